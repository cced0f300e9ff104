from .graph import GraphSpec  # noqa: F401
