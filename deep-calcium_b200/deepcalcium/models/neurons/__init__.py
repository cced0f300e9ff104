from .unet_2d_summary import UNet2DSummary, unet  # noqa: F401
