"""CPU oracle for the deep-calcium UNet2DS hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the reported CPU baseline.

PARITY UNPINNED: the reference (alexklibisz/deep-calcium) ships no tests, no
golden vectors and its arithmetic lives in un-vendored Keras 2.0.6 /
TensorFlow 1.2.1, neither of which is installable here.  The oracle is a
restatement of the reference's graph/loss/projection code with the documented
Keras-2.0.6 layer semantics; it is pinned only by algebraic identities and by
self-generated fixtures under ``tests/golden`` (see ``oracle/make_golden.py``).
"""
from .projection import project_mean_max, project_streaming_fp16, summarize_series  # noqa: F401
from .unet import (  # noqa: F401
    LAYER_ORDER, UNetSpec, init_weights, weights_to_keras_list, keras_list_to_weights,
    unet_forward, tta_predict, reflect_pad, train_step, keras_adam_update,
)
from .losses import (  # noqa: F401
    dice_loss, dicesq_loss, binary_crossentropy, weighted_binary_crossentropy,
    batch_metrics, INVERTIBLE_2D_AUGMENTATIONS,
)
