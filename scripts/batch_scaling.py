"""Forward time of the 512^2 inference graph against the batch size (1, 2, 4, 8 images = the per-rank share of the 8 TTA
transforms on 8, 4, 2, 1 GPUs), CUDA-graph replay, plus the eager per-layer times at one batch size.

    python scripts/batch_scaling.py [precision] [per-layer batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200')); sys.path.insert(0, ROOT)
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np  # noqa: E402
import torch  # noqa: E402
from deepcalcium.engine.graph import GraphSpec, he_normal_weights  # noqa: E402
from deepcalcium.engine.unet_engine import UNetEngine  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
detail = int(sys.argv[2]) if len(sys.argv) > 2 else 1
spec = GraphSpec(32)
eng = UNetEngine(spec, precision=precision)
eng.set_weights_dict(he_normal_weights(spec, seed=7535))
for nb in (1, 2, 4, 8):
    x = torch.randn(nb, 512, 512, device='cuda')
    for _ in range(4):
        eng.infer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.infer(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print('batch %d: %.4f ms per forward (incl. the 1 MiB/img input copy), %.4f ms per image, %.0f TFLOP/s'
          % (nb, ms, ms / nb, nb * spec.flops_forward(512, 512) / ms / 1e9))
import bench  # noqa: E402
sess = eng._session(detail, 512, 512, False)
rows, tot_f, tot_ms = bench.per_layer_profile(eng, sess, spec, detail, 512, 512)
print('per layer at batch %d (eager, CUDA events):' % detail)
for r in rows:
    print('  %-36s %8.4f ms %s' % (r['op'], r['ms'], r['tflops']))
