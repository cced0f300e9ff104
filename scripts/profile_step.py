"""One eager (non-graph) step of a hot-path workload, for ncu launch lists / captures.

    python scripts/profile_step.py infer|train|proj [precision]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np  # noqa: E402
import torch  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'infer'
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
from deepcalcium.engine.graph import GraphSpec, he_normal_weights  # noqa: E402
from deepcalcium.engine.unet_engine import UNetEngine  # noqa: E402

if what == 'proj':
    from deepcalcium.datasets.nf import summarize_movie_device
    movie = torch.rand((3000, 512, 512), device='cuda') * 4096
    for _ in range(3):
        summarize_movie_device(movie)
    torch.cuda.synchronize()
    sys.exit(0)

spec = GraphSpec(32)
eng = UNetEngine(spec, precision=precision, use_graphs=False)
eng.set_weights_dict(he_normal_weights(spec, seed=7535))
rng = np.random.default_rng(865)
if what == 'infer':
    s = torch.from_numpy(rng.standard_normal((512, 512)).astype(np.float32)).cuda()
    for _ in range(3):
        eng.predict_tta(s)
else:
    x = torch.from_numpy(rng.standard_normal((32, 128, 128)).astype(np.float32)).cuda()
    y = torch.from_numpy((rng.random((32, 128, 128)) < 0.126).astype(np.uint8)).cuda()
    for _ in range(3):
        eng.train_step(x, y, loss='dice_loss', dropout=True)
torch.cuda.synchronize()
print('launches', eng.launches)
