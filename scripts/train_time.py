"""Time the C3 training step (32 crops of 128x128, bf16, dice, Adam, dropout on) through the CUDA graph: ms per step and the loss
after 25 steps.    python scripts/train_time.py [policy=value ...]      env: DCB_PDL_TRAIN=0|1"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from deepcalcium import _native as nat  # noqa: E402
from deepcalcium.engine.graph import GraphSpec, he_normal_weights  # noqa: E402
from deepcalcium.engine.unet_engine import UNetEngine  # noqa: E402

for a in sys.argv[1:]:
    k, v = a.split('=')
    nat.set_policy(**{k: int(v)})
# what-if experiments (timing only, results are wrong): TT_SKIP=wgrad,bn_fwd,bn_bwd replaces those C-ABI calls by no-ops
from deepcalcium.engine import ops  # noqa: E402
_skip = [t for t in os.environ.get('TT_SKIP', '').split(',') if t]
for name in dir(ops):
    if any(t in name for t in _skip) and callable(getattr(ops, name)) and 'workspace' not in name:
        setattr(ops, name, lambda *a, **k: None)
spec = GraphSpec(32)
eng = UNetEngine(spec, precision='bf16')
eng.overlap_wgrad = os.environ.get('TT_OVERLAP', '1') != '0'
eng.set_weights_dict(he_normal_weights(spec, seed=7535))
rng = np.random.default_rng(865)
xs = [torch.from_numpy(rng.standard_normal((32, 128, 128)).astype(np.float32)).cuda() for _ in range(4)]
ys = [torch.from_numpy((rng.random((32, 128, 128)) < 0.126).astype(np.uint8)).cuda() for _ in range(4)]
for i in range(5):
    eng.train_step(xs[i % 4], ys[i % 4], loss='dice_loss', lr=0.002, dropout=True)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        m = eng.train_step(xs[i % 4], ys[i % 4], loss='dice_loss', lr=0.002, dropout=True)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print('train step: %.4f ms (best of 3 x 20), skip=%s overlap=%s, loss after %d steps %.6f, pdl_train=%s, policy %s'
      % (best, _skip, eng.overlap_wgrad, eng.iteration, float(m[0]), eng.pdl_train, {k: nat.get_policy(k) for k in ('fused_bn', 'bn_slab', 'pdl')}))
