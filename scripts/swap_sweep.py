"""Whole-step effect of the orientation policy (dcb_set_policy SWAP_MIN_COUT) on the 8x-TTA inference step and the 32-crop
training step: 64 = weights-as-M for 64 <= Cout <= 128 (round-1 default), 128 = only Cout = 128, 0 = never."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200')); sys.path.insert(0, ROOT)
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np  # noqa: E402
import torch  # noqa: E402
from deepcalcium import _native as nat  # noqa: E402
from deepcalcium.engine.graph import GraphSpec, he_normal_weights  # noqa: E402
from deepcalcium.engine.unet_engine import UNetEngine  # noqa: E402

spec = GraphSpec(32)
w = he_normal_weights(spec, seed=7535)
rng = np.random.default_rng(865)
img = torch.from_numpy(rng.standard_normal((512, 512)).astype(np.float32)).cuda()
x = torch.from_numpy(rng.standard_normal((32, 128, 128)).astype(np.float32)).cuda()
y = torch.from_numpy((rng.random((32, 128, 128)) < 0.126).astype(np.uint8)).cuda()


def timed(fn, n):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


SETS = {'A': ('enc0b',), 'B': ('enc0b', 'enc1b'), 'C': ('enc0b', 'enc2b'), 'D': ('enc0b', 'enc3b'), 'E': ('enc0b', 'enc1b', 'enc2b', 'enc3b'),
        'none': ()}
combos = [(dict(swap_min_cout=64), 'A'), (dict(swap_min_cout=0), 'A'), (dict(swap_min_cout=0), 'B'), (dict(swap_min_cout=0), 'C'),
          (dict(swap_min_cout=0), 'D'), (dict(swap_min_cout=0), 'E'), (dict(swap_min_cout=0), 'none')]
for pol, fs in combos:
    nat.reset_policy(); nat.set_policy(**pol)
    e1 = UNetEngine(spec, precision='fp16'); e1.set_weights_dict(w)
    e1._pool_fused = SETS[fs]
    pol = dict(pol, pools_fused=fs)
    ms_i = timed(lambda: e1.predict_tta(img), 30)
    e2 = UNetEngine(spec, precision='bf16'); e2.set_weights_dict(w)
    ms_t = timed(lambda: e2.train_step(x, y, loss='dice_loss', dropout=True), 20)
    print('%-44s inference %.4f ms/step (%.0f img/s)   training %.4f ms/step (%.0f crops/s)' % (pol, ms_i, 1e3 / ms_i, ms_t, 32e3 / ms_t))
    del e1, e2
