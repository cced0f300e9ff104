"""Per-tensor gradient error of one training step in 'upsampling' mode vs the fp64 oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200')); sys.path.insert(0, ROOT)
import numpy as np, torch
import oracle
from deepcalcium.engine.graph import GraphSpec
from deepcalcium.engine.unet_engine import UNetEngine
spec = oracle.UNetSpec(32, upsampling_or_transpose='upsampling')
w = oracle.init_weights(spec, seed=7535)
rng = np.random.default_rng(865)
x = rng.standard_normal((2, 64, 64)).astype(np.float32)
y = (rng.random((2, 64, 64)) < 0.126).astype(np.uint8)
L, nw, st, g, _ = oracle.train_step(w, x, y, spec=spec, loss='dice_loss')
L32, _, _, g32, _ = oracle.train_step(w, x, y, spec=spec, loss='dice_loss', dtype=torch.float32)
eng = UNetEngine(GraphSpec(32, upsampling_or_transpose='upsampling'), precision='fp32', use_graphs=False)
eng.set_weights_dict(w)
m = eng.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), loss='dice_loss', dropout=False)
print('loss', float(m[0].item()), L)
for key, g_ref in g.items():
    if key.endswith('/bias') and not key.startswith('head'):
        continue
    got = eng.G[key].cpu().numpy().astype(np.float64)
    rel = np.linalg.norm(got - g_ref) / (np.linalg.norm(g_ref) + 1e-30)
    rel32 = np.linalg.norm(g32[key] - g_ref) / (np.linalg.norm(g_ref) + 1e-30)
    print('  %-14s gpu-fp32 rel %.5f | cpu-fp32 rel %.5f | |g| %.3e' % (key, rel, rel32, np.linalg.norm(g_ref)))
