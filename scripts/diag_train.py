"""Per-tensor gradient error of one training step vs the fp64 oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200')); sys.path.insert(0, ROOT)
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np, torch
import oracle
from deepcalcium.engine.graph import GraphSpec
from deepcalcium.engine.unet_engine import UNetEngine

shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '4,32,32').split(','))
spec = oracle.UNetSpec(32)
w = oracle.init_weights(spec, seed=7535)
rng = np.random.default_rng(865)
x = rng.standard_normal(shape).astype(np.float32)
y = (rng.random(shape) < 0.126).astype(np.uint8)
L, nw, st, g, _ = oracle.train_step(w, x, y, spec=spec, loss='dice_loss')
for precision in ('fp32', 'bf16'):
    eng = UNetEngine(GraphSpec(32), precision=precision, use_graphs=False)
    eng.set_weights_dict(w)
    m = eng.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), loss='dice_loss', dropout=False)
    print(precision, 'loss', float(m[0].item()), 'oracle', L)
    for key, g_ref in g.items():
        if key.endswith('/bias') and not key.startswith('head'):
            continue
        got = eng.G[key].cpu().numpy().astype(np.float64)
        rel = np.linalg.norm(got - g_ref) / (np.linalg.norm(g_ref) + 1e-30)
        cos = float((got * g_ref).sum() / (np.linalg.norm(got) * np.linalg.norm(g_ref) + 1e-30))
        print('  %-14s rel %.4f cos %.5f |g| %.3e' % (key, rel, cos, np.linalg.norm(g_ref)))
