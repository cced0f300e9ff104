"""Layer-by-layer deviation of the bf16 inference activations from the bf16-emulating oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200')); sys.path.insert(0, ROOT)
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np, torch
import oracle
from deepcalcium.engine.graph import GraphSpec
from deepcalcium.engine.unet_engine import UNetEngine

spec = oracle.UNetSpec(32)
w = oracle.init_weights(spec, seed=7535)
x = np.random.default_rng(865).standard_normal((2, 64, 64)).astype(np.float32)
o16 = oracle.unet_forward(w, x, spec, dtype=torch.float64, emulate_bf16=True, return_intermediates=True)
o64 = oracle.unet_forward(w, x, spec, dtype=torch.float64, return_intermediates=True)
eng = UNetEngine(GraphSpec(32), precision='bf16', use_graphs=False)
eng.set_weights_dict(w)
_, logit = eng.infer(torch.from_numpy(x).cuda())
act = eng._sessions[(2, 64, 64, False)]['act']
for name in [b[0] for b in spec.blocks() if b[1] != 'head']:
    got = act[name].float().cpu().permute(0, 3, 1, 2).double()
    r16, r64 = o16['intermediates'][name], o64['intermediates'][name]
    d16, d64 = (got - r16).abs(), (got - r64).abs()
    exact = float((got == r16).double().mean())
    print('%-6s vs emu: max %.4f mean %.6f exact %.4f | vs f64: max %.4f mean %.6f | emu vs f64 mean %.6f | rms act %.3f'
          % (name, d16.max(), d16.mean(), exact, d64.max(), d64.mean(), (r16 - r64).abs().mean(), float(r64.pow(2).mean().sqrt())))
print('logit vs emu', float((logit.cpu().double() - o16['logit']).abs().mean()), 'vs f64', float((logit.cpu().double() - o64['logit']).abs().mean()))
