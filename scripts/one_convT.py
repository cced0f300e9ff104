"""Run one convT2x2 (stride 2) layer shape a few times (timing / ncu captures): python scripts/one_convT.py N h w Cin Cout [f16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import torch
from deepcalcium.engine import ops
N, h, w, Cin, Cout = (int(v) for v in sys.argv[1:6])
dt = torch.float16 if len(sys.argv) > 6 and sys.argv[6] == 'f16' else torch.bfloat16
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn(N, h, w, Cin, device='cuda', generator=g).to(dt)
k = torch.randn(2, 2, Cout, Cin, device='cuda', generator=g) * (1.0 / Cin) ** 0.5
wf = torch.empty(4 * Cin * Cout, dtype=dt, device='cuda')
wd = torch.empty(4 * Cin * Cout, dtype=dt, device='cuda')
ops.prep_convT2x2_weights(k, wf, wd, dt)
scale = torch.rand(Cout, device='cuda') + 0.5; shift = torch.randn(Cout, device='cuda') * 0.1
y = torch.empty(N, 2 * h, 2 * w, Cout, dtype=dt, device='cuda')
run = lambda: ops.convT2x2_fwd(x, wf, y, scale, shift, True)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
mb = (x.numel() + y.numel()) * 2 / 1e6
print('%s convT: %.4f ms per launch, %.0f MB -> %.0f GB/s' % (sys.argv[1:6], e0.elapsed_time(e1) / 10, mb, mb / (e0.elapsed_time(e1) / 10) ))
