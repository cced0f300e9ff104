"""Run one conv3x3 layer shape a few times (for ncu captures): python scripts/one_layer.py N H W Cin Cout [head|pool|plain]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import numpy as np, torch
from deepcalcium.engine import ops
from deepcalcium import _native as nat
for kv in filter(None, os.environ.get('DCB_POLICY', '').split(',')):     # e.g. DCB_POLICY=pair=1,strip=0
    nat.set_policy(**{kv.split('=')[0]: int(kv.split('=')[1])})
N, H, W, Cin, Cout = (int(v) for v in sys.argv[1:6])
mode = sys.argv[6] if len(sys.argv) > 6 else 'plain'
dt = torch.bfloat16
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn(N, H, W, Cin, device='cuda', generator=g).to(dt)
w = torch.randn(3, 3, Cin, Cout, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
wf = torch.empty(9 * Cin * Cout, dtype=dt, device='cuda')
ops.prep_conv3x3_weights(w, wf, None, dt)
scale = torch.rand(Cout, device='cuda') + 0.5; shift = torch.randn(Cout, device='cuda') * 0.1
y = torch.empty(N, H, W, Cout, dtype=dt, device='cuda')
hk = torch.randn(Cout, 2, device='cuda') * 0.3; hb = torch.tensor([0.1, -0.2], device='cuda')
logit = torch.empty(N, H, W, device='cuda'); prob = torch.empty(N, H, W, device='cuda')
pool = torch.empty(N, H // 2, W // 2, Cout, dtype=dt, device='cuda')
def run():
    if mode == 'head':
        ops.conv3x3_fwd_fused(x, None, wf, y, scale, shift, True, head_kernel=hk, head_bias=hb, logit=logit, prob=prob, need_y=False)
    elif mode == 'pool':
        ops.conv3x3_fwd_fused(x, None, wf, y, scale, shift, True, pool_out=pool)
    else:
        ops.conv3x3_fwd(x, None, wf, y, scale, shift, True)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
print('%s %s [%s]: %.4f ms per launch' % (sys.argv[1:6], mode, nat.last_kernel(), e0.elapsed_time(e1) / 10))
