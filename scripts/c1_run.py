import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'deep-calcium_b200'))
import torch
from deepcalcium.engine import ops
x = torch.randn(8, 512, 512, device='cuda'); w = torch.randn(3, 3, 1, 32, device='cuda')
sc = torch.rand(32, device='cuda'); sh = torch.randn(32, device='cuda')
out = torch.empty(8, 512, 512, 32, dtype=torch.float16, device='cuda')
for _ in range(5): ops.conv3x3_c1_fwd(x, w, out, sc, sh, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.conv3x3_c1_fwd(x, w, out, sc, sh, True)
e1.record(); torch.cuda.synchronize(); print('c1 fwd 8x512x512x32: %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))
