"""int16 projection: ms per 3000x512x512 movie for a few T-split counts (dcb_set_policy(DCB_POLICY_PROJ_I16_SPLITS))."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import torch
from deepcalcium.datasets.nf import summarize_movie_device
from deepcalcium.engine import ops
from deepcalcium import _native as nat
T, H, W = 3000, 512, 512
movie = (torch.rand((T, H, W), device='cuda') * 4096).to(torch.int16)
out = (torch.empty(H, W, device='cuda'), torch.empty(H, W, device='cuda'))
ws = torch.empty(ops.proj_workspace_bytes(T, H, W), dtype=torch.uint8, device='cuda')
for S in (0, 2, 4, 6, 8, 12, 17, 24, 32):
    nat.set_policy(proj_i16_splits=S)
    for _ in range(3): summarize_movie_device(movie, out=out, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): summarize_movie_device(movie, out=out, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('splits %2s: %.4f ms  %.0f GB/s' % (S or 'default', ms, (T * H * W * 2 + 2 * H * W * 4) / ms / 1e6))
