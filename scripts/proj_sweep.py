"""Time every projection variant x T-split on the C2 shape (3000x512x512 fp32); prints GB/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import torch  # noqa: E402
from deepcalcium.datasets.nf import summarize_movie_device  # noqa: E402
from deepcalcium.engine import ops  # noqa: E402

T, H, W = 3000, 512, 512
movie = torch.rand((T, H, W), device='cuda') * 4096
out = (torch.empty(H, W, device='cuda'), torch.empty(H, W, device='cuda'))
ws = torch.empty(ops.proj_workspace_bytes(T, H, W), dtype=torch.uint8, device='cuda')
nbytes = T * H * W * 4 + 2 * H * W * 4
res = []
variants = [int(v) for v in os.environ.get('VARIANTS', '0,1,2,3').split(',')]
for variant in variants:
    for splits in (1, 2, 4, 8, 16, 32):
        for _ in range(2):
            summarize_movie_device(movie, variant=variant, t_splits=splits, out=out, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            summarize_movie_device(movie, variant=variant, t_splits=splits, out=out, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res.append(dict(variant=variant, splits=splits, ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1)))
        print(res[-1], flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'proj_sweep.json'), 'w'), indent=1)
