"""Device times of the memory-bound training kernels (BN reductions / applies) at the training step's shapes.

Inputs are rotated over > L2-sized pools so every launch reads HBM.  Prints achieved GB/s (algorithmic bytes).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import torch
from deepcalcium.engine import ops

torch.manual_seed(0)
dev = 'cuda'
shapes = [(32 * 128 * 128, 32), (32 * 64 * 64, 64), (32 * 32 * 32, 128), (32 * 16 * 16, 256), (32 * 8 * 8, 512)]
print('ENV', {k: v for k, v in os.environ.items() if k.startswith('DCB_')})
for M, C in shapes:
    nbuf = max(2, int(400e6 // (M * C * 6)) + 1)
    xs = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    dys = [torch.randn(M, C, device=dev) for _ in range(nbuf)]
    ys = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    scale = torch.rand(C, device=dev) + 0.5; shift = torch.randn(C, device=dev) * 0.1
    mean = torch.randn(C, device=dev) * 0.1; rstd = torch.rand(C, device=dev) + 0.5
    dg = torch.empty(C, device=dev); db = torch.empty(C, device=dev)

    def timeit(fn, bytes_):
        n = 20
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()           # graph replay: no host launch overhead between the kernels
        with torch.cuda.graph(g):
            for i in range(n):
                fn(i % nbuf)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / n
        return '%6.1f us %5.0f GB/s' % (us, bytes_ / us / 1e3)

    r = []
    r.append('stats ' + timeit(lambda i: ops.bn_stats(xs[i], sums), M * C * 2))
    r.append('apply ' + timeit(lambda i: ops.bn_apply(xs[i], scale, shift, ys[i]), M * C * 4))
    r.append('bwd_reduce ' + timeit(lambda i: ops.bn_bwd_reduce(dys[i], C, 0, xs[i], scale, shift, mean, rstd, sums), M * C * 6))
    r.append('bwd_apply ' + timeit(lambda i: ops.bn_bwd_apply(dys[i], C, 0, xs[i], scale, shift, mean, rstd, sums, ys[i], dg, db), M * C * 8))
    print('M=%7d C=%3d | ' % (M, C) + ' | '.join(r))
    del xs, dys, ys
