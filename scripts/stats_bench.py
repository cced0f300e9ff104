"""Training-forward conv kernels with and without the batch-statistics epilogue, and the BatchNorm launch that follows
(single-launch kernel with grid barriers vs one pass from known sums), per layer shape of a 32 x 128^2 step.
    python scripts/stats_bench.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from deepcalcium import _native as nat  # noqa: E402
from deepcalcium.engine import ops  # noqa: E402

bf = torch.bfloat16
nat.set_policy(swap_min_cout=0)


def timeit(fn, n=20):
    """device time per call: n calls captured into one CUDA graph (no host launch overhead), best of 5 replays"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def rnd(*shape):
    return torch.randn(*shape, device='cuda').to(bf)


LAYERS = [('conv', 32, 128, 32, 0, 32), ('conv', 32, 128, 32, 32, 32), ('conv', 32, 64, 32, 0, 64), ('conv', 32, 64, 64, 0, 64),
          ('conv', 32, 64, 64, 64, 64), ('conv', 32, 32, 64, 0, 128), ('conv', 32, 32, 128, 0, 128), ('conv', 32, 32, 128, 128, 128),
          ('conv', 32, 16, 128, 0, 256), ('conv', 32, 16, 256, 0, 256), ('conv', 32, 16, 256, 256, 256), ('conv', 32, 8, 256, 0, 512),
          ('conv', 32, 8, 512, 0, 512), ('convT', 32, 8, 512, 256), ('convT', 32, 16, 256, 128), ('convT', 32, 32, 128, 64),
          ('convT', 32, 64, 64, 32), ('c1', 32, 128, 32)]
ws_bn = torch.zeros(ops.bn_train_workspace_bytes(512), dtype=torch.uint8, device='cuda')
for L in LAYERS:
    kind = L[0]
    if kind == 'conv':
        _, N, S, C0, C1, Co = L
        x0, x1 = rnd(N, S, S, C0), (rnd(N, S, S, C1) if C1 else None)
        wf = rnd(9 * (C0 + C1) * Co) * 0.05
        out = torch.empty(N, S, S, Co, dtype=bf, device='cuda')
        plain = lambda: ops.conv3x3_fwd(x0, x1, wf, out, None, bias, False)
        stats = lambda: (sums.zero_(), ops.conv3x3_fwd_stats(x0, x1, wf, out, sums, None, bias, False))
    elif kind == 'convT':
        _, N, S, Ci, Co = L
        x0 = rnd(N, S, S, Ci)
        wf = rnd(4 * Ci * Co) * 0.05
        out = torch.empty(N, 2 * S, 2 * S, Co, dtype=bf, device='cuda')
        plain = lambda: ops.convT2x2_fwd(x0, wf, out, None, bias, False)
        stats = lambda: (sums.zero_(), ops.convT2x2_fwd_stats(x0, wf, out, sums, None, bias, False))
    else:
        _, N, S, Co = L
        x0 = torch.randn(N, S, S, device='cuda')
        wk = torch.randn(3, 3, 1, Co, device='cuda')
        out = torch.empty(N, S, S, Co, dtype=bf, device='cuda')
        plain = lambda: ops.conv3x3_c1_fwd(x0, wk, out, None, bias, False)
        stats = lambda: (sums.zero_(), ops.conv3x3_c1_fwd_stats(x0, wk, out, sums, None, bias, False))
    bias = torch.randn(Co, device='cuda')
    sums = torch.zeros(2 * Co, dtype=torch.int64, device='cuda')
    g, b = torch.ones(Co, device='cuda'), torch.zeros(Co, device='cuda')
    f = lambda: torch.empty(Co, device='cuda')
    mm, mv, sc, sh, me, rs = f(), f(), f(), f(), f(), f()
    y = torch.empty_like(out)
    sync = torch.zeros(4, dtype=torch.int32, device='cuda')

    def bn_full():
        sync.zero_()
        ops.bn_train_fwd(out, g, b, 0.99, mm, mv, sc, sh, me, rs, y, ws_bn, sync, True, 0.0, 1, None, 1)

    def bn_sums():
        ops.bn_train_fwd_sums(out, sums, g, b, 0.99, mm, mv, sc, sh, me, rs, y, True, 0.0, 1, None, 1)
    tz = timeit(lambda: sync.zero_())
    tp = timeit(plain); kp = nat.last_kernel()
    ts = timeit(stats); ks = nat.last_kernel()
    ts -= tz            # the stats variant's timing includes zeroing its sums (one fill per STEP in the engine)
    tb, tbs = timeit(bn_full) - tz, timeit(bn_sums)
    print('%-34s conv %6.1f us (%s)  +stats %6.1f us (%s) | bn single-launch %6.1f us  from sums %6.1f us | pair %6.1f -> %6.1f'
          % (str(L), tp, kp, ts, ks, tb, tbs, tp + tb, ts + tbs))
