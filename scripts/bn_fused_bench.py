"""Training BatchNorm: single-launch kernels (dcb_bn_train_fwd / _bwd) against the separate passes, per layer shape of a
32 x 128^2 training step.  Times are per call, averaged over a loop of back-to-back launches (CUDA events around the loop).

    python scripts/bn_fused_bench.py [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
import torch  # noqa: E402
from deepcalcium.engine import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
PD = float(os.environ.get('BN_PDROP', '0.25'))   # dropout rate of the timed calls (7 of the 22 layers of a step have one)
only = os.environ.get('BN_ONLY')          # e.g. "32,8,8,512" to run a single shape (ncu captures)
dt = torch.bfloat16
shapes = [(32, 128, 128, 32), (32, 64, 64, 64), (32, 32, 32, 128), (32, 16, 16, 256), (32, 8, 8, 512)]
if only:
    shapes = [tuple(int(v) for v in only.split(','))]


def timed(fn, inner=20):
    """per-call device time with the launches replayed from a CUDA graph (no host launch overhead)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(1, reps // inner)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * inner) * 1e3


from deepcalcium import _native as nat  # noqa: E402
per_sm_list = [int(v) for v in os.environ.get('BN_CTAS', '4').split(',')]
nat.set_policy(bn_slab=int(os.environ.get('BN_SLAB', '1')))
for N, H, W, C in [sh for sh in shapes for _ in per_sm_list]:
    per_sm = per_sm_list[0]; per_sm_list = per_sm_list[1:] + per_sm_list[:1]
    nat.set_policy(bn_ctas_per_sm=per_sm)
    M = N * H * W
    x = torch.randn(N, H, W, C, device='cuda').to(dt)
    y = torch.empty_like(x); pool = torch.empty(N, H // 2, W // 2, C, dtype=dt, device='cuda')
    dy = torch.randn(M, C, device='cuda')
    draw = torch.empty_like(x)
    f = lambda: torch.empty(C, device='cuda')
    gamma, beta = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    sc, sh, mu, rs, dg, db = f(), f(), f(), f(), f(), f()
    sums = torch.zeros(4 * C, dtype=torch.float64, device='cuda')
    ws = torch.zeros(ops.bn_train_workspace_bytes(C), dtype=torch.uint8, device='cuda')
    sync = torch.zeros(8, dtype=torch.int32, device='cuda')
    seed_dev = torch.tensor([1], dtype=torch.int64, device='cuda')

    def sep_fwd():
        sums.zero_()
        ops.bn_stats(x, sums[:2 * C])
        ops.bn_finalize_apply(x, sums[:2 * C], M, gamma, beta, 0.99, None, None, sc, sh, mu, rs, y, True, PD, 7, seed_dev, 3)

    def fused_fwd():
        sync.zero_()
        ops.bn_train_fwd(x, gamma, beta, 0.99, None, None, sc, sh, mu, rs, y, ws, sync[:4], True, PD, 7, seed_dev, 3)

    def fused_fwd_pool():
        sync.zero_()
        ops.bn_train_fwd(x, gamma, beta, 0.99, None, None, sc, sh, mu, rs, y, ws, sync[:4], True, PD, 7, seed_dev, 3, pool_out=pool)

    def sep_bwd():
        sums.zero_()
        ops.bn_bwd_reduce(dy, C, 0, x, sc, sh, mu, rs, sums[2 * C:], PD, 7, seed_dev, 3)
        ops.bn_bwd_apply(dy, C, 0, x, sc, sh, mu, rs, sums[2 * C:], draw, dg, db, PD, 7, seed_dev, 3)

    def fused_bwd():
        sync.zero_()
        ops.bn_train_bwd(dy, C, 0, x, sc, sh, mu, rs, draw, dg, db, ws, sync[4:], PD, 7, seed_dev, 3)

    sep_fwd()
    mb = M * C * 2 / 1e6
    r = dict(sep_fwd=timed(sep_fwd), fused_fwd=timed(fused_fwd), fused_fwd_pool=timed(fused_fwd_pool), sep_bwd=timed(sep_bwd),
             fused_bwd=timed(fused_bwd), memset=timed(lambda: sync.zero_()))
    print('ctas/SM %d slab %d p_drop %.2f ' % (per_sm, nat.get_policy('bn_slab'), PD), end='')
    print('%-20s %6.1f MB bf16 | fwd: separate %6.1f us, fused %6.1f us, fused+pool %6.1f us | bwd: separate %6.1f us, fused %6.1f us | '
          'memset alone %.1f us | floors (HBM 6.5 TB/s): fwd %.1f us (R+W), bwd %.1f us (R dy fp32 + R x + W)'
          % ((N, H, W, C), mb, r['sep_fwd'], r['fused_fwd'], r['fused_fwd_pool'], r['sep_bwd'], r['fused_bwd'], r['memset'],
             2 * mb / 6.5, 4 * mb / 6.5))
