"""Time conv3x3_fwd on 8x512x512 (and 8x256x256) inputs for several channel splits: per-tile cost vs MMA count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import torch
from deepcalcium.engine import ops
bf = torch.bfloat16
def run(N, H, W, C0, C1, Cout, fused=None):
    x0 = torch.randn(N, H, W, C0, device='cuda').to(bf)
    x1 = torch.randn(N, H, W, C1, device='cuda').to(bf) if C1 else None
    w = torch.randn(3, 3, C0 + C1, Cout, device='cuda') * 0.05
    wf = torch.empty(9 * (C0 + C1) * Cout, dtype=bf, device='cuda')
    ops.prep_conv3x3_weights(w, wf, None, bf)
    y = torch.empty(N, H, W, Cout, dtype=bf, device='cuda')
    sc = torch.ones(Cout, device='cuda'); sh = torch.zeros(Cout, device='cuda')
    hk = torch.randn(Cout, 2, device='cuda'); hb = torch.zeros(2, device='cuda')
    lg = torch.empty(N, H, W, device='cuda'); pr = torch.empty(N, H, W, device='cuda')
    def call():
        if fused == 'head_noy':
            ops.conv3x3_fwd_fused(x0, x1, wf, y, sc, sh, True, head_kernel=hk, head_bias=hb, logit=lg, prob=pr, need_y=False)
        else:
            ops.conv3x3_fwd(x0, x1, wf, y, sc, sh, True)
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * N * H * W * 9 * (C0 + C1) * Cout
    print('N%d %dx%d C0=%d C1=%d Cout=%d %s: %.4f ms  %.0f TFLOP/s' % (N, H, W, C0, C1, Cout, fused or '', ms, fl / ms / 1e9), flush=True)
for cfg in [(8, 512, 512, 32, 0, 32), (8, 512, 512, 32, 32, 32), (8, 512, 512, 64, 32, 32), (8, 512, 512, 64, 0, 32), (8, 512, 512, 64, 64, 32),
            (8, 512, 512, 32, 0, 64), (8, 512, 512, 64, 0, 64), (8, 256, 256, 64, 0, 64), (8, 256, 256, 64, 64, 64)]:
    run(*cfg)
run(8, 512, 512, 32, 0, 32, 'head_noy')
