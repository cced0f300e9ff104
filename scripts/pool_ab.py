import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'deep-calcium_b200'))
import torch
from deepcalcium.engine import ops
from deepcalcium import _native as nat
dt = torch.bfloat16
def bench(N,H,W,Cin,Cout,pool,pol):
    x = torch.randn(N,H,W,Cin,device='cuda').to(dt); w = torch.randn(3,3,Cin,Cout,device='cuda')*(2/(9*Cin))**.5
    wf = torch.empty(9*Cin*Cout,dtype=dt,device='cuda'); ops.prep_conv3x3_weights(w,wf,None,dt)
    sc = torch.rand(Cout,device='cuda')+.5; sh = torch.randn(Cout,device='cuda')*.1
    y = torch.empty(N,H,W,Cout,dtype=dt,device='cuda'); p = torch.empty(N,H//2,W//2,Cout,dtype=dt,device='cuda')
    def run():
        if pool == 'fused': ops.conv3x3_fwd_fused(x,None,wf,y,sc,sh,True,pool_out=p)
        elif pool == 'sep': ops.conv3x3_fwd(x,None,wf,y,sc,sh,True); ops.maxpool2x2(y,p)
        else: ops.conv3x3_fwd(x,None,wf,y,sc,sh,True)
    with nat.policy(**pol):
        for _ in range(3): run()
        k = nat.last_kernel()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10): run()
        g.replay(); torch.cuda.synchronize()
        e0,e1 = torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
    print('%-28s pool=%-5s policy=%-22s kernel=%-18s %.1f us' % ((N,H,W,Cin,Cout),pool,pol,k,e0.elapsed_time(e1)/50*1e3))
for shape in [(8,128,128,128,128),(8,256,256,64,64),(8,64,64,256,256)]:
    for pol in ({}, {'swap_min_cout':0}):
        for pool in ('none','sep','fused'):
            bench(*shape,pool,pol)
