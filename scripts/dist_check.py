"""Multi-GPU parity check (run with torchrun, one rank per GPU):
  1. 8x TTA of one image sharded over the ranks == the single-GPU mask, bit for bit;
  2. data-parallel training (SyncBN sums, loss sums, gradient all-reduce) == single-GPU training on the
     concatenated batch (fp32 check mode, dropout off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200')); sys.path.insert(0, ROOT)
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np, torch, torch.distributed as dist
from deepcalcium.engine.graph import GraphSpec, he_normal_weights
from deepcalcium.engine.unet_engine import UNetEngine
from deepcalcium.engine.dist import Comm, predict_tta_sharded, shard_range, sync_parameters, attach_peers

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
comm = Comm()
spec = GraphSpec(32)
w = he_normal_weights(spec, seed=7535)
rng = np.random.default_rng(7535)
for blk in spec.blocks:
    if blk.kind != 'head':
        w[blk.name + '/moving_mean'] = (0.1 * rng.standard_normal(blk.cout)).astype(np.float32)
        w[blk.name + '/moving_var'] = rng.uniform(0.5, 1.5, blk.cout).astype(np.float32)
ok = True
# ---- 1. sharded TTA (bf16 tensor-core mode)
eng = UNetEngine(spec, precision='bf16')
eng.set_weights_dict(w)
s = torch.from_numpy(np.random.default_rng(865).standard_normal((500, 480)).astype(np.float32)).cuda()
for it in range(4):     # eager, capture, replay, replay
    mask, act = predict_tta_sharded(eng, s if comm.rank == 0 else None, comm, shape=tuple(s.shape))
if comm.rank == 0:
    mask, act = mask.clone(), act.clone()
    m1, a1 = eng.predict_tta(s)
    same = bool(torch.equal(mask, m1)) and bool(torch.equal(act, a1))
    print('sharded TTA over %d ranks bit-identical to 1 GPU: %s' % (comm.world, same))
    ok &= same
# ---- 2. data-parallel training (fp32 check mode)
B, H = 8, 32
x = np.random.default_rng(1).standard_normal((B, H, H)).astype(np.float32)
y = (np.random.default_rng(2).random((B, H, H)) < 0.126).astype(np.uint8)
f, c = shard_range(B, comm.world, comm.rank)
dp = UNetEngine(spec, precision='fp32', use_graphs=False)
dp.set_weights_dict(w)
dp.comm = comm
if os.environ.get('DCB_DP_PEERS', '1') == '1':
    attach_peers(dp, comm)
sync_parameters(dp, comm)
m = dp.train_step(torch.from_numpy(x[f:f + c]).cuda(), torch.from_numpy(y[f:f + c]).cuda(), loss='dice_loss', dropout=False)
loss_dp = float(m[0].item())
if comm.rank == 0:
    from deepcalcium import _native as nat
    ref = UNetEngine(spec, precision='fp32', use_graphs=False)
    ref.set_weights_dict(w)
    # default dispatch on the single device (channel-slab cluster BatchNorm kernels on the small tensors, grid-barrier
    # kernels in the data-parallel run): every BatchNorm sum of the check mode is an exact fixed-point integer sum, so the
    # kernel family and the split of the batch over ranks do not change it
    loss_ref = float(ref.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), loss='dice_loss', dropout=False)[0].item())
    # Adam's first step is lr * sign(g) for every weight, so weights with |g| ~ 0 are not comparable between two
    # summation orders; the all-reduced GRADIENTS and the BN moving statistics are.
    worst = ('', 0.0)
    for k in dp.G:
        a_, b_ = dp.G[k].double().cpu().numpy(), ref.G[k].double().cpu().numpy()
        r = float(np.linalg.norm(a_ - b_) / (np.linalg.norm(b_) + 1e-30)) if np.linalg.norm(b_) > 0 else float(np.abs(a_).max())
        if r > worst[1]:
            worst = (k, r)
    wd, wr = dp.get_weights_dict(), ref.get_weights_dict()
    stat = max(float(np.max(np.abs(wd[k] - wr[k]))) for k in wd if 'moving' in k)
    print('DP loss %.8f single-GPU loss %.8f; worst gradient rel. L2 diff %s %.3g; max |BN moving stat diff| %.3g'
          % (loss_dp, loss_ref, worst[0], worst[1], stat))
    good = abs(loss_dp - loss_ref) < 1e-5 and worst[1] < 1e-4 and stat < 1e-5
    print('data-parallel training matches the single-device batch: %s' % good)
    ok &= bool(good)
# ---- 3. the same in the bf16 tensor-core mode under the default policy (batch statistics from the conv epilogues, in-kernel
# SyncBN exchange of the fixed-point sums): against the single-device bf16 step on the concatenated batch.  The two runs
# round the same fp32 accumulations at different tile boundaries, so the comparison is at bf16 tolerance; the BatchNorm
# moving statistics (the exchanged sums) must agree to fp32 rounding.  B = 32 crops of 64 x 64 so that the level-0..2
# tensors pass the size gate of the statistics epilogue.
B, H = 32, 64
x = np.random.default_rng(3).standard_normal((B, H, H)).astype(np.float32)
y = (np.random.default_rng(4).random((B, H, H)) < 0.126).astype(np.uint8)
f, c = shard_range(B, comm.world, comm.rank)
dpb = UNetEngine(spec, precision='bf16', use_graphs=False)
dpb.set_weights_dict(w)
dpb.comm = comm
attach_peers(dpb, comm)
sync_parameters(dpb, comm)
from deepcalcium import _native as nat
with nat.policy(fused_bn=3):
    m = dpb.train_step(torch.from_numpy(x[f:f + c]).cuda(), torch.from_numpy(y[f:f + c]).cuda(), loss='dice_loss', dropout=False)
loss_dp = float(m[0].item())
if comm.rank == 0:
    ref = UNetEngine(spec, precision='bf16', use_graphs=False)
    ref.set_weights_dict(w)
    with nat.policy(fused_bn=3):
        loss_ref = float(ref.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), loss='dice_loss', dropout=False)[0].item())
    errs = {}
    for k in dpb.G:
        a_, b_ = dpb.G[k].double().cpu().numpy(), ref.G[k].double().cpu().numpy()
        if np.linalg.norm(b_) > 0:
            errs[k] = float(np.linalg.norm(a_ - b_) / np.linalg.norm(b_))
    wd, wr = dpb.get_weights_dict(), ref.get_weights_dict()
    stat = max(float(np.max(np.abs(wd[k] - wr[k]))) for k in wd if 'moving' in k)
    worst = max(errs, key=errs.get)
    # (two 16-bit evaluations of this random-init dice network that round differently anywhere - here: other kernels for the
    # smaller per-rank tensors - differ by tens of percent in the gradients, like each of them differs from fp64: DESIGN.md 4
    # "bf16 numerics"; loss and BatchNorm statistics are the meaningful agreement in this mode)
    print('bf16 DP (epilogue statistics + in-kernel SyncBN) loss %.6f single-GPU loss %.6f; gradient rel. L2 diff worst %s %.3g, median %.3g '
          '(16-bit noise floor of this network); max |BN moving stat diff| %.3g'
          % (loss_dp, loss_ref, worst, errs[worst], float(np.median(list(errs.values()))), stat))
    good = abs(loss_dp - loss_ref) < 2e-3 and stat < 2e-2
    print('bf16 data-parallel step consistent with the single-device batch: %s' % good)
    ok &= bool(good)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
