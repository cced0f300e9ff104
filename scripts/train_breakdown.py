"""Per-call device time of one eager training step (CUDA events around every C-ABI call), grouped by op.

    python scripts/train_breakdown.py [bf16|fp32] [policy=value ...]      e.g.  fused_bn=0

Unlike an ncu launch list the caches stay warm and the kernels are not serialised by a profiler; unlike the graph replay
the launches are issued one by one (the gaps between calls are not counted: each call is bracketed by its own events).
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np  # noqa: E402
import torch  # noqa: E402
from deepcalcium import _native as nat  # noqa: E402
from deepcalcium.engine import ops  # noqa: E402
from deepcalcium.engine.graph import GraphSpec, he_normal_weights  # noqa: E402
from deepcalcium.engine.unet_engine import UNetEngine  # noqa: E402

precision = 'bf16'
for a in sys.argv[1:]:
    if '=' in a:
        k, v = a.split('=')
        nat.set_policy(**{k: int(v)})
    else:
        precision = a
mode = os.environ.get('BREAKDOWN_MODE', 'train')

events = []
skip = {'proj_workspace_bytes', 'conv3x3_wgrad_workspace_bytes', 'convT2x2_wgrad_workspace_bytes',
        'conv3x3_c1_wgrad_workspace_bytes', 'bn_train_workspace_bytes'}
for name in dir(ops):
    f = getattr(ops, name)
    if callable(f) and not name.startswith('_') and getattr(f, '__module__', '') == ops.__name__ and name not in skip \
            and not isinstance(f, type):
        def make(name, f):
            def g(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = f(*a, **k)
                e1.record()
                shape = next((tuple(t.shape) for t in a if torch.is_tensor(t) and t.dim() >= 3), ())
                events.append((name, nat.last_kernel() if 'conv' in name or 'wgrad' in name else '', shape, e0, e1))
                return r
            return g
        setattr(ops, name, make(name, f))

spec = GraphSpec(32)
eng = UNetEngine(spec, precision=precision, use_graphs=False)
eng.set_weights_dict(he_normal_weights(spec, seed=7535))
rng = np.random.default_rng(865)
if mode == 'train':
    x = torch.from_numpy(rng.standard_normal((32, 128, 128)).astype(np.float32)).cuda()
    y = torch.from_numpy((rng.random((32, 128, 128)) < 0.126).astype(np.uint8)).cuda()
    run = lambda: eng.train_step(x, y, loss='dice_loss', dropout=True)
else:
    s = torch.from_numpy(rng.standard_normal((512, 512)).astype(np.float32)).cuda()
    run = lambda: eng.predict_tta(s)
for _ in range(3):
    del events[:]
    run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
tot = 0.0
for name, kern, shape, e0, e1 in events:
    us = e0.elapsed_time(e1) * 1e3
    tot += us
    key = name + (':' + kern if kern else '')
    a = agg.setdefault(key, [0.0, 0])
    a[0] += us; a[1] += 1
print('policy', {k: nat.get_policy(k) for k in nat.POLICY_KEYS}, 'precision', precision, 'mode', mode)
print('sum of per-call device times %.1f us over %d calls' % (tot, len(events)))
for key, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-46s %8.1f us %4d calls' % (key, us, n))
if os.environ.get('BREAKDOWN_LIST'):
    for name, kern, shape, e0, e1 in events:
        print('%-28s %-18s %-24s %7.1f' % (name, kern, shape, e0.elapsed_time(e1) * 1e3))
