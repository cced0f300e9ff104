"""Per-layer device times of one 8-image 512x512 forward (eager, CUDA events) under the current env."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'deep-calcium_b200'))
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np, torch
import bench
from deepcalcium.engine.graph import GraphSpec, he_normal_weights
from deepcalcium.engine.unet_engine import UNetEngine
spec = GraphSpec(32)
eng = UNetEngine(spec, precision='bf16', use_graphs=False)
eng.set_weights_dict(he_normal_weights(spec, seed=7535))
s = torch.from_numpy(np.random.default_rng(1).standard_normal((512, 512)).astype(np.float32)).cuda()
for _ in range(3):
    eng.predict_tta(s)
sess = eng._session(8, 512, 512, False)
rows, tot_f, tot_ms = bench.per_layer_profile(eng, sess, spec, 8, 512, 512)
print('ENV', {k: v for k, v in os.environ.items() if k.startswith('DCB_')}, 'conv ms %.4f  TFLOP/s %.1f' % (tot_ms, tot_f / tot_ms / 1e9))
print(' '.join('%s=%.3f' % (r['op'].replace('conv3x3 ', 'c').replace('convT2x2 ', 'T').replace(' ', '_'), r['ms']) for r in rows))
