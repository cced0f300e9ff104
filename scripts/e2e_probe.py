import os, sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/deep-calcium_b200')
os.environ.setdefault('DEEP_CALCIUM_HOME', '/tmp/deep-calcium-home')
import numpy as np, torch
if 'WORLD_SIZE' in os.environ:      # under torchrun: one rank per GPU, NCCL initialised like bench.py does
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
    dist.barrier()
from deepcalcium.engine.graph import GraphSpec, he_normal_weights
from deepcalcium.engine.unet_engine import UNetEngine
from deepcalcium.models.neurons import UNet2DSummary
from deepcalcium.models.neurons.unet_2d_summary import UNetModel
spec = GraphSpec(32)
eng = UNetEngine(spec, precision='fp16')
eng.set_weights_dict(he_normal_weights(spec, seed=7535))
rng = np.random.default_rng(1)
host_imgs = {('img%d' % i): rng.standard_normal((512, 512)).astype(np.float32) for i in range(4)}
model = UNetModel.__new__(UNetModel)
model.window_shape, model.spec, model.engine = (512, 512), spec, eng
api = UNet2DSummary(cpdir='/tmp/deep-calcium-bench-cp-x', dataset_name_func=lambda p: p, series_summary_func=lambda p: host_imgs[p])
paths = [('img%d' % (i % 4)) for i in range(20)]
api.predict(paths[:6], model, augmentation=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
api.predict(paths, model, augmentation=True)
torch.cuda.synchronize()
print('threads', torch.get_num_threads(), 'OMP', os.environ.get('OMP_NUM_THREADS'), 'e2e img/s', 20 / (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable()
api.predict(paths, model, augmentation=True)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
