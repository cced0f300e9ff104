"""Host-side logic of the multi-GPU partitioning with a world-size-2 gloo group on CPU:
shard arithmetic, gather order for the sharded 8x TTA, all-reduce bookkeeping of data-parallel training."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_range_covers_everything_once():
    from deepcalcium.engine.dist import shard_range
    for n in (0, 1, 7, 8, 32, 256, 3000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert [shard_range(8, 8, r) for r in range(8)] == [(r, 1) for r in range(8)]
    assert [shard_range(8, 2, r) for r in range(2)] == [(0, 4), (4, 4)]
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from deepcalcium.engine.dist import Comm, shard_range
        from deepcalcium.utils.neurons import INVERTIBLE_2D_AUGMENTATIONS as TABLE
        comm = Comm()
        assert (comm.world, comm.rank) == (world, rank)
        # --- sharded TTA: each rank "predicts" its slice of the 8 transforms, root combines in order 0..7
        S = 16
        rng = np.random.default_rng(123)
        probs = rng.random((8, S, S)).astype(np.float32)           # what a 1-GPU run would have produced
        first, count = shard_range(8, world, rank)
        parts = comm.gather_to_root(torch.from_numpy(probs[first:first + count].copy()),
                                    [shard_range(8, world, r)[1] for r in range(world)])
        if rank == 0:
            allp = torch.cat(parts, 0).numpy()
            assert np.array_equal(allp, probs)
            mp_ = np.zeros((S, S))
            for k, (_, aug, inv) in enumerate(TABLE):
                mp_ += inv(allp[k:k + 1])[0] / len(TABLE)
            ref = np.zeros((S, S))
            for k, (_, aug, inv) in enumerate(TABLE):
                ref += inv(probs[k:k + 1])[0] / len(TABLE)
            assert np.array_equal(mp_, ref)
        else:
            assert parts is None
        # --- uneven shards (3 ranks would give 3,3,2; with world 2 use 5 items -> 3,2)
        f5, c5 = shard_range(5, world, rank)
        items = torch.arange(5, dtype=torch.float32)[f5:f5 + c5].reshape(c5, 1)
        got = comm.gather_to_root(items, [shard_range(5, world, r)[1] for r in range(world)])
        if rank == 0:
            assert torch.equal(torch.cat(got, 0).flatten(), torch.arange(5, dtype=torch.float32))
        # --- data-parallel bookkeeping: BN sums all-reduced, dgamma pre-scaled by 1/world, gradients summed
        x_all = np.random.default_rng(7).standard_normal((8, 5))
        fb, cb = shard_range(8, world, rank)
        local = torch.tensor(np.stack([x_all[fb:fb + cb].sum(0), (x_all[fb:fb + cb] ** 2).sum(0)]))
        comm.allreduce_sum(local)
        assert np.allclose(local[0].numpy(), x_all.sum(0)) and np.allclose(local[1].numpy(), (x_all ** 2).sum(0))
        dgamma_local = local[0] / world                             # what bn_bwd_apply writes with dgb_scale = 1/world
        comm.allreduce_sum(dgamma_local)                            # the flat-gradient all-reduce
        assert np.allclose(dgamma_local.numpy(), x_all.sum(0))
        # --- projection of one movie split into row bands: band projection + all_gather == full projection
        import oracle
        from deepcalcium.engine.dist import summarize_movie_sharded
        movie = (np.random.default_rng(5).random((13, 21, 16)) * 4096).astype(np.float32)      # 21 rows: uneven bands

        def cpu_project(band, floor):                               # stand-in for the CUDA kernel in this CPU test
            m, x = oracle.project_mean_max(band.numpy(), floor_max_at_zero=floor)
            return torch.from_numpy(m), torch.from_numpy(x)
        fr, nr = shard_range(21, world, rank)
        mean, mx = summarize_movie_sharded(torch.from_numpy(movie[:, fr:fr + nr].copy()), comm, 21, project_fn=cpu_project)
        omean, omx = oracle.project_mean_max(movie)
        assert np.array_equal(mean.numpy(), omean) and np.array_equal(mx.numpy(), omx)
        try:
            summarize_movie_sharded(torch.from_numpy(movie[:, :3].copy()), comm, 21, project_fn=cpu_project)
            raise AssertionError('a band of the wrong height must be rejected')
        except ValueError:
            pass
        t = torch.full((3,), float(rank + 1))
        comm.broadcast(t)
        assert torch.all(t == 1.0)
        q.put((rank, 'ok'))
    except Exception as ex:     # noqa: BLE001
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
