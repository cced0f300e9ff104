"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing;
NCCL over NVLink on the GPU box, gloo in the CPU tests).  New functionality - the reference is
single-device - specified in SURVEY.md section 8e:

  * independent units (movies, images, windows): ``shard_range`` splits them, no collective;
  * one image's 8 TTA transforms over n ranks (BASELINE config C4): transform k -> rank owning
    ``shard_range(8, n, r)``; probability maps are gathered to rank 0 and combined in the fixed
    order 0..7, so the mask is bit-identical to the 1-GPU result;
  * data-parallel training (C5): rank r trains on crops ``shard_range(B, n, r)``; BatchNorm batch
    statistics, the loss sums and the gradients are all-reduced (sum) so every rank applies the
    single-device update of the whole global batch.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, world, rank):
    """Contiguous balanced split: returns (first, count) of rank's share of n_items."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world %d' % (rank, world))
    base, rem = divmod(n_items, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class Comm(object):
    """The three collectives the path needs.  A Comm of world 1 is a no-op."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    def allreduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def all_gather_rows(self, t, counts):
        """every rank receives every rank's [K, rows_r, W] tensor (rows_r = counts[r], may differ by one)"""
        pad = max(counts)
        buf = t
        if t.shape[1] < pad:
            buf = torch.zeros((t.shape[0], pad) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
            buf[:, :t.shape[1]] = t
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf.contiguous(), group=self.group)
        return [o[:, :c] for o, c in zip(out, counts)]

    def gather_to_root(self, t, counts=None):
        """Rank 0 receives every rank's tensor (in rank order) and returns them as a list; other
        ranks return None.  ``counts``: leading-dimension size per rank when shards are uneven."""
        if self.world == 1:
            return [t]
        if counts is None:
            counts = [t.shape[0]] * self.world
        # all_gather keeps the call symmetric (gather is not implemented by every backend/version)
        pad = max(counts)
        buf = t
        if t.shape[0] < pad:
            buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            buf[:t.shape[0]] = t
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf.contiguous(), group=self.group)
        if self.rank != 0:
            return None
        return [o[:c] for o, c in zip(out, counts)]


def predict_tta_sharded(engine, summ_dev, comm, window=512, threshold=0.5):
    """One summary image, its 8 TTA transforms sharded over comm.world ranks (1, 2, 4 or 8).
    Returns (mask, act) device tensors on rank 0, (None, None) elsewhere."""
    from . import ops
    world, rank = comm.world, comm.rank
    if world == 1:
        return engine.predict_tta(summ_dev, window=window, augmentation=True, threshold=threshold)
    first, count = shard_range(8, world, rank)
    probs = engine.predict_tta(summ_dev, window=window, augmentation=True, threshold=threshold,
                               transforms=(first, count))
    parts = comm.gather_to_root(probs, [shard_range(8, world, r)[1] for r in range(world)])
    if rank != 0:
        return None, None
    hs, ws = summ_dev.shape
    allp = torch.cat(parts, dim=0).contiguous()
    mask = torch.empty(hs, ws, dtype=torch.uint8, device=summ_dev.device)
    act = torch.empty(hs, ws, dtype=torch.float64, device=summ_dev.device)
    ops.tta_combine(allp, window, hs, ws, threshold, 8, act, mask)
    return mask, act


def summarize_movie_sharded(band, comm, H, floor_max_at_zero=False, project_fn=None):
    """Projection of ONE movie split into row bands over the ranks (reference loop: datasets/nf.py:126-130).
    The reduction is per pixel, so rank r projects its band ``movie[:, first:first+rows, :]`` with
    ``(first, rows) = shard_range(H, world, r)`` without any exchange; the two [rows, W] float32 maps are then
    all-gathered (2 x 4 x H x W / world bytes per rank) so every rank holds the full (mean, max) images.
    ``band``: [T, rows, W] tensor on this rank's device.  ``project_fn(band, floor_max_at_zero) -> (mean, max)`` defaults
    to the CUDA kernel (deepcalcium.datasets.nf.summarize_movie_device); the CPU tests of the host logic inject one."""
    if project_fn is None:
        from ..datasets.nf import summarize_movie_device
        project_fn = summarize_movie_device
    world, rank = comm.world, comm.rank
    first, rows = shard_range(H, world, rank)
    if band.dim() != 3 or band.shape[1] != rows:
        raise ValueError('rank %d of %d holds rows [%d, %d) of %d: expected a [T, %d, W] band, got %s'
                         % (rank, world, first, first + rows, H, rows, tuple(band.shape)))
    mean, mx = project_fn(band, floor_max_at_zero)
    if world == 1:
        return mean, mx
    both = torch.stack([mean, mx], dim=0)                         # [2, rows, W]
    counts = [shard_range(H, world, r)[1] for r in range(world)]
    parts = comm.all_gather_rows(both, counts)                    # list over ranks of [2, rows_r, W]
    full = torch.cat(parts, dim=1)
    return full[0].contiguous(), full[1].contiguous()


def sync_parameters(engine, comm):
    """All replicas start from rank 0's weights and optimiser state."""
    for t in (engine.params, engine.nontrain, engine.adam_m, engine.adam_v, engine.step_state):
        comm.broadcast(t)
    engine._weights_dirty = True
