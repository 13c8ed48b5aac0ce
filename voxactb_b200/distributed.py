"""Batch sharding across GPUs: the path is embarrassingly parallel over the batch,
so inference needs no data-path collective -- every rank voxelises and evaluates its own contiguous slice
with replicated weights.  ``torch.distributed`` is used only for rendezvous, barriers and the max-over-ranks
of device timings (NCCL on GPUs, gloo in the CPU tests).  Mirrors how the reference gives every DDP rank its
own strided shard of the replay (YARR task_uniform_replay_buffer.py:103-108)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous slice [begin, end) of n samples owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('bad rank/world %r/%r' % (rank, world))
    base, rem = divmod(int(n), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_observation(obs, rank, world):
    """Slice every batched tensor (or list of tensors) of an observation dict along dim 0."""
    n = obs['proprio'].shape[0]
    b, e = shard_range(n, rank, world)

    def cut(v):
        if isinstance(v, (list, tuple)):
            return [cut(t) for t in v]
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
            return v[b:e].contiguous()
        return v
    return {k: cut(v) for k, v in obs.items()}


def max_over_ranks(value, device=None):
    """Whole-job time of a step = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(t, n_total):
    """All-gather per-rank result rows (e.g. selected actions) into batch order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(e - b for b, e in sizes)
    pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)], 0)
