"""Multi-GPU sharding of the env batch: one process per GPU (torch.distributed, NCCL on B200 / gloo in CPU tests).

The step has no cross-env term, so ranks never exchange data on the step path: rank r owns the
contiguous slice [offset, offset+count) of the global env index space and keys its RNG streams by GLOBAL env index
(`env_index_offset`), which makes results independent of the world size.  Collectives exist only for consumers that want
global views: scalar statistics (all-reduce) or, optionally, the gathered observation batch (all-gather).
"""
import os


def shard_envs(total_envs, rank, world_size):
    """Contiguous, balanced partition: returns (count, offset) for `rank`."""
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return count, offset


def rank_world():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_sharded(env_id, total_envs, **kwargs):
    """`make` for this rank's shard of a global batch of `total_envs` environments (device = LOCAL_RANK by default)."""
    from .envs import make

    rank, world = rank_world()
    count, offset = shard_envs(total_envs, rank, world)
    kwargs.setdefault("device", int(os.environ.get("LOCAL_RANK", "0")))
    return make(env_id, num_envs=count, env_index_offset=offset, **kwargs)


def global_stats(reward, terminated):
    """(mean reward, number of terminated envs) over ALL ranks: one all-reduce of two scalars."""
    import torch
    import torch.distributed as dist

    t = torch.stack([reward.double().sum(), terminated.double().sum(), torch.tensor(float(reward.numel()), device=reward.device, dtype=torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return float(t[0] / t[2]), int(t[1].item())


def all_gather_batch(*tensors):
    """Optional single all-gather of per-rank [n_local, ...] tensors into global [N, ...] tensors (same n_local on every
    rank).  For PMSM at N=2^20 this moves ~72 MB per step — several times the step itself; data-parallel
    learners should keep observations rank-local and use global_stats instead."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensors
    out = []
    for t in tensors:
        g = torch.empty((dist.get_world_size() * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t.contiguous())
        out.append(g)
    return tuple(out)
