"""Multi-GPU plumbing: environment batches are sharded over ranks (one process per GPU) with
no communication inside a step; the only collective on the env path is the all-gather of the
per-environment rewards / losses.  Works with the `nccl` backend on GPUs and with
`gloo` on CPU tensors (used by the world_size-2 tests)."""
import torch
import torch.distributed as dist


def shard_bounds(num_envs, rank, world):
    """Contiguous, balanced [lo, hi) slice of the env axis owned by `rank` (first `num_envs %
    world` ranks hold one extra environment)."""
    base, extra = divmod(int(num_envs), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_env_values(local, num_envs=None, group=None):
    """All-gather a per-environment tensor [b_local, ...] into [num_envs, ...] on every rank.
    Shards may be ragged (see shard_bounds): they are padded to the largest shard for the
    collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if num_envs is None:
        n = torch.tensor([local.shape[0]], device=local.device)
        dist.all_reduce(n, group=group)
        num_envs = int(n.item())
    sizes = [shard_bounds(num_envs, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * bmax : r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    del rank
    return torch.cat(parts, dim=0)
