"""Multi-GPU plumbing: instances are independent, so an ensemble shards by contiguous instance-id range with
NO collective inside the step loop.  The only exchange is one gather of per-instance statistics
(and optionally final states) at the end, over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(total, rank, world):
    """contiguous, balanced id range [lo, hi) of `rank`; the first `total % world` ranks get one extra"""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total, world):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_per_instance(local, total, group=None):
    """all-gather a per-instance tensor (last dim = this rank's instances, in id order) into the global
    id order.  `local`: torch tensor [..., n_local].  Returns [..., total] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    pad = max(sizes)
    lead = tuple(local.shape[:-1])
    buf = torch.zeros(lead + (pad,), dtype=local.dtype, device=local.device)
    buf[..., : local.shape[-1]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf.contiguous(), group=group)
    return torch.cat([o[..., :s] for o, s in zip(out, sizes)], dim=-1)


def local_ids(total, rank, world):
    lo, hi = shard_range(total, rank, world)
    return np.arange(lo, hi, dtype=np.uint64)
