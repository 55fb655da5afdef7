"""Multi-GPU plumbing: the vehicle batch is sharded in contiguous ranges over the ranks (one process per GPU) with no collective on
the hot path; the only exchange is the final gather of controls (3 f64 / vehicle) and solver statistics over torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np


def shard_range(total, world, rank):
    """Contiguous range [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_arrays(arrays, world, rank):
    lo, hi = shard_range(len(arrays[0]), world, rank)
    return [a[lo:hi] for a in arrays]


def gather_batch(local, total, dist, device=None):
    """All-gather per-vehicle rows (local: (b_local, k) float64 / int32 numpy array or torch tensor) into the global (total, k) array
    in shard order.  Ragged shards are padded to the largest shard for the collective."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    squeeze = t.dim() == 1
    if squeeze:
        t = t[:, None]
    sizes = [shard_range(total, world, r) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((bmax, t.shape[1]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    full = torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
    return full[:, 0] if squeeze else full
