"""Episode-return statistics across GPUs: the only collective on this path.

Each rank's kernel accumulates {sum of returns, sum of squared returns, finished episodes} into a device-resident
double[3]; once per episode boundary the ranks all-reduce(SUM) those 24 bytes over NCCL (gloo in the CPU tests).
"""
import numpy as np


def wrap_device_doubles(ptr, n, device):
    """torch float64 view of `n` doubles at device address `ptr` (memory owned by libquadsim)."""
    import torch

    class _Holder(object):
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    with torch.cuda.device(device):
        return torch.as_tensor(h, device=device)


def all_reduce_stats(stats, group=None):
    """stats: tensor [3] (device or CPU) -> dict(mean, std, episodes) over all ranks.  No-op without a process group."""
    import torch
    import torch.distributed as dist
    t = stats.clone() if isinstance(stats, torch.Tensor) else torch.as_tensor(np.asarray(stats, dtype=np.float64))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    s, s2, n = (float(x) for x in t.tolist())
    if n <= 0:
        return {"mean": float("nan"), "std": float("nan"), "episodes": 0.0}
    mean = s / n
    var = max(s2 / n - mean * mean, 0.0)
    return {"mean": mean, "std": var ** 0.5, "episodes": n}


def shard_range(n_total, rank, world_size):
    """Contiguous env-index block of `rank`: sizes differ by at most one, union = [0, n_total)."""
    base, rem = divmod(int(n_total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
