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


class EpisodeStatsReducer(object):
    """The path's only collective, kept off the step stream.

    The caller steps its envs on `step_stream` (typically replaying CUDA graphs) and calls `after_steps(step_count)`
    whenever it has enqueued more steps.  Each time the count crosses a multiple of `period` (= ep_len + 1 steps: the
    envs of a batch were reset together, so that is when their episodes end, quadrotor.py:986-987) the reducer

      1. snapshots the device-resident double[3] blocks of all its handles into one slot of a small ring with ONE
         copy kernel on the step stream (so the snapshot is consistent: it sits between two steps), and
      2. all-reduces (SUM) that slot on its own side stream, ordered after the snapshot by an event.

    The step stream never waits for the collective: neither its host enqueue nor inter-rank skew sits between two
    steps.  Without CUDA (gloo, CPU tests) the same cadence logic runs with a synchronous all-reduce of a clone.
    """

    SLOTS = 4

    def __init__(self, stats_tensors, period, group=None):
        import torch
        self.torch = torch
        self.tensors = list(stats_tensors)
        self.period = int(period)
        self.group = group
        self.next_at = self.period
        self.fired = 0
        self.cuda = self.tensors[0].is_cuda
        dev = self.tensors[0].device
        self.slots = [torch.zeros((len(self.tensors), 3), dtype=torch.float64, device=dev) for _ in range(self.SLOTS)]
        self.last = None
        if self.cuda:
            self.stream = torch.cuda.Stream(device=dev)
            self.free = [None] * self.SLOTS           # event: the collective that last read the slot has finished
            self.work = [None] * self.SLOTS

    def due(self, step_count):
        """True (once) for every call whose step_count has reached the next multiple of `period`."""
        if step_count < self.next_at:
            return False
        while self.next_at <= step_count:
            self.next_at += self.period
        return True

    def after_steps(self, step_count, step_stream=None):
        if self.due(step_count):
            self.reduce_now(step_stream)
            return True
        return False

    def reduce_now(self, step_stream=None):
        torch = self.torch
        import torch.distributed as dist
        have_pg = dist.is_available() and dist.is_initialized()
        k = self.fired % self.SLOTS
        self.fired += 1
        slot = self.slots[k]
        if not self.cuda:
            torch.stack([t.reshape(3) for t in self.tensors], out=slot)
            if have_pg:
                dist.all_reduce(slot, op=dist.ReduceOp.SUM, group=self.group)
            self.last = slot
            return
        step_stream = step_stream or torch.cuda.current_stream(slot.device)
        with torch.cuda.stream(step_stream):
            if self.free[k] is not None:
                step_stream.wait_event(self.free[k])           # finished long ago: SLOTS episodes back
            torch.stack(self.tensors, out=slot)                # one small kernel between two steps
            snap = torch.cuda.Event()
            snap.record(step_stream)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(snap)
            if have_pg:
                self.work[k] = dist.all_reduce(slot, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self.work[k].wait()                            # orders the side stream (not the host) after the collective
            done = torch.cuda.Event()
            done.record(self.stream)
            self.free[k] = done
        self.last = slot

    def result(self):
        """Job-wide {mean, std, episodes} of the most recent reduction (synchronises the side stream only)."""
        if self.last is None:
            return {"mean": None, "std": None, "episodes": 0.0}
        if self.cuda:
            self.stream.synchronize()
        s, s2, n = (float(x) for x in self.last.sum(0).tolist())
        if n <= 0:
            return {"mean": None, "std": None, "episodes": 0.0}                # (None, not NaN: the result goes into JSON lines)
        mean = s / n
        return {"mean": mean, "std": max(s2 / n - mean * mean, 0.0) ** 0.5, "episodes": n}
