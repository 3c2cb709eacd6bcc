"""Env-index sharding across the GPUs of one box: one process per GPU, no collective on the step path.

The envs of a job are independent, so rank r owns the contiguous env-index block
shard_range(n_total, r, world) on cuda:LOCAL_RANK.  Philox streams are keyed by the GLOBAL env id
(QsConfig.env_id_offset), which makes every env's trajectory independent of the GPU count.  The only exchange is
the per-episode all-reduce of {sum of returns, sum of squared returns, episodes} (24 bytes; NCCL on GPUs, gloo in
the CPU tests) in gym_art_b200.stats.
"""
import os

from .stats import all_reduce_stats, shard_range


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def make_sharded_env(n_total, rank=None, world_size=None, local_rank=None, **env_kwargs):
    """This rank's QuadVecEnv over its block of the job's `n_total` envs (swarms are never split)."""
    from .vec_env import QuadVecEnv
    r, w, lr = rank_world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    local_rank = lr if local_rank is None else local_rank
    m = int(env_kwargs.get("n_agents", 1))
    if n_total % m:
        raise ValueError("n_total must be a multiple of n_agents")
    b, e = shard_range(n_total // m, rank, world_size)
    env = QuadVecEnv((e - b) * m, env_id_offset=b * m, device="cuda:%d" % local_rank, **env_kwargs)
    env.shard = (b * m, e * m)
    return env


def episode_stats(env, group=None, clear=False):
    """Job-wide finished-episode statistics: all-reduce(SUM) of the device-resident double[3] of every rank."""
    import torch
    t = env.stats_tensor().clone()
    if clear:
        env.stats(clear=True)
    out = all_reduce_stats(t, group=group)
    torch.cuda.current_stream(env.device).synchronize()
    return out
