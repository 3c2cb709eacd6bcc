"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path - contiguous env-index sharding, the
per-episode statistics all-reduce (the path's only collective) and invariance of the results to the rank count.
The per-rank compute here is the NumPy oracle with injected noise (the checker), standing in for the CUDA step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_art_b200.stats import all_reduce_stats, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, T, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import quadsim_oracle as qo
    from gym_art_b200.quadrotor import quad_models
    b, e = shard_range(n_total, rank, world)
    rng = np.random.RandomState(0)                       # every rank draws the job-wide tape and takes its block
    env = qo.OracleVecEnv(quad_models.defaultquad_params(), e - b, controller=qo.CTRL_RAW, ep_time=0.1, auto_reset=True)
    rz = rng.uniform(-1, 1, (n_total, qo.RESET_Z))
    env.reset(reset_z=rz[b:e])
    obs = None
    for t in range(T):
        a = rng.uniform(-1, 1, (n_total, 4))
        ou = rng.standard_normal((n_total, 2, 4))
        rz = rng.uniform(-1, 1, (n_total, qo.RESET_Z))
        obs, rew, done = env.step(a[b:e], ou_z=ou[b:e], reset_z=rz[b:e])
    stats = all_reduce_stats(torch.as_tensor(env.stats))
    gathered = [None] * world
    dist.all_gather_object(gathered, (b, e, obs))
    if rank == 0:
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        np.savez(os.path.join(out_dir, "w%d.npz" % world), obs=full, mean=stats["mean"], std=stats["std"], n=stats["episodes"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_job_equals_one_rank_job(tmp_path):
    n_total, T = 37, 45              # odd size: the shards differ by one env; 11-step episodes, four finish per env
    for world in (1, 2):
        mp.spawn(_worker, args=(world, _free_port(), n_total, T, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "w1.npz"), np.load(tmp_path / "w2.npz")
    assert np.array_equal(a["obs"], b["obs"])                                  # trajectories do not depend on the rank count
    assert float(a["n"]) == float(b["n"]) == 4 * n_total
    assert abs(float(a["mean"]) - float(b["mean"])) < 1e-12 and abs(float(a["std"]) - float(b["std"])) < 1e-9


def test_shard_range_partitions():
    for n in (1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_all_reduce_stats_without_process_group():
    s = all_reduce_stats(np.array([10.0, 60.0, 4.0]))
    assert s["episodes"] == 4 and abs(s["mean"] - 2.5) < 1e-12 and abs(s["std"] - (15.0 - 6.25) ** 0.5) < 1e-12
    assert all_reduce_stats(np.zeros(3))["episodes"] == 0
