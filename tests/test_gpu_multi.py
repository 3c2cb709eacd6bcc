"""GPU: one process per rank with env-index sharding and the per-episode statistics all-reduce - over NCCL on two devices (needs
2+ GPUs, skipped on a single-GPU box), and over gloo with both ranks sharing cuda:0 (runs anywhere); two handles on two devices
inside one process (needs 2+ GPUs)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, T, out, share_gpu=False):
    import torch
    import torch.distributed as dist
    local = 0 if share_gpu else rank
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local))
    torch.cuda.set_device(local)
    if share_gpu:       # both ranks on cuda:0 (NCCL refuses two ranks on one device): gloo carries the 24-byte all-reduce
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gym_art_b200.sharding import episode_stats, make_sharded_env
    env = make_sharded_env(n_total, dynamics_params="Crazyflie", controller="mellinger", sense_noise="default",
                           init_random_state=True, ep_time=0.1, seed=21)
    env.reset()
    for _ in range(T):
        o, r, d = env.step(None)
    st = episode_stats(env)
    b, e = env.shard
    gathered = [None] * world
    dist.all_gather_object(gathered, (b, o.cpu().numpy()))
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])])
        np.savez(out, obs=full, mean=st["mean"], std=st["std"], n=st["episodes"])
    dist.barrier()
    dist.destroy_process_group()


def _two_rank_job_equals_single_gpu_job(tmp_path, share_gpu):
    import torch
    import torch.multiprocessing as mp
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.stats import all_reduce_stats
    n_total, T = 100000 + 37, 45
    out = str(tmp_path / "w2.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_total, T, out, share_gpu), nprocs=2, join=True)
    env = QuadVecEnv(n_total, dynamics_params="Crazyflie", controller="mellinger", sense_noise="default",
                     init_random_state=True, ep_time=0.1, seed=21)
    env.reset()
    for _ in range(T):
        o, r, d = env.step(None)
    one = all_reduce_stats(env.stats_tensor())
    two = np.load(out)
    assert np.array_equal(two["obs"], o.cpu().numpy())          # Philox is keyed by the global env id
    assert float(two["n"]) == one["episodes"] == 4 * n_total
    assert abs(float(two["mean"]) - one["mean"]) < 1e-9 and abs(float(two["std"]) - one["std"]) < 1e-7


@needs2
def test_two_rank_nccl_job_equals_single_gpu_job(tmp_path):
    _two_rank_job_equals_single_gpu_job(tmp_path, share_gpu=False)


def test_two_rank_job_sharing_one_gpu_equals_single_gpu_job(tmp_path):
    """The same two-process sharded job on a single-GPU box: both ranks step their env-index block on cuda:0 (two processes,
    two handles, kernels of both interleaving on one device) and all-reduce their episode statistics over gloo - trajectories
    and job statistics equal the one-process job bit for bit.  (Runs where the NCCL variant above has to be skipped.)"""
    _two_rank_job_equals_single_gpu_job(tmp_path, share_gpu=True)


@needs2
def test_two_handles_on_two_devices_in_one_process():
    import torch
    from gym_art_b200 import QuadVecEnv
    n = 150000                                                    # pipelined instantiation on both devices
    envs = [QuadVecEnv(n, device="cuda:%d" % k, seed=3) for k in range(2)]
    outs = []
    for k, e in enumerate(envs):
        with torch.cuda.device(k):
            e.reset()
            a = torch.zeros((n, 4), device="cuda:%d" % k)
            for _ in range(5):
                o, r, d = e.step(a)
            torch.cuda.synchronize(k)
            e.check_finite()
            outs.append(o.cpu())
    assert torch.equal(outs[0], outs[1])
