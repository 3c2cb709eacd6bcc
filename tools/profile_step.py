"""Small driver for ncu: steps one QuadVecEnv a few times with plain launches (no CUDA graph).

    python tools/profile_step.py --envs 4194304 --steps 8 --config cfg2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gym_art_b200 import QuadVecEnv
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4 * 1024 * 1024)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg5", "cfg4", "drag", "noise", "vertical", "info", "t2w", "catchall", "pem_t2w"])
    ap.add_argument("--copy", action="store_true")
    a = ap.parse_args()
    kw = dict(device="cuda:0", ep_time=5, seed=1)
    if a.config == "cfg2":
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="raw", **kw)
    elif a.config == "cfg3":
        env = QuadVecEnv(a.envs, dynamics_params="Crazyflie", controller="omega_thrust", sense_noise="default",
                         init_random_state=True, **kw)
    elif a.config == "cfg5":
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="mellinger", **kw)
    elif a.config == "drag":      # one feature group each (qs::Feat): the configurations of tools/generic_perf.py
        env = QuadVecEnv(a.envs, dynamics_params="Crazyflie", controller="raw", dynamics_change={
            "motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424}, "damp": {"vel": 0.001, "omega_quadratic": 0.015}}, **kw)
    elif a.config == "noise":
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="raw",
                         sense_noise=dict(pos_unif_range=0.01, quat_norm_std=0.01, gyro_norm_std=0.1), **kw)
    elif a.config == "vertical":
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="vertical", dim_mode="1D", **kw)
    elif a.config == "t2w":
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="raw", obs_repr="xyz_vxyz_R_omega_t2w_t2t", sense_noise="default", **kw)
    elif a.config == "catchall":
        env = QuadVecEnv(a.envs, dynamics_params="Crazyflie", controller="raw", dynamics_change={
            "motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424}, "damp": {"vel": 0.001, "omega_quadratic": 0.015}},
            sense_noise=dict(pos_unif_range=0.01, quat_norm_std=0.01, gyro_norm_std=0.1), **kw)
    elif a.config == "pem_t2w":
        import numpy as np
        np.random.seed(7)
        env = QuadVecEnv(a.envs, dynamics_params="RandomQuad", per_env_dynamics=256, controller="raw", obs_repr="xyz_vxyz_R_omega_t2w",
                         sense_noise="default", **kw)
    elif a.config == "info":
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="raw", info=True, **kw)
    else:
        from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
        env = QuadVecEnv(a.envs, dynamics_params="DefaultQuad", controller="raw", n_agents=8, formation=ring_goals(8),
                         **kw)
    # same-run yardstick: STREAM-style copy bandwidth of this very GPU (method of MEASURED_PEAKS.json)
    if a.copy:
        x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda:0")
        y = torch.empty_like(x)
        best = 0.0
        for _ in range(10):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(); y.copy_(x); c1.record(); torch.cuda.synchronize()
            best = max(best, 2 * x.numel() * 2 / (c0.elapsed_time(c1) * 1e-3) / 1e9)
        print("copy bandwidth on this GPU: %.0f GB/s" % best)
        del x, y
    env.reset()
    act = torch.rand((a.envs, env.act_dim), device="cuda:0") * 2 - 1
    if a.config == "cfg3":
        act[:, 0] *= 0.2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        env.step(act)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        env.step(act)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print("variant %s  %.1f us/step  %.3e env-steps/s  %.0f GB/s algorithmic (349 B/env-step)" % (
        env.kernel_variant, ms * 1e3, a.envs / (ms * 1e-3), a.envs * 349 / (ms * 1e-3) / 1e9))
    env.check_finite()


if __name__ == "__main__":
    main()
