"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck): ragged sizes, multi-tile
persistent path, one-wave path, swarm, per-env models, reset, dynamics_step, host pipeline."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_art_b200 import QuadVecEnv  # noqa: E402
from gym_art_b200.quadrotor import quad_models  # noqa: E402
from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals  # noqa: E402

big = 2 * 148 * 4 * 128 + 3 * 128 + 17      # enough tiles for the persistent pipeline, ragged tail
for n, kw in ((big, dict(controller="raw")), (1000 + 5, dict(controller="omega_thrust", sense_noise="default", init_random_state=True)),
              (big - 17 + 8, dict(controller="raw", n_agents=8, formation=ring_goals(8, 0.3), collision_k=2.0)),
              (517, dict(controller="mellinger", obs_repr="xyz_vxyz_R_omega_acc_act", sense_noise="default")),
              (300, dict(controller="raw", dtype="float64", info=True)),
              (260, dict(controller="raw", per_env_dynamics=2)),
              # the feature-family kernels (qs::Feat), pipelined and one-wave, ragged tails
              (big + 5, dict(controller="raw", dynamics_change={"motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424},
                                                                "damp": {"vel": 0.001, "omega_quadratic": 0.015}})),
              (big + 9, dict(controller="raw", sense_noise=dict(pos_unif_range=0.01, quat_norm_std=0.01, gyro_norm_std=0.1),
                             obs_repr="xyz_vxyz_R_omega_acc_act")),
              (big + 3, dict(controller="vertical", dim_mode="1D", init_random_state=True)),
              (700 + 3, dict(controller="vertplane", dim_mode="2D", obs_repr="xyzr_vxyzr_quat_omega_h", sense_noise="default")),
              (big + 1, dict(controller="raw", info=True, obs_repr="xyz_vxyz_R_omega_t2w_t2t", resample_goal=True)),
              (big - 17 + 4, dict(controller="raw", n_agents=4, formation=ring_goals(4, 0.3), collision_k=2.0, per_env_dynamics=2))):
    env = QuadVecEnv(n, dynamics_params="DefaultQuad", ep_time=0.05, seed=3, **kw)
    if kw.get("per_env_dynamics"):
        env.set_models([quad_models.crazyflie_params(), quad_models.defaultquad_params()])
    env.reset()
    a = torch.rand((n, env.act_dim), device="cuda", dtype=env.tdtype) * 2 - 1
    for t in range(14):
        env.step(a)
    env.reset(torch.arange(n, device="cuda") % 3 == 0)
    env.dynamics_step(a, controller=env.controller_name, n_substeps=3)
    env.affine_fg(torch.rand((n % 1000 + 3, 21), device="cuda", dtype=env.tdtype), 0.005)
    if env.tdtype == torch.float32 and n < 2000:
        env.step_k(3, a.expand(3, n, env.act_dim).contiguous())
        ah, oh = a.cpu().pin_memory(), torch.empty((n, env.obs_dim)).pin_memory()
        rh, dh = torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
        env.step_host(ah, oh, rh, dh)
        env.step_host(a.cpu(), torch.empty((n, env.obs_dim)), torch.empty(n), torch.empty(n, dtype=torch.uint8))
    torch.cuda.synchronize()
    env.check_finite()
    print("ok", n, env.kernel_variant, flush=True)
    env.close()
