"""Step time of configurations outside the fast subset (generic kernel) and of the non-plain fast variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_art_b200 import QuadVecEnv
dev = torch.device("cuda", 0)
peak, _ = bench.load_peaks()
n = 4 << 20
DRAG = {"motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424}, "damp": {"vel": 0.001, "omega_quadratic": 0.015}}
cases = [
    ("cli-default: mellinger + default noise + obs ..._act (S=36,O=22: 397 B)", dict(controller="mellinger", sense_noise="default", obs_repr="xyz_vxyz_R_omega_act"), 397 - 16, "float32"),
    ("raw + obs xyz_vxyz_R_omega_acc_act + noise (S=36,O=25: 409 B)", dict(controller="raw", sense_noise="default", obs_repr="xyz_vxyz_R_omega_acc_act"), 409, "float32"),
    ("raw, sim_steps=4", dict(controller="raw", sim_steps=4), 349, "float32"),
    ("generic: rotor drag + linearity + damping", dict(controller="raw", dynamics_change=DRAG), 349, "float32"),
    ("generic: full sensor noise (uniform, quaternion, gyro bias)", dict(controller="raw", sense_noise=dict(pos_unif_range=0.01, quat_norm_std=0.01, gyro_norm_std=0.1)), 349 + 24, "float32"),
    ("generic: vertical controller (1D)", dict(controller="vertical", dim_mode="1D"), 349 - 12, "float32"),
    ("misc: info buffer bound (+192 B)", dict(controller="raw", info=True), 349 + 192, "float32"),
    ("misc: quaternion observation xyz_vxyz_quat_omega (O=13: 329 B)", dict(controller="raw", obs_repr="xyz_vxyz_quat_omega"), 329, "float32"),
    ("misc: xyz_vxyz_R_omega_t2w_t2t + default noise (O=20: 357 B)", dict(controller="raw", obs_repr="xyz_vxyz_R_omega_t2w_t2t", sense_noise="default"), 357, "float32"),
    ("pair family: drag + full sensor noise", dict(controller="raw", dynamics_change=DRAG, sense_noise=dict(pos_unif_range=0.01, quat_norm_std=0.01, gyro_norm_std=0.1)), 349 + 24, "float32"),
    ("catch-all: quaternion observation + full sensor noise (O=13)", dict(controller="raw", obs_repr="xyz_vxyz_quat_omega", sense_noise=dict(pos_unif_range=0.01, quat_norm_std=0.01, gyro_norm_std=0.1)), 329 + 24, "float32"),
    ("per-env models (RandomQuad x256), raw, plain shape (+112 B model rows)", dict(controller="raw", dynamics_params="RandomQuad", per_env_dynamics=256), 461, "float32"),
    ("per-env models, perturbed motors (RelativeSampler: linearity < 1) (+144 B)", dict(controller="raw", per_env_dynamics=256, dyn_sampler_1={"class": "RelativeSampler", "noise_ratio": 0.1}), 349 + 144, "float32"),
    ("per-env models + t2w observation + default noise (O=19, +128 B)", dict(controller="raw", dynamics_params="RandomQuad", per_env_dynamics=256, obs_repr="xyz_vxyz_R_omega_t2w", sense_noise="default"), 353 + 128, "float32"),
    ("per-env models in swarms of 8                  ", dict(controller="raw", dynamics_params="RandomQuad", per_env_dynamics=256, n_agents=8, formation="ring8"), 461 + 16, "float32"),
    ("generic fp64 (verification build), raw", dict(controller="raw"), 2 * 349 - 1, "float64"),
]
for name, kw, nbytes, dtype in cases:
    m = n if dtype == "float32" else n // 4
    kw = dict(kw)
    kw.setdefault("dynamics_params", "Crazyflie" if "drag" in name else "DefaultQuad")   # (the drag model is the CrazyFlie's)
    if kw.get("formation") == "ring8":
        from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
        kw["formation"] = ring_goals(8)
    import numpy as np
    np.random.seed(7)
    env = QuadVecEnv(m, ep_time=5, seed=1, dtype=dtype, **kw)
    env.reset()
    a = None if env.controller_name == "mellinger" else torch.rand((m, env.act_dim), device=dev, dtype=env.tdtype) * 2 - 1
    bench.RING_PERIOD[0] = 1
    ms, _ = bench.time_graph(torch, lambda i: env.step(a), 20, 4, dev, chunk=20)
    us = ms[0] * 1e3 / 20
    print("%-78s %-22s %8.1f us  %.3e env-steps/s  frac %.2f" % (name, env.kernel_variant, us, m / us * 1e6, m * nbytes / us * 1e-3 / peak), flush=True)
    env.close()
