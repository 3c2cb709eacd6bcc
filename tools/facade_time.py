import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gym_art_b200.quadrotor.quadrotor import QuadrotorEnv
for kw in (dict(raw_control=True), dict(raw_control=False, sense_noise="default", obs_repr="xyz_vxyz_R_omega_act")):
    env = QuadrotorEnv(dynamics_params="DefaultQuad", ep_time=5, seed=1, **kw)
    a = np.zeros(4)
    for _ in range(50):
        env.step(a)
    t0 = time.perf_counter()
    n = 0
    for _ in range(1000):
        o, r, d, info = env.step(a)
        n += 1
        if d:
            env.reset()
    dt = time.perf_counter() - t0
    print(kw, "%.1f us/step  %.0f steps/s" % (dt / n * 1e6, n / dt))
    env.close()
