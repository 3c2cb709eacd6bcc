"""Single-env steps/s of the LIVE reference (under the import stubs) vs the oracle port on BASELINE configs[0], in this
container (needs /root/reference).  The figures are recorded in bench.py (PORT_VS_LIVE) so that the reference arm's JSON
line states how the port it times relates to the reference's own Python step.

    python tools/port_vs_live.py
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def live(seconds=6.0, cfg="cfg1"):
    from oracle import ref_shim
    ref = ref_shim.import_reference()
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        if cfg == "cfg1":
            env = ref.quadrotor.QuadrotorEnv(dynamics_params="DefaultQuad", raw_control=False, ep_time=5, sim_freq=200.,
                                             sim_steps=2, sense_noise="default", obs_repr="xyz_vxyz_R_omega_act")
        else:
            env = ref.quadrotor.QuadrotorEnv(dynamics_params="DefaultQuad", raw_control=True, raw_control_zero_middle=True,
                                             ep_time=5, sim_freq=200., sim_steps=2, sense_noise=None, obs_repr="xyz_vxyz_R_omega")
        env.reset()
    a = np.zeros(4)
    for _ in range(200):
        env.step(a)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(100):
                _, _, done, _ = env.step(a)
                if done:
                    env.reset()
        n += 100
    return n / (time.perf_counter() - t0)


def port(seconds=6.0, cfg="cfg1"):
    import bench
    return bench._cpu_worker((seconds, 1, 1, cfg))


if __name__ == "__main__":
    for cfg in ("cfg1", "cfg2"):
        lv, pt = live(cfg=cfg), port(cfg=cfg)
        print(cfg, {"live_reference_steps_per_s": round(lv, 1), "oracle_port_steps_per_s": round(pt, 1), "cpu_count": os.cpu_count()})
