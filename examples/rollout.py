#!/usr/bin/env python
"""Rollouts on the B200 path, in the spirit of the reference's `./quadrotor.py -m mellinger -q DefaultQuad`
(test_rollout / benchmark, gym_art/quadrotor/quadrotor.py:1187-1432), minus the rendering.

    python examples/rollout.py                        # one env, reference-shaped API (facade), Mellinger controller
    python examples/rollout.py -n 1048576 -m raw      # a million envs, random raw actions, CUDA-graph stepped
    python examples/rollout.py -n 65536 -q RandomQuad --per-env-dynamics 256
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-m", "--mode", default="mellinger", choices=["mellinger", "raw", "omega_thrust"])
    ap.add_argument("-q", "--quad", default="DefaultQuad", help="DefaultQuad | Crazyflie | MediumQuad | RandomQuad")
    ap.add_argument("-n", "--num-envs", type=int, default=1)
    ap.add_argument("-e", "--episodes", type=int, default=2)
    ap.add_argument("-t", "--ep-time", type=float, default=5.0)
    ap.add_argument("-sn", "--sense-noise", action="store_true")
    ap.add_argument("--per-env-dynamics", type=int, default=None)
    a = ap.parse_args()
    noise = "default" if a.sense_noise else None

    if a.num_envs == 1:
        import numpy as np
        from gym_art_b200.quadrotor.quadrotor import QuadrotorEnv
        env = QuadrotorEnv(dynamics_params=a.quad, raw_control=a.mode != "mellinger", ep_time=a.ep_time, sense_noise=noise)
        for ep in range(a.episodes):
            s, ret, steps, t0 = env.reset(), 0.0, 0, time.perf_counter()
            while True:
                act = np.zeros(4) if a.mode == "mellinger" else env.action_space.sample()
                s, r, done, info = env.step(act)
                ret += r
                steps += 1
                if done:
                    break
            dt = time.perf_counter() - t0
            print("episode %d: %d steps, return %.4f, final distance to goal %.4f m, %.0f steps/s" % (
                ep, steps, ret, float(np.linalg.norm(env.goal - env.dynamics.pos)), steps / dt))
        env.close()
        return

    import torch
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.stats import all_reduce_stats
    env = QuadVecEnv(a.num_envs, dynamics_params=a.quad, controller=a.mode, ep_time=a.ep_time, sense_noise=noise,
                     per_env_dynamics=a.per_env_dynamics)
    env.reset()
    steps = (env.ep_len + 1) * a.episodes
    act = None if a.mode == "mellinger" else torch.zeros((a.num_envs, env.act_dim), device=env.device)
    chunk = 50
    g = env.capture_steps(chunk, act)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done_steps = 1
    while done_steps < steps:
        if act is not None:
            act.uniform_(-1, 1)             # a stand-in policy: fresh random actions every 50 steps
            if a.mode == "omega_thrust":
                act[:, 0] *= 0.2
        g.replay()
        done_steps += chunk
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    env.check_finite()
    st = all_reduce_stats(env.stats_tensor())
    print("%d envs x %d steps in %.3f s = %.3e env-steps/s (kernel %s)" % (a.num_envs, done_steps, dt, a.num_envs * done_steps / dt,
                                                                          env.kernel_variant))
    print("finished episodes %d, mean return %.4f +- %.4f" % (st["episodes"], st["mean"], st["std"]))


if __name__ == "__main__":
    main()
