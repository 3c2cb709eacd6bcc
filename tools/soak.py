"""Long-horizon soak: 1 M envs x 100,200 steps (200 episodes each) through CUDA-graph replays; checks bookkeeping stays
exact and the state stays sane (orthonormal R, finite, inside the room)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gym_art_b200 import QuadVecEnv  # noqa: E402

n, episodes = 1 << 20, 200
env = QuadVecEnv(n, dynamics_params="Crazyflie", controller="omega_thrust", sense_noise="default", init_random_state=True,
                 ep_time=5, seed=9)
env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
a[:, 0] *= 0.2
g = env.capture_steps(501, a)            # runs one warm-up step itself
t0 = time.perf_counter()
for ep in range(episodes):
    g.replay()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
env.check_finite()
steps = 1 + 501 * episodes
R = env.rot
err = float((R.transpose(1, 2) @ R - torch.eye(3, device="cuda")).abs().amax())
st = env.stats()
svd_t, step_count = env.clock
print("steps %d  %.2f s  %.3e env-steps/s  ortho err %.2e  episodes %.0f (expected %d)  tick %d  svd_timer %.4f  step_count %d" % (
    steps, dt, n * 501 * episodes / dt, err, st[2], n * (steps // 501), int(env.tick[0]), svd_t, step_count))
assert st[2] == n * (steps // 501) and err < 2e-5 and bool(torch.isfinite(env.state).all())
assert bool((env.tick == steps % 501).all()) and step_count == steps + 1
pos = env.pos
assert float(pos.abs().max()) <= 10.0 and float(pos[:, 2].min()) >= 0.0
print("soak ok")
