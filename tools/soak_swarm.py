"""Long-horizon soak of the swarm path: 131,072 swarms of 8 x 20,040 steps (40 episodes; most of each episode in the contact-heavy
regime under random actions) through CUDA-graph replays; bookkeeping exact, state finite / orthonormal / inside the room."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gym_art_b200 import QuadVecEnv  # noqa: E402
from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals  # noqa: E402

n, episodes = (1 << 17) * 8, 40
env = QuadVecEnv(n, dynamics_params="DefaultQuad", controller="raw", n_agents=8, formation=ring_goals(8), ep_time=5, seed=11)
env.reset()
a = torch.rand((n, 4), device="cuda") * 2 - 1
g = env.capture_steps(501, a)
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
print("steps %d  %.2f s  %.3e drone-steps/s  ortho err %.2e  episodes %.0f (expected %d)  tick %d" % (
    steps, dt, n * 501 * episodes / dt, err, st[2], n * (steps // 501), int(env.tick[0])))
assert st[2] == n * (steps // 501) and err < 2e-5 and bool(torch.isfinite(env.state).all())
assert bool((env.tick == steps % 501).all())
pos = env.pos
assert float(pos.abs().max()) <= 10.0 and float(pos[:, 2].min()) >= 0.0
print("swarm soak ok")
