import os, sys
sys.path.insert(0, "/root/repo")
import torch
from gym_art_b200.groups import QuadVecEnvGroup
dev = torch.device("cuda", 0)
n, G, K = 65536, 16, 320
grp = QuadVecEnvGroup(G, n, dynamics_params="DefaultQuad", controller="raw", ep_time=5, seed=1)
grp.reset()
acts = [torch.rand((n, 4), device=dev) * 2 - 1 for _ in range(G)]
g = grp.capture_steps(K, acts)
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("%d groups x %d envs, %d steps each: %.3f us per 65,536-env step, %.3e env-steps/s, frac %.3f" % (
    G, n, K, ms * 1e3 / (G * K), G * K * n / (ms * 1e-3), G * K * n * 349 / (ms * 1e-3) / 6533.2e9))
for e in grp.envs:
    e.check_finite()
