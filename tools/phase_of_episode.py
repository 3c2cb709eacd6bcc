"""Step time as a function of the phase of the episode (graph-replayed launches, CUDA events): trajectories evolve - drones under
random actions end up on the floor, swarms pile up in contact - and data-dependent paths may cost more late in an episode.
    python tools/phase_of_episode.py cfg4 2097152 [cfg n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peak, _ = bench.load_peaks()
args = sys.argv[1:] or ["cfg4", "2097152"]
for cfg, n in zip(args[0::2], args[1::2]):
    n = int(n)
    e = bench.make_env(n, dev, seed=1, cfg=cfg)
    e.reset()
    a = None
    if cfg != "cfg5":
        a = torch.rand((n, 4), device=dev) * 2 - 1
        if cfg == "cfg3":
            a[:, 0] *= 0.2
    bench.RING_PERIOD[0] = 1
    K = 20
    g = e.capture_steps(K, a)
    done = 1
    line = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(52):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        line.append("%d:%.1f" % (done, us))
        done += K
    print("%s n=%d  us/step by first step of each %d-step window: %s" % (cfg, n, K, "  ".join(line)), flush=True)
    e.close()
    del e, a, g
    torch.cuda.empty_cache()
