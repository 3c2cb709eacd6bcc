"""Step time vs batch size (graph-replayed launches, CUDA events): python tools/sweep.py cfg2 cfg3 ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peak, _ = bench.load_peaks()
for cfg in sys.argv[1:] or ["cfg2"]:
    bytes_per = 333 if cfg == "cfg5" else 349
    for n in (131072, 262144, 524288, 1048576, 2097152, 4194304, 8388608):
        # a ring of envs large enough that every step streams from HBM (> 2 x L2)
        ring = max(1, min(16, (300 << 20) // (n * 349) + 1))
        envs, acts = [], []
        for b in range(ring):
            e = bench.make_env(n, dev, seed=1, env_id_offset=b * n, cfg=cfg)
            e.reset()
            envs.append(e)
            a = None
            if cfg != "cfg5":
                a = torch.rand((n, 4), device=dev) * 2 - 1
                if cfg == "cfg3":
                    a[:, 0] *= 0.2
            acts.append(a)
        bench.RING_PERIOD[0] = ring
        K = max(32, min(2000, (1 << 28) // n))
        K -= K % ring
        ms = bench.time_graph(torch, lambda i: envs[i % ring].step(acts[i % ring]), K, 4 * ring, dev, chunk=K)[0][0]
        us = ms * 1e3 / K
        print("%s n=%8d ring=%2d  %8.2f us/step  %.3e env-steps/s  frac %.3f" % (cfg, n, ring, us, n / us * 1e6, n * bytes_per / us * 1e-3 / peak), flush=True)
        for e in envs:
            e.close()
        del envs, acts
        torch.cuda.empty_cache()
