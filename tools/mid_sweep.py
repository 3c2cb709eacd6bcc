"""Step time at mid-size batches (between one wave and a deep tile pipeline): python tools/mid_sweep.py [cfg]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
for n in [int(x) for x in os.environ.get("MID_SIZES", "73728,81920,90112,98304,114688,131072,147456,163840,196608,262144").split(",")]:
    ring = max(1, min(16, (300 << 20) // (n * 349) + 1))
    envs, acts = [], []
    for b in range(ring):
        e = bench.make_env(n, dev, seed=1, env_id_offset=b * n, cfg=cfg)
        e.reset()
        envs.append(e)
        a = torch.rand((n, 4), device=dev) * 2 - 1
        if cfg == "cfg3":
            a[:, 0] *= 0.2
        acts.append(a)
    bench.RING_PERIOD[0] = ring
    K = 1600 - 1600 % ring
    ms = bench.time_graph(torch, lambda i: envs[i % ring].step(acts[i % ring]), K, 4 * ring, dev, chunk=K)[0][0]
    print("%s n=%8d tiles/slot %.2f  %8.2f us/step" % (cfg, n, n / 128 / 592, ms * 1e3 / K), flush=True)
    for e in envs:
        e.close()
    del envs, acts
    torch.cuda.empty_cache()
