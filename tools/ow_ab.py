"""A/B of the one-wave launch knobs on the bench headline (ring of 16 x 65,536-env cfg2 batches, CUDA-graph replays):
QUADSIM_OW_THREADS (CTA size), QUADSIM_PREFETCH (L2 prefetch ahead of griddepcontrol.wait), QUADSIM_SPEC_DRAW (draws ahead
of the wait from a speculative clock read).  The knobs are read at qs_create, so every combination gets fresh handles.

    python tools/ow_ab.py [n_envs] [steps]
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def run(n, steps, ring, cfg, repeats=3, lead_in=0, **knobs):
    for k, v in knobs.items():
        os.environ[k] = str(v)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    envs, acts = [], []
    for b in range(ring):
        e = bench.make_env(n, dev, seed=1, env_id_offset=b * n, cfg=cfg)
        e.reset()
        envs.append(e)
        acts.append(None if cfg == "cfg5" else torch.rand((n, 4), device=dev, generator=gen) * 2 - 1)
    bench.RING_PERIOD[0] = ring
    ms, _ = bench.time_graph(torch, lambda i: envs[i % ring].step(acts[i % ring]), steps, 20, dev, repeats=repeats, lead_in=lead_in)
    v = envs[0].kernel_variant
    for e in envs:
        e.check_finite()
        e.close()
    return min(ms) * 1e3 / steps, bench.median(ms) * 1e3 / steps, v


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4008
    cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
    ring = max(1, min(16, (256 << 20) // (n * 221)))
    print("n_envs %d, ring %d, %d steps per timing, %s" % (n, ring, steps, cfg))
    best, med, v = run(n, steps, ring, cfg, QUADSIM_OW_THREADS=64, QUADSIM_PREFETCH=3)
    print("steady, %d steps : %.3f us/step (median %.3f)  frac %.3f  [%s]" % (steps, best, med, n * 349 / (best * 1e-6) / 1e9 / 6533.2, v), flush=True)
    for lead in (0, 4, 16):                                                                           # the driver's regime
        best, med, v = run(n, 20, ring, cfg, repeats=9, lead_in=lead, QUADSIM_OW_THREADS=64, QUADSIM_PREFETCH=3)
        print("K=20 lead_in %2d : %.3f us/step (median %.3f)" % (lead, best, med), flush=True)
