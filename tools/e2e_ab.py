"""A/B of QUADSIM_HOST_CTAS (resident CTAs per SM of a qs_step_host zero-copy launch) on the e2e number of bench.py:
65,536 envs, pinned host buffers, one synchronous call per step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    n, dev = bench.N_BATCH, torch.device("cuda", 0)
    a_h = (torch.rand((n, 4)) * 2 - 1).pin_memory()
    o_h, r_h = torch.empty((n, 18)).pin_memory(), torch.empty((n,)).pin_memory()
    d_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
    for rep in range(2):
        for k in (0, 1, 2, 3, 4, 6, 8):
            os.environ["QUADSIM_HOST_CTAS"] = str(k)
            e = bench.make_env(n, dev, seed=1)
            e.reset()
            for _ in range(20):
                e.step_host(a_h, o_h, r_h, d_h)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(1000):
                e.step_host(a_h, o_h, r_h, d_h)
            torch.cuda.synchronize()
            us = (time.perf_counter() - t0) / 1000 * 1e6
            print("host_ctas %d : %.1f us per step  %.3e env-steps/s  %.1f GB/s over the link  [%s]" % (k, us, n / us * 1e6, n * 93 / us * 1e-3, e.host_path), flush=True)
            e.close()
