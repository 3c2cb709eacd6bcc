"""A/B of the one-wave CTA size (QUADSIM_OW_THREADS = 32 / 64 / 128) on the bench headline."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.ow_ab import run  # noqa: E402

if __name__ == "__main__":
    n, ring = 65536, 16
    for rep in range(2):
        for th in (32, 64, 128):
            best, med, v = run(n, 2004, ring, "cfg2", QUADSIM_OW_THREADS=th, QUADSIM_PREFETCH=3)
            print("threads %3d : %.3f us/step (median %.3f)  frac %.3f  [%s]" % (th, best, med, n * 349 / (best * 1e-6) / 1e9 / 6533.2, v), flush=True)
