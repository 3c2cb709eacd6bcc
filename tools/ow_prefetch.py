"""A/B of the L2 prefetch mode (QUADSIM_PREFETCH = 0 off / 1 bulk by thread 0 / 3 one line per thread) on the bench headline."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.ow_ab import run  # noqa: E402

if __name__ == "__main__":
    n, ring = 65536, 16
    for rep in range(3):
        for pf in (0, 1, 3):
            best, med, v = run(n, 4008, ring, "cfg2", QUADSIM_OW_THREADS=64, QUADSIM_PREFETCH=pf)
            print("prefetch %d : %.3f us/step (median %.3f)  frac %.3f  [%s]" % (pf, best, med, n * 349 / (best * 1e-6) / 1e9 / 6533.2, v), flush=True)
