"""A/B of QUADSIM_OW_EXTRA (CTA slots per SM left free for the dependent launch of a one-wave step) on the bench headline.

    python tools/ow_extra.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.ow_ab import run  # noqa: E402

if __name__ == "__main__":
    n, ring = 65536, 16
    for extra in (0, 1, 2, 3):
        os.environ["QUADSIM_OW_EXTRA"] = str(extra)
        best, med, v = run(n, 2004, ring, "cfg2", QUADSIM_OW_THREADS=64, QUADSIM_PREFETCH=3)
        print("extra %d : %.3f us/step (median %.3f)  frac %.3f  [%s]" % (extra, best, med, n * 349 / (best * 1e-6) / 1e9 / 6533.2, v), flush=True)
