"""GPU: coarse guards on the launch knobs of the bench headline (65,536-env cfg2 steps cycling through a ring of batches
larger than L2, CUDA-graph replays).  Not a benchmark: the point is that a default silently switched off is caught here and
not in the round-end bench - this file exists because a stray preprocessor line once dropped `spec_draw` to 0 (5.8 -> 6.7 us)
while every parity test stayed green."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_speculative_draws_and_prefetch_are_on_by_default_and_pay():
    sys.path.insert(0, ROOT)
    from tools.ow_ab import run
    for k in ("QUADSIM_SPEC_DRAW", "QUADSIM_PREFETCH", "QUADSIM_OW_THREADS", "QUADSIM_OW_EXTRA"):
        os.environ.pop(k, None)
    best_default, _, variant = run(65536, 1002, 16, "cfg2")
    assert variant == "tma-raw"
    best_off, _, _ = run(65536, 1002, 16, "cfg2", QUADSIM_SPEC_DRAW=0, QUADSIM_PREFETCH=0)
    for k in ("QUADSIM_SPEC_DRAW", "QUADSIM_PREFETCH"):
        os.environ.pop(k, None)
    print("65,536-env step: defaults %.3f us, speculative draws + L2 prefetch off %.3f us" % (best_default, best_off))
    assert best_default < 0.93 * best_off, (best_default, best_off)      # measured 5.8 vs 7.4
    assert best_default < 6.3, best_default                              # 0.55 of the HBM roofline: a floor, not the target


def test_mid_size_launch_is_one_even_wave_up_to_the_occupancy_the_runtime_reports():
    """81,920 envs = 1,280 64-thread CTAs: nine per SM fit (the one-wave instantiation needs ~90 registers), so the launch is ONE
    evenly spread wave; the launch-bound formula of round 2 (eight per SM) left 96 CTAs to a second wave (9.6 vs 7.2 us)."""
    sys.path.insert(0, ROOT)
    from tools.ow_ab import run
    for k in ("QUADSIM_SPEC_DRAW", "QUADSIM_PREFETCH", "QUADSIM_OW_THREADS", "QUADSIM_OW_EXTRA", "QUADSIM_OW_REGCAP"):
        os.environ.pop(k, None)
    n = 81920
    ring = max(1, min(16, (256 << 20) // (n * 221)))
    steps = 1008 - 1008 % ring
    best_default, _, variant = run(n, steps, ring, "cfg2")
    assert variant == "tma-raw"
    print("81,920-env step: %.3f us" % best_default)
    assert best_default < 8.4, best_default                               # measured 7.2 (round 2: 9.6)
