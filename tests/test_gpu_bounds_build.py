"""GPU: the QS_DEBUG_BOUNDS build of the library (python -m gym_art_b200._build --variant bounds -> _exp/libquadsim_bounds.so).
compute-sanitizer is closed on the GPU pool, so memory safety of the tile ring, the work counter / ticket hand-off and
every derived env / row / tile index is argued by (a) this build, in which each of them is checked against its bound before
the access and failures are reported through qs_check, run over the awkward sizes of tools/sanitize_case.py (ragged tails,
one-wave and pipelined launches, swarms, per-env models, K-step, host pipeline), and (b) the parity suites over the same
sizes.  The self-test proves the reporting path: a deliberately failing check must surface as an error."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gym_art_b200", "_exp", "libquadsim_bounds.so")


def _run(extra_env):
    env = dict(os.environ, QUADSIM_LIB=LIB, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_case.py")], capture_output=True, text=True, env=env,
                          timeout=600)


@pytest.mark.skipif(not os.path.exists(LIB), reason="bounds-checked library not built (python -m gym_art_b200._build --variant bounds)")
def test_bounds_checked_build_runs_clean_over_awkward_sizes():
    r = _run({})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok ") >= 8, r.stdout


@pytest.mark.skipif(not os.path.exists(LIB), reason="bounds-checked library not built (python -m gym_art_b200._build --variant bounds)")
def test_a_failing_check_is_reported():
    r = _run({"QUADSIM_BOUNDS_SELFTEST": "1"})
    assert r.returncode != 0 and "QS_DEBUG_BOUNDS" in (r.stdout + r.stderr), r.stdout[-2000:] + r.stderr[-2000:]
