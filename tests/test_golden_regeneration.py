"""CPU (build container only: needs /root/reference): oracle/gen_golden.py is a hermetic recipe - regenerating every
committed fixture from the live reference reproduces it array for array, in any order (a fixture is a function of its
case and seed alone; see ReferenceRunner.__init__ for the state that used to leak)."""
import os

import numpy as np
import pytest

from oracle import ref_shim
from tests.golden_util import ANSWER_FILES, GOLDEN_DIR

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="needs the reference tree (/root/reference)")


def _same(name, fresh):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    assert sorted(z.files) == sorted(fresh.keys()), name
    for k in z.files:
        a, b = np.asarray(fresh[k]), np.asarray(z[k])
        assert a.shape == b.shape and a.dtype == b.dtype, (name, k)
        assert np.array_equal(a, b), (name, k)


def test_every_rollout_fixture_regenerates_bit_for_bit_in_any_order():
    from oracle import gen_golden as gg
    names = list(gg.CASES)
    assert sorted(names) == sorted(os.path.splitext(f)[0] for f in os.listdir(GOLDEN_DIR)
                                   if f.endswith(".npz") and f not in ANSWER_FILES)
    for name in reversed(names):                 # not the order the committed files were written in
        _same(name, gg.rollout(name))
    _same(names[0], gg.rollout(names[0]))        # and once more at the end of the process


def test_answer_files_regenerate():
    from oracle import gen_golden as gg
    assert sorted(n + ".npz" for n in gg.ANSWER_FILES) == sorted(ANSWER_FILES)
    for name, fn in gg.ANSWER_FILES.items():
        _same(name, fn())


def test_all_fixtures_carry_every_key():
    for f in os.listdir(GOLDEN_DIR):
        if f.endswith(".npz") and f not in ANSWER_FILES:
            z = np.load(os.path.join(GOLDEN_DIR, f))
            for k in ("goal", "goal_z", "reset_goal_z", "ou_z", "obs_z", "reset_z", "meta_json"):
                assert k in z.files, (f, k)
