"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-golden (GPU) tests.

The fixtures under tests/golden/ were produced from the live reference by oracle/gen_golden.py.
"""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# fixtures that are not single-env rollouts
ANSWER_FILES = ("known_answer.npz", "model_answers.npz", "random_models.npz", "affine_answers.npz", "restricted_state.npz")

ROLLOUT_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                       if os.path.basename(p) not in ANSWER_FILES)


class Golden(object):
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.a = {k: z[k] for k in z.files if k != "meta_json"}
        self.meta = json.loads(str(z["meta_json"]))
        self.params = self.meta["params"]
        self.oracle_kwargs = self.meta["oracle_kwargs"]
        self.controller = self.meta["controller"]
        self.T = self.a["obs"].shape[0]

    def __getattr__(self, k):
        try:
            return self.__dict__["a"][k]
        except KeyError:
            raise AttributeError(k)

    def reset_index_after(self, t):
        """index into reset_z / reset_obs of the reset that follows step t (which must be a done step)."""
        return int(np.searchsorted(self.reset_step, t + 1))


def load_npz(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def obs_tolerance(obs_repr, obs_row, base):
    """Per-element absolute tolerance for one observation row.  The quaternion builders use the reference's single-branch
    R2quat (quad_utils.py:101-109): w = sqrt(1 + trace) / 2, (x, y, z) = (...) / (4 w).  Near half-turn attitudes w -> 0 and
    rounding of the matrix entries is amplified by ~1 / w^2 in all four components - a property of the formula, in the
    reference as much as here - so those four elements get base * max(1, 1 / w^2)."""
    tol = np.full(np.shape(obs_row), float(base))
    if "quat" in obs_repr.split("_"):
        w = np.abs(np.asarray(obs_row)[..., 6])
        tol[..., 6:10] = (base * np.maximum(1.0, 1.0 / np.maximum(w, 1e-6) ** 2))[..., None]
    return tol
