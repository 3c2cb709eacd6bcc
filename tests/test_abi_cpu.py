"""CPU: the C-ABI library loads without a GPU, exports every symbol include/quadsim.h declares, and its pure
host entry points (layout, validation) behave.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from gym_art_b200 import _lib
from gym_art_b200.quadrotor.model import ModelConstants
from gym_art_b200.vec_env import make_config, resolve_dynamics_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gym_art_b200 import _build
    _build.build()
    return _lib.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "quadsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qs_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    declared = header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "libquadsim.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared, "binding list and header disagree"


def test_abi_version_and_struct_size(lib):
    assert lib.qs_abi_version() == _lib.QS_ABI_VERSION
    header = open(os.path.join(ROOT, "include", "quadsim.h")).read()
    assert "#define QS_ABI_VERSION %d" % _lib.QS_ABI_VERSION in header
    assert "#define QS_OBS_Z %d" % _lib.QS_OBS_Z in header and "#define QS_RESET_Z %d" % _lib.QS_RESET_Z in header


def _cfg(**kw):
    model = ModelConstants(resolve_dynamics_params(kw.pop("quad", "DefaultQuad")))
    return make_config(kw.pop("n", 1024), model, **kw)[0]


def test_state_layout_rows(lib):
    lay = _lib.state_layout(_cfg())
    assert (lay.rows, lay.row_ou, lay.row_prev_action, lay.row_gyro_bias, lay.row_goal) == (8, 7, -1, -1, -1)
    assert (lay.obs_dim, lay.act_dim) == (18, 4)
    lay = _lib.state_layout(_cfg(obs_repr="xyz_vxyz_R_omega_act", sense_noise={"gyro_norm_std": 0.1}, per_env_goal=True,
                                 controller="vertplane"))
    assert (lay.rows, lay.row_ou, lay.row_prev_action, lay.row_gyro_bias, lay.row_goal) == (11, 7, 8, 9, 10)
    assert (lay.obs_dim, lay.act_dim) == (22, 2)
    dims = {"xyz_vxyz_R_omega": 18, "xyz_vxyz_R_omega_h": 19, "xyzr_vxyzr_R_omega": 18, "xyzr_vxyzr_R_omega_h": 19,
            "xyz_vxyz_R_omega_acc_act": 25, "xyz_vxyz_R_omega_act": 22, "xyz_vxyz_quat_omega": 13, "xyzr_vxyzr_quat_omega": 13,
            "xyzr_vxyzr_quat_omega_h": 14, "xyzr_vxyzr_R_omega_t2w": 19, "xyz_vxyz_R_omega_t2w": 19, "xyz_vxyz_R_omega_t2w_t2t": 20}
    assert sorted(dims) == sorted(_lib.OBS_REPRS)
    for rep, d in dims.items():
        assert _lib.state_layout(_cfg(obs_repr=rep)).obs_dim == d


def test_validation_errors(lib):
    lay = _lib.QsStateLayout()
    cfg = _cfg()
    cfg.abi_version = 1
    assert lib.qs_state_layout(C.byref(cfg), C.byref(lay)) == _lib.QS_EABI
    cfg = _cfg()
    cfg.n_agents = 3
    assert lib.qs_state_layout(C.byref(cfg), C.byref(lay)) == _lib.QS_EINVAL
    assert b"n_agents" in lib.qs_last_error(None)
    cfg = _cfg()
    cfg.sim_steps = 0
    assert lib.qs_state_layout(C.byref(cfg), C.byref(lay)) == _lib.QS_EINVAL
    cfg = _cfg(dim_mode="1D", controller="vertical")
    cfg.init_random_state = _lib.QS_INIT_RANDOM                          # quadrotor.py:1101-1103: legal in every dim_mode
    assert lib.qs_state_layout(C.byref(cfg), C.byref(lay)) == _lib.QS_OK
    cfg.init_random_state = _lib.QS_INIT_PITCH_ROLL_RESTRICTED           # a 3D law (Euler angles -> R)
    assert lib.qs_state_layout(C.byref(cfg), C.byref(lay)) == _lib.QS_EINVAL
    cfg.init_random_state = 7
    assert lib.qs_state_layout(C.byref(cfg), C.byref(lay)) == _lib.QS_EINVAL
    with pytest.raises(_lib.QuadsimError):
        _lib.check(_lib.QS_EINVAL)
    # NULL handles are rejected, not dereferenced
    assert lib.qs_step(None, None, None, None, None, None) == _lib.QS_EINVAL
    assert lib.qs_destroy(None) == _lib.QS_EINVAL


def test_reference_value_errors():
    with pytest.raises(ValueError, match="Unknown dimensionality"):
        _cfg(dim_mode="4D")
    with pytest.raises(ValueError, match="sense_noise"):
        _cfg(sense_noise=3.0)
    with pytest.raises(ValueError):
        _cfg(obs_repr="xyz_vxyz_euler_omega")                             # cannot run in the reference (undefined local `euler`, get_state.py:426)
    with pytest.raises(AssertionError):
        _cfg(rew_coeff={"poss": 1.0})


def test_env_needs_cuda_and_does_not_fall_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_art_b200 import QuadVecEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        QuadVecEnv(8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gym_art_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_every_translation_unit_is_in_the_build_list_and_every_family_launcher_is_defined():
    """A .cu file under csrc/ that the build script does not list would silently drop its kernels (and every launcher the device
    header declares must be defined by exactly one translation unit, or the library would not load)."""
    import glob
    import re
    from gym_art_b200 import _build
    csrc = os.path.join(ROOT, "gym_art_b200", "csrc")
    on_disk = sorted(os.path.basename(p) for p in glob.glob(os.path.join(csrc, "*.cu")))
    assert on_disk == sorted(_build.SOURCES)
    declared = re.findall(r"QS_DECL_FAST\((launch_\w+)\);", open(os.path.join(csrc, "qs_device.cuh")).read())
    defined = []
    for f in on_disk:
        m = re.search(r"#define QS_(?:FAST|GEN)_NAME (launch_\w+)", open(os.path.join(csrc, f)).read())
        if m:
            defined.append(m.group(1))
    assert sorted(declared) == sorted(defined) and len(set(defined)) == len(defined)
