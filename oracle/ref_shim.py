"""TEST INFRASTRUCTURE ONLY - not part of the product path.

Imports the *live* reference (amolchanov86/gym_art, mounted read-only at /root/reference in the build
container) under ``sys.modules`` stubs for the packages it imports but never uses for arithmetic
(gym, matplotlib, transforms3d, noise, tensorflow, pyglet).

The live reference exists only in the build container.  It is used
  * by ``oracle/gen_golden.py`` to generate the committed fixtures under ``tests/golden/``, and
  * by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent) to pin the
    NumPy restatement ``oracle/quadsim_oracle.py`` against the real code.
Nothing under ``gym_art_b200/`` may import this file.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("QUADSIM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_art", "quadrotor"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    """Install import-time-only stand-ins.  None of them takes part in the simulated arithmetic."""
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_quadsim_stub", False):
        return

    class Box(object):
        def __init__(self, low, high, dtype=np.float32, shape=None):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = dtype

    class Env(object):
        metadata = {}
        spec = None

    class EzPickle(object):
        def __init__(*args, **kwargs):
            pass

    class EnvSpec(object):
        def __init__(self, id, max_episode_steps=None, **kw):
            self.id = id
            self.max_episode_steps = max_episode_steps

    def np_random(seed=None):
        # (deterministic also without a seed: an unseeded generator made recorded rollouts differ from run to run)
        return np.random.RandomState(0 if seed is None else seed), seed

    class _Err(Exception):
        pass

    spaces = _mod("gym.spaces", Box=Box)
    seeding = _mod("gym.utils.seeding", np_random=np_random)
    utils = _mod("gym.utils", EzPickle=EzPickle, seeding=seeding)
    registration = _mod("gym.envs.registration", EnvSpec=EnvSpec)
    envs = _mod("gym.envs", registration=registration)
    error = _mod("gym.error", Error=_Err, DependencyNotInstalled=_Err)
    _mod("gym", Env=Env, spaces=spaces, utils=utils, envs=envs, error=error, _quadsim_stub=True)

    pyplot = _mod("matplotlib.pyplot")
    _mod("matplotlib", pyplot=pyplot)
    # transforms3d is absent here; QuadrotorDynamics.pitch_roll_restricted_random_state (quadrotor.py:257) needs
    # euler.euler2mat with its default 'sxyz' axes - restated from its documented convention in the oracle
    def _euler2mat(ai, aj, ak, axes="sxyz"):
        assert axes == "sxyz"
        from . import quadsim_oracle as _qo
        return _qo.euler2mat_sxyz(np.float64(ai), np.float64(aj), np.float64(ak))

    _mod("transforms3d", euler=_mod("transforms3d.euler", euler2mat=_euler2mat))
    _mod("noise")
    _mod("tensorflow")

    class Group(object):
        def __init__(self, parent=None):
            self.parent = parent

    class Batch(object):
        pass

    gl = _mod("pyglet.gl")
    graphics = _mod("pyglet.graphics", Group=Group, Batch=Batch)
    _mod("pyglet", options={}, gl=gl, graphics=graphics)


def import_reference():
    """Returns the reference modules as a namespace: .quadrotor .control .sensor_noise .quad_utils ..."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        import gym_art.quadrotor.quadrotor as quadrotor
        import gym_art.quadrotor.quadrotor_control as control
        import gym_art.quadrotor.sensor_noise as sensor_noise
        import gym_art.quadrotor.quad_utils as quad_utils
        import gym_art.quadrotor.quad_models as quad_models
        import gym_art.quadrotor.inertia as inertia
        import gym_art.quadrotor.get_state as get_state
        import gym_art.quadrotor.quadrotor_randomization as quad_rand
        import gym_art.quadrotor_multi.quadrotor_multi as quadrotor_multi
    # get_state.py uses three names it never imports (`normal` :334, `R2quat` :277, `copy` :217): supply them from where
    # the rest of the reference takes them, so the quaternion and t2w / t2t observation builders run as written.  `normal`
    # is looked up on np.random at call time, so a NoiseTape sees its draws.
    if not hasattr(get_state, "normal"):
        import copy as _copy
        get_state.normal = lambda *a, **k: np.random.normal(*a, **k)
        get_state.R2quat = quad_utils.R2quat
        get_state.copy = _copy
    ns = types.SimpleNamespace(
        quadrotor=quadrotor, control=control, sensor_noise=sensor_noise, quad_utils=quad_utils,
        quad_models=quad_models, inertia=inertia, get_state=get_state, quad_rand=quad_rand,
        quadrotor_multi=quadrotor_multi)
    return ns


class NoiseTape(object):
    """Record/replay of the reference's global-RNG draws ("RNG interception points").

    While active, the module-level names the reference draws from are replaced by functions that
    return fresh draws from a private ``RandomState`` *and* log (kind, unit-scale values).  The
    test harness regroups the log into the injected-noise slot layout of include/quadsim.h.
    kinds: "normal" (unit normal, before scaling), "uniform" (unit uniform in [-1,1) for symmetric
    ranges, else raw in [low, high) mapped to [0,1)), "randn".
    """

    def __init__(self, ref, seed=0):
        self.ref = ref
        self.rs = np.random.RandomState(seed)
        self.log = []
        self._saved = []

    # --- replacement draw functions -------------------------------------------------------
    def normal(self, loc=0.0, scale=1.0, size=None):
        z = self.rs.standard_normal(size)
        self.log.append(("normal", np.array(z, dtype=np.float64, copy=True)))
        return loc + scale * z

    def uniform(self, low=0.0, high=1.0, size=None):
        u = self.rs.uniform(-1.0, 1.0, size)  # unit symmetric uniform
        self.log.append(("uniform", np.array(u, dtype=np.float64, copy=True)))
        low = np.asarray(low, dtype=np.float64)
        high = np.asarray(high, dtype=np.float64)
        return 0.5 * (low + high) + 0.5 * (high - low) * u

    def randn(self, *shape):
        z = self.rs.standard_normal(shape if shape else None)
        self.log.append(("randn", np.array(z, dtype=np.float64, copy=True)))
        return z

    def rand(self, *shape):
        u = self.rs.uniform(-1.0, 1.0, shape if shape else None)
        self.log.append(("uniform", np.array(u, dtype=np.float64, copy=True)))
        return 0.5 + 0.5 * u

    # --- patching ---------------------------------------------------------------------------
    def _patch(self, obj, name, value):
        self._saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def __enter__(self):
        ref = self.ref
        self._patch(ref.sensor_noise, "normal", self.normal)
        self._patch(ref.sensor_noise, "uniform", self.uniform)
        fake_nr = types.SimpleNamespace(randn=self.randn)
        self._patch(ref.quad_utils, "nr", fake_nr)
        self._patch(np.random, "uniform", self.uniform)
        self._patch(np.random, "normal", self.normal)
        self._patch(np.random, "rand", self.rand)
        return self

    def __exit__(self, *exc):
        for obj, name, value in reversed(self._saved):
            setattr(obj, name, value)
        self._saved = []
        return False

    def attach_env(self, env):
        """The env's private RandomState draws the spawn position (quadrotor.py:1087)."""
        env.np_random = types.SimpleNamespace(uniform=self.uniform)

    def take(self):
        log, self.log = self.log, []
        return log
