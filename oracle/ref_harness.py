"""TEST INFRASTRUCTURE ONLY - drives the live reference under a NoiseTape and regroups its RNG draws
into the injected-noise slot layout shared by the oracle and the CUDA path (include/quadsim.h).

Only usable where /root/reference exists (the build container).  Used by oracle/gen_golden.py and
tests/test_oracle_vs_reference.py.
"""
import contextlib
import io

import numpy as np

from . import ref_shim
from . import quadsim_oracle as qo


def _norm3(x):
    return (x[0] ** 2 + x[1] ** 2 + x[2] ** 2) ** 0.5


def _normalize(x):
    n = _norm3(x)
    return x if n < 0.00001 else x / n


class _TapeReader(object):
    def __init__(self, log):
        self.log = log
        self.i = 0

    def pop(self, kind, shape):
        k, v = self.log[self.i]
        kinds = (kind,) if isinstance(kind, str) else kind
        assert k in kinds and np.shape(v) == shape, "tape mismatch at %d: got %s%s want %s%s" % (
            self.i, k, np.shape(v), kind, shape)
        self.i += 1
        return v

    def peek(self):
        return self.log[self.i] if self.i < len(self.log) else (None, None)

    def done(self):
        return self.i == len(self.log)


def _read_add_noise(tr, gyro_bias_mode, z=None):
    """One SensorNoise.add_noise call (sensor_noise.py:116-157) -> OBS_Z unit draws."""
    z = np.zeros(qo.OBS_Z) if z is None else z
    z[qo.Z_NPOS:qo.Z_NPOS + 3] = tr.pop("normal", (3,))
    z[qo.Z_UPOS:qo.Z_UPOS + 3] = tr.pop("uniform", (3,))
    z[qo.Z_NVEL:qo.Z_NVEL + 3] = tr.pop("normal", (3,))
    z[qo.Z_UVEL:qo.Z_UVEL + 3] = tr.pop("uniform", (3,))
    z[qo.Z_NGYRO:qo.Z_NGYRO + 3] = tr.pop("normal", (3,))
    if gyro_bias_mode:
        z[qo.Z_NGYRO_RW:qo.Z_NGYRO_RW + 3] = tr.pop("normal", (3,))
    z[qo.Z_NTHETA:qo.Z_NTHETA + 3] = tr.pop("normal", (3,))
    z[qo.Z_UTHETA:qo.Z_UTHETA + 3] = tr.pop("uniform", (3,))
    z[qo.Z_NACC_S:qo.Z_NACC_S + 3] = tr.pop("normal", (3,))
    z[qo.Z_NACC_D:qo.Z_NACC_D + 3] = tr.pop("normal", (3,))
    return z


class ReferenceRunner(object):
    """One live reference QuadrotorEnv, stepped with recorded noise."""

    def __init__(self, seed=0, controller=None, multi=False, **env_kwargs):
        self.ref = ref_shim.import_reference()
        self.tape = ref_shim.NoiseTape(self.ref, seed=seed)
        self.env_kwargs = env_kwargs
        with self.tape, contextlib.redirect_stdout(io.StringIO()):
            mod = self.ref.quadrotor_multi if multi else self.ref.quadrotor
            self.env = mod.QuadrotorEnv(**env_kwargs)
            self.tape.attach_env(self.env)
        self.tape.take()                       # draws made by the ctor's own _reset() are not replayed
        # ... and how many of them there were must not shift the stream the rollout is recorded from (the ctor's default
        # spawn uses the env's own RandomState and a rejection loop on it): start the rollout's draws from a fresh,
        # seeded generator, so that a fixture is a function of (case, seed) alone - whatever ran before in the process
        self.tape.rs = np.random.RandomState(seed)
        self.noise_on = not self.env.sense_noise.bypass
        self.gyro_bias_mode = self.noise_on and self.env.sense_noise.gyro_norm_std != 0.0
        comps = self.env.obs_repr.split("_")
        self.obs_t2w, self.obs_t2t = "t2w" in comps, "t2t" in comps       # get_state.py:334, :379: one normal(size=1) each
        self.obs_draws = self.noise_on or self.obs_t2w
        # optional controller the reference env cannot select itself
        self.ext_controller = None
        if controller == "omega_thrust":
            self.ext_controller = self.ref.control.OmegaThrustControl(self.env.dynamics)
        elif controller == "shifted":
            self.ext_controller = self.ref.control.ShiftedMotorControl(self.env.dynamics)
        elif controller == "velocity_yaw":
            c = self.ref.control.VelocityYawControl(self.env.dynamics)
            # quadrotor_control.py:276 calls an undefined name; restate with the commented line :275
            self.ref.control.get_blas_funcs = lambda name, args: np.dot(args[0], args[1])
            self.ext_controller = c
        if self.ext_controller is not None:
            ext = self.ext_controller

            class _Adapter(object):          # gives step_func the keyword signature _step uses (:966-970)
                action = np.zeros(4)

                def step_func(self, dynamics, action, goal, dt, observation=None):
                    ext.step(dynamics, np.array(action, dtype=np.float64), dt)

            self.env.controller = _Adapter()
        if not hasattr(self.env.controller, "action"):
            # Vertical/VertPlaneControl never set .action, which quadrotor.py:1002 reads for the info
            # dict (reference bug: AttributeError on the first step).  Harness-only workaround.
            self.env.controller.action = np.zeros(4)

    def _read_state_call(self, tr):
        """One state_<obs_repr>(env) call: add_noise (unless bypassed), then the t2w / t2t normals -> OBS_Z unit draws."""
        z = np.zeros(qo.OBS_Z)
        if self.noise_on:
            _read_add_noise(tr, self.gyro_bias_mode, z)
        if self.obs_t2w:
            z[qo.Z_NT2W] = tr.pop("normal", (1,))[0]
        if self.obs_t2t:
            z[qo.Z_NT2T] = tr.pop("normal", (1,))[0]
        return z

    # ---------------------------------------------------------------------------------------------
    def _reset_z_from_tape(self, log, pos_goal):
        tr = _TapeReader(log)
        rz = np.zeros(qo.RESET_Z)
        env = self.env
        self.last_goal_z = np.zeros(qo.GOAL_Z)
        if env.resample_goal:                                             # quadrotor.py:1078-1079, drawn before the spawn
            self.last_goal_z[3] = tr.pop("uniform", ())
        rz[qo.RZ_UPOS:qo.RZ_UPOS + 3] = tr.pop("uniform", (3,))            # quadrotor.py:1087
        if env.init_random_state and env.dim_mode == "1D":                # quadrotor.py:1101-1103
            rz[qo.RZ_UVELMAG] = tr.pop("uniform", ())
        elif env.init_random_state and env.dim_mode == "2D":              # :1104-1110: omega, vel (3 draws), theta
            rz[qo.RZ_UOMMAG] = tr.pop("uniform", ())
            rz[qo.RZ_UVEL:qo.RZ_UVEL + 3] = tr.pop("uniform", (3,))
            rz[qo.RZ_YAW] = tr.pop("uniform", ())
        elif env.init_random_state:                                       # quadrotor.py:227-239
            tr.pop("uniform", (3,))                                       # pos draw, discarded (:228, :1113 `_`)
            rz[qo.RZ_UVEL:qo.RZ_UVEL + 3] = tr.pop("uniform", (3,))
            rz[qo.RZ_UVELMAG] = tr.pop("uniform", ())
            rz[qo.RZ_UOM:qo.RZ_UOM + 3] = tr.pop("uniform", (3,))
            rz[qo.RZ_UOMMAG] = tr.pop("uniform", ())
            up_z = tr.pop("normal", (3,))                                 # quad_utils.py:47-57
            up = _normalize(up_z)
            while True:
                fwd_z = tr.pop("normal", (3,))
                if not np.dot(_normalize(fwd_z), up) > 0.95:
                    break
            rz[qo.RZ_NUP:qo.RZ_NUP + 3] = up_z
            rz[qo.RZ_NFWD:qo.RZ_NFWD + 3] = fwd_z
        elif env.dim_mode == "3D":                                        # quadrotor.py:1124-1126
            while True:
                u = tr.pop("uniform", ())
                k, v = tr.peek()
                if not (k == "uniform" and np.shape(v) == ()):
                    break
            rz[qo.RZ_YAW] = u
        if self.obs_draws:
            rz[qo.RZ_OBS:qo.RZ_OBS + qo.OBS_Z] = self._read_state_call(tr)
        assert tr.done(), "unparsed reset draws"
        return rz

    def reset(self):
        with self.tape:
            obs = self.env.reset()
        rz = self._reset_z_from_tape(self.tape.take(), None)
        return np.array(obs, dtype=np.float64), rz

    def step(self, action):
        env = self.env
        with self.tape:
            obs, rew, done, info = env.step(np.array(action, dtype=np.float64))
        tick_before = env.tick - 1
        tr = _TapeReader(self.tape.take())
        zs = []
        self.last_goal_z = np.zeros(qo.GOAL_Z)
        if self.obs_draws:                                                # quadrotor.py:947: a bare add_noise call, no t2w draw
            zs.append(_read_add_noise(tr, self.gyro_bias_mode) if self.noise_on else np.zeros(qo.OBS_Z))
        if env.excite and tick_before % 5 == 0:                           # :957-962
            self.last_goal_z[0:2] = tr.pop("uniform", (2,))
            self.last_goal_z[2] = tr.pop("uniform", (1,))[0]
        if self.obs_draws:
            zs.append(self._read_state_call(tr))                          # :970
        ou_z = np.zeros((env.sim_steps, 4))
        for s in range(env.sim_steps):
            ou_z[s] = tr.pop("randn", (4,))                               # quad_utils.py:199
        obs_z = np.zeros(qo.OBS_Z)
        if self.obs_draws:
            zs.append(self._read_state_call(tr))                          # :988
            obs_z = zs[2].copy()
            if self.gyro_bias_mode:
                # fold the three bias-walk updates into one effective unit normal (see oracle add_noise)
                pi1, _ = qo.gyro_bias_coeffs(vars(env.sense_noise), env.dt)
                w = np.array([pi1 ** 2, pi1, 1.0])
                obs_z[qo.Z_NGYRO:qo.Z_NGYRO + 3] = sum(
                    w[k] * zs[k][qo.Z_NGYRO:qo.Z_NGYRO + 3] for k in range(3)) / np.sqrt(np.sum(w ** 2))
        assert tr.done(), "unparsed step draws"
        return np.array(obs, dtype=np.float64), float(rew), bool(done), info, ou_z, obs_z

    def state(self):
        d = self.env.dynamics
        return dict(pos=d.pos.copy(), vel=d.vel.copy(), rot=d.rot.copy(), omega=np.array(d.omega, dtype=np.float64),
                    rot_damp=d.thrust_rot_damp.copy(), cmds_damp=d.thrust_cmds_damp.copy(),
                    ou=d.thrust_noise.state.copy(), accelerometer=np.array(d.accelerometer, dtype=np.float64),
                    svd_timer=float(d.since_last_svd))


def oracle_kwargs_from_env(env, controller_id):
    """Maps a live reference env's configuration onto OracleVecEnv kwargs."""
    sn = None
    if not env.sense_noise.bypass:
        sn = {k: getattr(env.sense_noise, k) for k in qo.DEFAULT_SENSE_NOISE}
    return dict(controller=controller_id, zero_middle=env.raw_control_zero_middle, sim_freq=1.0 / env.dt,
                sim_steps=env.sim_steps, obs_repr=env.obs_repr, ep_time=env.ep_time, room_size=env.room_size,
                init_random_state=env.init_random_state, rew_coeff=dict(env.rew_coeff), sense_noise=sn,
                gravity=env.gravity, dim_mode=env.dim_mode, multi_reward="pos_log_weight" in env.rew_coeff,
                resample_goal=bool(env.resample_goal), excite=bool(env.excite),
                dynamics_simplification=bool(env.dynamics_simplification), t2w_std=float(getattr(env, "t2w_std", 0.005)),
                t2t_std=float(getattr(env, "t2t_std", 0.0005)))
