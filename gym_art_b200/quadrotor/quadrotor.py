"""Reference-shaped facades over the B200 env-step path: QuadrotorDynamics and QuadrotorEnv.

Same constructor arguments, methods and attributes as gym_art/quadrotor/quadrotor.py (QuadrotorDynamics :60-541,
QuadrotorEnv :647-1156), for code written against the reference's single-env objects.  Nothing is computed here:
every step / reset is one launch of the fused CUDA kernels through the C-ABI (include/quadsim.h), on a one-drone
(or `num_envs`-drone) QuadVecEnv whose state tensor lives on the GPU; attribute reads copy the field back.  The
facades default to the fp64 build, which tracks the float64 reference to 1e-12 per step; throughput work belongs
on QuadVecEnv (fp32, thousands to millions of envs per launch).

Differences from the reference, all forced by the device path and listed in DESIGN.md:
  * randomness comes from on-device Philox keyed by `seed` (or from injected draws in the parity tests), not
    from the global NumPy MT19937 stream;
  * rendering, obstacles and the TensorFlow controller are not part of this path.
"""
import copy

import numpy as np

from .. import _lib
from ..vec_env import QuadVecEnv
from . import quadrotor_randomization as quad_rand
from .model import GRAV, ModelConstants
from .quadrotor_control import (NonlinearPositionController, OmegaThrustControl, RawControl,  # noqa: F401  (re-exported:
                                ShiftedMotorControl, VelocityYawControl, VerticalControl,    # the reference module star-imports them)
                                VertPlaneControl)
from .sensor_noise import SensorNoise  # noqa: F401  (re-exported like the reference module does)
from .spaces import Box

EPS = 1e-6


class _OUView(object):
    """dynamics.thrust_noise: the OU process lives in the state tensor (quad_utils.py:179-201)."""

    def __init__(self, dyn):
        self._dyn = dyn
        self.theta, self.mu = dyn.model.ou_theta, np.zeros(4)
        self.sigma = dyn.model.ou_sigma

    @property
    def state(self):
        if self._dyn._vec.layout.row_ou < 0:
            return np.zeros(4)
        return self._dyn._get("ou_state")

    def reset(self):
        if self._dyn._vec.layout.row_ou >= 0:
            self._dyn._vec.set_field("ou_state", [0., 0., 0., 0.])


class QuadrotorDynamics(object):
    """quadrotor.py:60-541.  `step(thrust_cmds, dt)` runs `dynamics_steps_num` substeps of step1 on the device."""

    def __init__(self, model_params, room_box=None, dynamics_steps_num=1, dim_mode="3D", gravity=GRAV,
                 dynamics_simplification=False, device="cuda:0", dtype="float64", seed=0, num_envs=1,
                 noise_mode="philox", _vec=None, _vec_kwargs=None):
        if dim_mode not in ("1D", "2D", "3D"):
            raise ValueError('QuadEnv: Unknown dimensionality mode %s' % dim_mode)            # :131
        self.dynamics_steps_num = dynamics_steps_num
        self.dynamics_simplification = dynamics_simplification
        self.dim_mode = dim_mode
        self.control_mx = {"1D": np.ones([4, 1]), "2D": np.array([[1., 0.], [1., 0.], [0., 1.], [0., 1.]]),
                           "3D": np.eye(4)}[dim_mode]                                          # :121-129
        self.vxyz_max, self.acc_max = 3., 3. * GRAV                                            # :92-94
        self.gravity = gravity
        self.since_last_svd_limit = 0.5
        self.room_box = (np.array([[-10., -10., 0.], [10., 10., 10.]]) if room_box is None
                         else np.array(room_box, dtype=np.float64).copy())                    # :116-119
        self.num_envs = int(num_envs)
        self._device, self._dtype, self._seed = device, dtype, seed
        self._handles = {}
        self._vec = None
        self._view = _vec is not None
        self._fresh = True
        self.update_model(model_params)
        if _vec is not None:                       # a view on the env's drones: same state tensor, same handle
            self._vec_kwargs = dict(_vec_kwargs)
            self._vec = _vec
            self._handles[(round(1.0 / _vec.dt, 9), bool(_vec.cfg.zero_middle))] = _vec
        else:
            self._vec_kwargs = dict(dynamics_params=self.model_params, controller="raw", dim_mode=dim_mode,
                                    sim_steps=dynamics_steps_num, room_size=float(self.room_box[1][2]), gravity=gravity,
                                    device=device, dtype=dtype, seed=seed, auto_reset=False, info=True,
                                    noise_mode=noise_mode, goals=np.zeros((self.num_envs, 3)),
                                    host_mapped=self.num_envs <= 64, dynamics_simplification=dynamics_simplification)
            self._handle(1.0 / 200.0, True)        # allocates the state tensor; re-keyed by the dt of the first step
            self.set_state(np.zeros(3), np.zeros(3), np.eye(3), np.zeros(3))

    # ---- model (quadrotor.py:142-208) --------------------------------------------------------------------
    def update_model(self, model_params):
        self.model_params = model_params
        m = ModelConstants(model_params, gravity=self.gravity, dynamics_simplification=self.dynamics_simplification)
        self.model = m                                           # carries link geometry + derived constants
        for k in ("mass", "inertia", "thrust_to_weight", "torque_to_thrust", "motor_linearity", "C_rot_drag",
                  "C_rot_roll", "motor_damp_time_up", "motor_damp_time_down", "thrust_noise_ratio", "vel_damp",
                  "damp_omega_quadratic", "motor_assymetry", "thrust_max", "torque_max", "prop_pos", "prop_ccw",
                  "prop_crossproducts", "arm", "omega_max", "G_omega", "torque_to_inertia"):
            setattr(self, k, getattr(m, k))
        if self._handles:                          # live handles hold the old constants: rebuild over the same drones
            if self._view:
                raise RuntimeError("this QuadrotorDynamics is a view on a QuadrotorEnv: call env.update_dynamics(params)")
            old, clock, key = self._vec.state.clone(), self._vec.clock, self._last_key
            self.close()
            self._vec = None
            self._vec_kwargs["dynamics_params"] = model_params
            h = self._handle(1.0 / key[0], key[1])
            if h.state.shape == old.shape:         # same optional rows (OU on/off unchanged): the drones keep their state
                h.state.copy_(old)
            h.clock = clock

    # ---- device plumbing ---------------------------------------------------------------------------------
    def _handle(self, dt, zero_middle):
        key = (round(1.0 / dt, 9), bool(zero_middle))
        h = self._handles.get(key)
        if h is None:
            kw = dict(self._vec_kwargs)
            kw.update(sim_freq=1.0 / dt, raw_control_zero_middle=bool(zero_middle))
            state = None
            for other in self._handles.values():
                state = other.state
                break
            h = QuadVecEnv(self.num_envs, state=state, **kw)
            if state is not None:
                h.set_field("goals", self._vec.get_field("goals"))
            self._handles[key] = h
        if self._vec is not None and h is not self._vec:
            h.clock = self._vec.clock              # the re-orthonormalisation timer is one global phase (:104, :381-386)
        self._vec = h
        self._last_key = key
        return h

    @property
    def thrust_noise(self):
        return _OUView(self)

    def _get(self, name):
        vec = self._vec
        if vec.host_mapped and self.num_envs == 1:           # state is host memory: read the words in place
            first, count = vec._fields()[name]
            if first < 0:
                raise AttributeError("state has no %s row in this configuration" % name)
            w = vec.state.numpy()[:, 0, :].reshape(-1)[first:first + count].astype(np.float64)
            return w.reshape(3, 3) if name == "rot" else w
        v = vec.get_field(name).double().cpu().numpy()
        return v[0] if self.num_envs == 1 else v

    def _put(self, name, value):
        vec = self._vec
        if vec.host_mapped and self.num_envs == 1:
            first, count = vec._fields()[name]
            st = vec.state.numpy()
            v = np.asarray(value, dtype=np.float64).reshape(-1)
            for k in range(count):
                st[(first + k) // 4, 0, (first + k) % 4] = v[k]
        else:
            vec.set_field(name, np.asarray(value, dtype=np.float64).reshape(-1, vec._fields()[name][1]))

    def _info(self, first, count):
        if self._vec.host_mapped:
            v = self._vec.info.numpy()[:, first:first + count].astype(np.float64)
        else:
            v = self._vec.info[:, first:first + count].double().cpu().numpy()
        return v[0] if self.num_envs == 1 else v

    def _controller_step(self, name, action, dt, goal=None, zero_middle=True):
        """controller mixing + dynamics_steps_num x step1 (what every reference controller's step() does)."""
        h = self._handle(dt, zero_middle)
        if goal is not None:
            self._put("goals", goal)
        a = None if action is None else np.asarray(action, dtype=np.float64).reshape(self.num_envs, -1)
        h.dynamics_step(a, controller=name, n_substeps=self.dynamics_steps_num)
        self._fresh = False
        return self._info(_lib.QS_INFO_CMDS, 4)

    # ---- state (quadrotor.py:213-224, :438-440) ----------------------------------------------------------
    def set_state(self, position, velocity, rotation, omega, thrusts=np.zeros((4,))):
        position, velocity, rotation, omega = (np.asarray(x, dtype=np.float64) for x in (position, velocity, rotation, omega))
        if self.num_envs == 1:
            for v in (position, velocity, omega):
                assert v.shape == (3,)
            assert rotation.shape == (3, 3)
        self._put("pos", position)
        self._put("vel", velocity)
        self._put("rot", rotation)
        self._put("omega", omega.astype(np.float32).astype(np.float64))                            # :223
        self.thrusts = copy.deepcopy(thrusts)
        self._fresh = True

    def reset(self):
        self._put("thrust_cmds_damp", np.zeros(4 * self.num_envs))
        self._put("thrust_rot_damp", np.zeros(4 * self.num_envs))

    pos = property(lambda self: self._get("pos"), lambda self, v: self._put("pos", v))
    vel = property(lambda self: self._get("vel"), lambda self, v: self._put("vel", v))
    omega = property(lambda self: self._get("omega"), lambda self, v: self._put("omega", v))
    rot = property(lambda self: self._get("rot"), lambda self, v: self._put("rot", v))
    thrust_cmds_damp = property(lambda self: self._get("thrust_cmds_damp"))
    thrust_rot_damp = property(lambda self: self._get("thrust_rot_damp"))

    @property
    def acc(self):
        return np.zeros(3) if self._fresh else self._info(_lib.QS_INFO_ACC, 3)

    @property
    def accelerometer(self):
        return np.array([0., 0., GRAV]) if self._fresh else self._info(_lib.QS_INFO_ACCELEROMETER, 3)

    @property
    def torque(self):
        return self._info(_lib.QS_INFO_TORQUE, 3)

    @property
    def omega_dot(self):
        return self._info(_lib.QS_INFO_OMEGA_DOT, 3)

    @property
    def since_last_svd(self):
        return self._vec.clock[0]

    @since_last_svd.setter
    def since_last_svd(self, t):
        self._vec.clock = (float(t), self._vec.clock[1])

    # ---- stepping ----------------------------------------------------------------------------------------
    def step(self, thrust_cmds, dt):
        """quadrotor.py:261-262: [step1(thrust_cmds, dt) for _ in range(dynamics_steps_num)]."""
        h = self._handle(dt, bool(self._vec.cfg.zero_middle))
        h.dynamics_step(np.asarray(thrust_cmds, dtype=np.float64).reshape(self.num_envs, 4), controller=None,
                        n_substeps=self.dynamics_steps_num)
        self._fresh = False

    def step1(self, thrust_cmds, dt):
        """quadrotor.py:273-436, one substep."""
        h = self._handle(dt, bool(self._vec.cfg.zero_middle))
        h.dynamics_step(np.asarray(thrust_cmds, dtype=np.float64).reshape(self.num_envs, 4), controller=None, n_substeps=1)
        self._fresh = False

    def random_state(self, box, vel_max=15.0, omega_max=2 * np.pi):
        """quadrotor.py:227-239, drawn by the device's spawn sampler (Philox): returns (pos, vel, rot, omega) without
        touching this object's state.  pos ~ U(-box, box)^3, vel/omega uniform direction-ish with U(0, max) magnitude,
        rot uniform on SO(3) (quad_utils.py:47-57)."""
        tmp = QuadVecEnv(self.num_envs, dynamics_params=self.model_params, init_random_state=True, goal=(0., 0., 0.),
                         box=float(box), device=self._device, dtype=self._dtype, auto_reset=False,
                         seed=int(np.random.randint(0, 2 ** 31 - 1)), spawn_z_min=-1e30, max_init_vel=vel_max,
                         max_init_omega=omega_max)
        tmp.reset()
        out = tuple(tmp.get_field(k).double().cpu().numpy() for k in ("pos", "vel", "rot", "omega"))
        tmp.close()
        return tuple(v[0] for v in out) if self.num_envs == 1 else out

    def pitch_roll_restricted_random_state(self, box, vel_max=15.0, omega_max=2 * np.pi, pitch_max=0.5, roll_max=0.5,
                                           yaw_max=3.14):
        """quadrotor.py:243-259, drawn by the device's spawn sampler like random_state: attitude from Euler angles
        U(-pitch_max, pitch_max), U(-roll_max, roll_max), U(-yaw_max, yaw_max), R = Rz(yaw) Ry(pitch) Rx(roll)."""
        tmp = QuadVecEnv(self.num_envs, dynamics_params=self.model_params, init_random_state="pitch_roll_restricted",
                         goal=(0., 0., 0.), box=float(box), device=self._device, dtype=self._dtype, auto_reset=False,
                         seed=int(np.random.randint(0, 2 ** 31 - 1)), spawn_z_min=-1e30, max_init_vel=vel_max,
                         max_init_omega=omega_max, pitch_max=pitch_max, roll_max=roll_max, yaw_max=yaw_max)
        tmp.reset()
        out = tuple(tmp.get_field(k).double().cpu().numpy() for k in ("pos", "vel", "rot", "omega"))
        tmp.close()
        return tuple(v[0] for v in out) if self.num_envs == 1 else out

    # ---- affine representation s' = dt (F(s) + G(s) u)  (quadrotor.py:457-518) --------------------------
    def F(self, s, dt):
        """Unforced dynamics F(s, dt) of one state vector s[21] = xyz, Vxyz, R, omega, goal (:461-497), or of a batch [n, 21]."""
        s = np.asarray(s, dtype=np.float64)
        F, _ = self._vec.affine_fg(s.reshape(-1, 21), dt, want_G=False)
        out = F.double().cpu().numpy()
        return out[0] if s.ndim == 1 else out

    def G(self, s):
        """Forced dynamics G(s) [21, C] (:501-518), C = columns of control_mx for this dim_mode; a batch gives [n, 21, C]."""
        s = np.asarray(s, dtype=np.float64)
        _, G = self._vec.affine_fg(s.reshape(-1, 21), 1.0, want_F=False)
        out = G.double().cpu().numpy()
        return out[0] if s.ndim == 1 else out

    def state_vector(self):                                                                  # :520-525
        return np.concatenate([self.pos, self.vel, self.rot.flatten(), self.omega])

    def action_space(self):                                                                  # :536-540
        return Box(np.zeros(4), np.ones(4), dtype=np.float32)

    @staticmethod
    def angvel2thrust(w, linearity=0.424):                                                   # :133-140
        return (1 - linearity) * w ** 2 + linearity * w

    def look_at(self):                                                                       # :527-534
        """eye, center, up of the onboard camera (45 degrees down), for gluLookAt-style renderers."""
        R = self.rot
        eye = self.pos + R @ np.array([0., 0., -0.02])
        theta = np.radians(45.0)
        to = np.cos(theta) * R[:, 0] - np.sin(theta) * R[:, 2]
        n = np.linalg.norm(to)
        to = to / n if n >= 1e-5 else to
        return eye, eye + to, np.cross(to, R[:, 1])

    def close(self):
        for h in self._handles.values():
            h.close()
        self._handles = {}


# obs component table (quadrotor.py:900-914)
def _obs_bounds(env):
    d, rb = env.dynamics, env.room_box
    return {
        "xyz": [-(rb[1] - rb[0]), rb[1] - rb[0]], "xyzr": [-(rb[1] - rb[0]), rb[1] - rb[0]],
        "vxyz": [-d.vxyz_max * np.ones(3), d.vxyz_max * np.ones(3)], "vxyzr": [-d.vxyz_max * np.ones(3), d.vxyz_max * np.ones(3)],
        "acc": [-d.acc_max * np.ones(3), d.acc_max * np.ones(3)], "R": [-np.ones(9), np.ones(9)],
        "omega": [-d.omega_max * np.ones(3), d.omega_max * np.ones(3)], "t2w": [0. * np.ones(1), 5. * np.ones(1)],
        "t2t": [0. * np.ones(1), 1. * np.ones(1)], "h": [0. * np.ones(1), rb[1][2] * np.ones(1)],
        "act": [np.zeros(4), np.ones(4)], "quat": [-np.ones(4), np.ones(4)], "euler": [-np.pi * np.ones(3), np.pi * np.ones(3)],
    }


REW_NAMES = ("pos", "action", "crash", "orient", "yaw", "rot", "attitude", "spin", "act_change", "vel")   # :607-631


class _Spec(object):
    def __init__(self, id, max_episode_steps):
        self.id, self.max_episode_steps = id, max_episode_steps


class QuadrotorEnv(object):
    """quadrotor.py:647-1156: old-gym API, `step(action) -> (obs, reward, done, info)`, `reset() -> obs`.

    Extra keyword-only arguments (not in the reference): device, dtype ("float64" default), seed, noise_mode
    ("philox" | "injected" - the parity tests replay recorded reference noise through `_vec.set_noise`).
    """
    metadata = {'render.modes': ['human', 'rgb_array'], 'video.frames_per_second': 50}
    spec = None
    _multi_reward = False
    _sampler_key = "class"

    def __init__(self, dynamics_params="DefaultQuad", dynamics_change=None, dynamics_randomize_every=None,
                 dyn_sampler_1=None, dyn_sampler_2=None, raw_control=True, raw_control_zero_middle=True, dim_mode='3D',
                 tf_control=False, sim_freq=200., sim_steps=2, obs_repr="xyz_vxyz_R_omega", ep_time=7, obstacles_num=0,
                 room_size=10, init_random_state=False, rew_coeff=None, sense_noise=None, verbose=False, gravity=GRAV,
                 resample_goal=False, t2w_std=0.005, t2t_std=0.0005, excite=False, dynamics_simplification=False,
                 device="cuda:0", dtype="float64", seed=None, noise_mode="philox"):
        if obstacles_num:
            raise NotImplementedError("obstacles are not part of the env-step path (dead code in the reference, SURVEY.md section 2)")
        self.init_random_state, self.room_size, self.obs_repr = init_random_state, room_size, obs_repr
        self.sim_steps, self.dim_mode, self.raw_control_zero_middle = sim_steps, dim_mode, raw_control_zero_middle
        self.tf_control, self.dynamics_randomize_every, self.verbose = False, dynamics_randomize_every, verbose
        self.obstacles_num, self.raw_control, self.scene, self.obstacles = 0, raw_control, None, None
        self.gravity, self.resample_goal, self.excite = gravity, resample_goal, excite
        self.t2w_std, self.t2w_min, self.t2w_max = t2w_std, 1.5, 10.0
        self.t2t_std, self.t2t_min, self.t2t_max = t2t_std, 0.005, 1.0
        self.dynamics_simplification = dynamics_simplification
        self.max_init_vel, self.max_init_omega = 1., 2 * np.pi                                # :717-718
        self.room_box = np.array([[-room_size, -room_size, 0], [room_size, room_size, room_size]], dtype=np.float64)
        self.box, self.box_scale = 2.0, 1.0                                                   # :728-729
        self.traj_count = 0
        self._device, self._dtype, self._noise_mode = device, dtype, noise_mode
        self._sim_freq, self._rew_coeff_arg = float(sim_freq), rew_coeff
        self._vec, self.dynamics = None, None
        self.update_sense_noise(sense_noise)
        self._seed(seed)
        # ---- dynamics (and randomisation) :738-763 ----
        self.dyn_base_sampler = getattr(quad_rand, dynamics_params)() if isinstance(dynamics_params, str) \
            else quad_rand.ConstValueSampler(None, {})
        self._dyn_params_arg = dynamics_params
        self.dynamics_change = copy.deepcopy(dynamics_change)
        base = self._sample_base()
        self.dyn_sampler_1, self.dyn_sampler_2 = (self._make_sampler(s, base) for s in (dyn_sampler_1, dyn_sampler_2))
        # ---- episode params :787-795 ----
        self.ep_time = ep_time
        self.dt = 1.0 / sim_freq
        self.metadata = dict(self.metadata)
        self.metadata["video.frames_per_second"] = sim_freq / self.sim_steps
        self.ep_len = int(self.ep_time / (self.dt * self.sim_steps))
        self.tick, self.crashed = 0, False
        self.control_freq = sim_freq / sim_steps
        self.goal = np.array([0., 0., 2.])
        self.actions = [np.zeros(4), np.zeros(4)]
        self.resample_dynamics()
        self.observation_space = self.make_observation_space()
        self.rew_coeff = dict(self._vec.rew_coeff)
        self._reset()
        if self.spec is None:
            self.spec = _Spec('Quadrotor-v0', self.ep_len)

    # ---- configuration -----------------------------------------------------------------------------------
    def _sample_base(self):
        if isinstance(self._dyn_params_arg, str):
            params = self.dyn_base_sampler.sample()
        else:
            params = copy.deepcopy(self._dyn_params_arg)
        if self.dynamics_change is not None:
            quad_rand._update_existing(params, self.dynamics_change)
        return params

    def _make_sampler(self, spec, params):
        if spec is None:
            return None
        kw = copy.deepcopy(spec)
        cls = kw.pop(self._sampler_key)
        return getattr(quad_rand, cls)(params=params, **kw)

    def update_sense_noise(self, sense_noise):                                                # :838-849
        from ..vec_env import resolve_sense_noise
        self._sense_noise_arg = sense_noise
        self.sense_noise = resolve_sense_noise(sense_noise)

    def _seed(self, seed=None):                                                               # :938-940
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 63))
        self.np_random = np.random.RandomState(seed % (2 ** 32))
        self._philox_seed = int(seed)
        return [seed]

    seed = _seed

    def resample_dynamics(self):                                                              # :1030-1056
        """MUST be followed by reset() (same warning as the reference)."""
        self.dynamics_params = self._sample_base()
        for s in (self.dyn_sampler_1, self.dyn_sampler_2):
            if s is not None:
                self.dynamics_params = s.sample(self.dynamics_params)
        quad_rand.check_quad_param_limits(self.dynamics_params)
        self.update_dynamics(self.dynamics_params)

    def update_dynamics(self, dynamics_params):                                               # :852-891
        self.dynamics_params = dynamics_params
        if self.raw_control:
            controller = {"1D": "vertical", "2D": "vertplane", "3D": "raw"}.get(self.dim_mode)
            if controller is None:
                raise ValueError('QuadEnv: Unknown dimensionality mode %s' % self.dim_mode)
        else:
            controller = "mellinger"
        old_state, old_clock = (None, None) if self._vec is None else (self._vec.state, self._vec.clock)
        kw = dict(dynamics_params=dynamics_params, controller=controller,
                  raw_control_zero_middle=self.raw_control_zero_middle, dim_mode=self.dim_mode, sim_freq=self._sim_freq,
                  sim_steps=self.sim_steps, obs_repr=self.obs_repr, ep_time=self.ep_time, room_size=self.room_size,
                  init_random_state=self.init_random_state, rew_coeff=self._rew_coeff_arg, sense_noise=self._sense_noise_arg,
                  gravity=self.gravity, device=self._device, dtype=self._dtype, seed=self._philox_seed,
                  noise_mode=self._noise_mode, auto_reset=False, info=True, multi_reward=self._multi_reward,
                  goals=np.asarray(self.goal, dtype=np.float64).reshape(1, 3), box=self.box, host_mapped=True,
                  resample_goal=self.resample_goal, excite=self.excite, dynamics_simplification=self.dynamics_simplification,
                  t2w_std=self.t2w_std, t2t_std=self.t2t_std)
        vec = QuadVecEnv(1, **kw)
        if old_state is not None:
            if old_state.shape == vec.state.shape:      # OU state, gyro bias and the SVD timer survive a model change
                vec.state.copy_(old_state)
            vec.clock = old_clock
            self.dynamics.close()
            self._vec.close()
        self._vec = vec
        self.dynamics = QuadrotorDynamics(model_params=dynamics_params, dynamics_steps_num=self.sim_steps,
                                          room_box=self.room_box, dim_mode=self.dim_mode, gravity=self.gravity,
                                          dynamics_simplification=self.dynamics_simplification, device=self._device,
                                          dtype=self._dtype, _vec=vec, _vec_kwargs=kw)
        zm = self.raw_control_zero_middle
        if self.raw_control:
            cls = {"1D": VerticalControl, "2D": VertPlaneControl, "3D": RawControl}[self.dim_mode]
            self.controller = cls(self.dynamics, zero_action_middle=zm)
        else:
            self.controller = NonlinearPositionController(self.dynamics, tf_control=False)
        self.action_space = self.controller.action_space(self.dynamics)

    def make_observation_space(self):                                                         # :898-936
        self.wall_offset = 0.3
        self.obs_space_low_high = _obs_bounds(self)
        self.obs_comp_names = list(self.obs_space_low_high.keys())
        self.obs_comp_sizes = [self.obs_space_low_high[n][1].size for n in self.obs_comp_names]
        comps = self.obs_repr.split("_")
        low = np.concatenate([self.obs_space_low_high[c][0] for c in comps])
        high = np.concatenate([self.obs_space_low_high[c][1] for c in comps])
        self.obs_comp_sizes_dict, self.obs_space_comp_indx, self.obs_comp_end, end = {}, {}, [], 0
        for i, n in enumerate(self.obs_comp_names):
            end += self.obs_comp_sizes[i]
            self.obs_comp_sizes_dict[n], self.obs_space_comp_indx[n] = self.obs_comp_sizes[i], i
            self.obs_comp_end.append(end)
        self.observation_space = Box(low, high, dtype=np.float32)
        return self.observation_space

    def save_dyn_params(self, filename):                                                      # :829-836
        import yaml

        def conv(x):
            return {k: conv(v) for k, v in x.items()} if isinstance(x, dict) else str(x)

        with open(filename, "w") as f:
            f.write(yaml.dump(conv(self.dynamics_params), default_flow_style=False))

    # ---- the env API ---------------------------------------------------------------------------------------
    def _np(self, t):
        return t.numpy().astype(np.float64) if t.device.type == "cpu" else t.double().cpu().numpy()

    def state_vector(self, env=None):
        """Observation of the current state (get_state.state_<obs_repr>): the last vector the device produced."""
        return self._np(self._vec.obs)[0].copy()

    def _step(self, action):                                                                  # :942-1028
        action = np.asarray(action, dtype=np.float64)
        self.actions[1] = copy.deepcopy(self.actions[0])
        self.actions[0] = copy.deepcopy(action)
        # (excite, :957-962: the kernel draws the new goal every 5th tick itself)
        obs, rew, done = self._vec.step(action.reshape(1, -1))         # host-mapped: results are host memory, already synced
        if not np.isfinite(float(rew[0])):                                                    # :633-636
            raise ValueError('QuadEnv: reward is Nan')
        self.dynamics._fresh = False
        if self.excite:
            self.goal = self.dynamics._get("goals")
        info_row = self._np(self._vec.info)[0]
        self.crashed = bool(info_row[_lib.QS_INFO_CRASHED])
        self.controller.action = info_row[_lib.QS_INFO_CMDS:_lib.QS_INFO_CMDS + 4].copy()
        self.tick = int(info_row[_lib.QS_INFO_TICK])
        self.time_remain = self.ep_len - (self.tick - 1)
        reward, done = float(self._np(rew)[0]), bool(done[0].item())
        sv = self._np(obs)[0].copy()
        self.traj_count += int(done)
        w, raw = info_row[_lib.QS_INFO_COSTS:_lib.QS_INFO_COSTS + 10], info_row[_lib.QS_INFO_COSTS_RAW:_lib.QS_INFO_COSTS_RAW + 10]
        rew_info = {"rew_main": -w[0], "rewraw_main": -raw[0]}
        for k, name in enumerate(REW_NAMES):
            rew_info["rew_" + name] = -w[k]
            rew_info["rewraw_" + name] = -raw[k]
        d = self.dynamics
        cd = d.thrust_cmds_damp
        obs_comp = {"xyz": [d.pos], "vxyz": [d.vel], "acc": [d.accelerometer], "omega": [d.omega], "omega_dot": [d.omega_dot],
                    "R": [d.rot.flatten()], "act": [action], "act_clipped": [np.clip(self.controller.action, 0., 1.)],
                    "act_filtered": [cd], "act_torque": [d.prop_ccw * cd], "torque": [d.torque]}
        dyn_params = {"mass": [d.mass], "motor_linearity": [d.motor_linearity], "motor_time_up": [d.motor_damp_time_up],
                      "motor_time_down": [d.motor_damp_time_down], "motor_assymetry": [d.motor_assymetry],
                      "motor_pos": [d.prop_pos.flatten()], "motor_ccw": [d.prop_ccw], "t2w": [d.thrust_to_weight],
                      "t2t": [d.torque_to_thrust], "t2i": [d.torque_to_inertia], "inertia": [d.inertia],
                      "thrust_max": [np.mean(d.thrust_max)], "torque_max": [np.mean(d.torque_max)], "arm": [d.arm],
                      "grav": [GRAV], "dt": [self.dt * self.sim_steps]}
        return sv, reward, done, {'rewards': rew_info, "obs_comp": obs_comp, "dyn_params": dyn_params}

    def _reset(self):                                                                         # :1059-1144
        if self.dynamics_randomize_every is not None and (self.traj_count + 1) % self.dynamics_randomize_every == 0:
            self.resample_dynamics()
        if not (self.resample_goal or self.excite):                                           # :1078-1081
            self.goal = np.array([0., 0., 2.])
            self.dynamics._put("goals", self.goal)
        obs = self._vec.reset()                      # resample_goal / excite: the reset kernel re-makes the goal itself
        if self.resample_goal or self.excite:
            self.goal = self.dynamics._get("goals")
        self.dynamics._fresh = True
        self.crashed, self.tick = False, 0
        self.actions = [np.zeros(4), np.zeros(4)]
        return self._np(obs)[0].copy()

    def reset(self):
        return self._reset()

    def step(self, action):
        return self._step(action)

    def render(self, mode='human', close=False):
        raise NotImplementedError("rendering (pyglet scene, quadrotor_visualization.py) is outside the env-step path")

    def close(self):
        if self.dynamics is not None:
            self.dynamics.close()
        if self._vec is not None:
            self._vec.close()
