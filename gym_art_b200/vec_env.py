"""QuadVecEnv: N quadrotor envs stepped by one fused sm_100a kernel per env step.

The batched counterpart of the reference's QuadrotorEnv (gym_art/quadrotor/quadrotor.py:647-1156): same ctor
vocabulary, same per-env semantics (see DESIGN.md "behaviour checklist"), torch tensors in and out, state owned
by torch in the structure-of-arrays layout of include/quadsim.h.  All arithmetic happens in libquadsim.so; this
file only builds the QsConfig, owns the tensors and forwards pointers.  There is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib
from .quadrotor import quadrotor_randomization as quad_rand
from .quadrotor.model import GRAV, ModelConstants
from .quadrotor.sensor_noise import SensorNoise

CONTROLLERS = {"raw": _lib.QS_CTRL_RAW, "vertical": _lib.QS_CTRL_VERTICAL, "vertplane": _lib.QS_CTRL_VERTPLANE,
               "shifted": _lib.QS_CTRL_SHIFTED, "omega_thrust": _lib.QS_CTRL_OMEGA_THRUST,
               "velocity_yaw": _lib.QS_CTRL_VELOCITY_YAW, "mellinger": _lib.QS_CTRL_MELLINGER}

REW_COEFF_DEFAULT = {"pos": 1., "effort": 0.05, "action_change": 0., "crash": 1., "orient": 1., "yaw": 0., "rot": 0.,
                     "attitude": 0., "spin": 0.1, "vel": 0.}                          # quadrotor.py:799-806
REW_COEFF_MULTI_DEFAULT = {"pos": 1., "pos_offset": 0.1, "pos_log_weight": 1., "pos_linear_weight": 0.1,
                           "effort": 0.01, "action_change": 0., "crash": 1., "orient": 1., "yaw": 0., "rot": 0.,
                           "attitude": 0., "spin": 0., "vel": 0.}                     # quadrotor_multi.py:810-818


def resolve_dynamics_params(dynamics_params, dynamics_change=None, check=True):
    # (the limits check assumes the full link geometry; simplified-geometry dicts skip the parts they do not have)
    """quadrotor.py:738-744 + :1030-1056: sampler name (or dict) -> checked parameter dict."""
    if isinstance(dynamics_params, str):
        params = getattr(quad_rand, dynamics_params)().sample()
    else:
        import copy
        params = copy.deepcopy(dynamics_params)
    if dynamics_change is not None:
        quad_rand._update_existing(params, dynamics_change)
    return quad_rand.check_quad_param_limits(params) if check else params


def resolve_sense_noise(sense_noise):
    """quadrotor.py:838-849."""
    if isinstance(sense_noise, SensorNoise):
        return sense_noise
    if isinstance(sense_noise, dict):
        return SensorNoise(**sense_noise)
    if isinstance(sense_noise, str):
        if sense_noise == "default":
            return SensorNoise(bypass=False)
        raise ValueError("ERROR: QuadEnv: sense_noise parameter is of unknown type: " + str(sense_noise))
    if sense_noise is None:
        return SensorNoise(bypass=True)
    raise ValueError("ERROR: QuadEnv: sense_noise parameter is of unknown type: " + str(sense_noise))


def resolve_rew_coeff(rew_coeff, multi_reward):
    """quadrotor.py:799-818."""
    base = dict(REW_COEFF_MULTI_DEFAULT if multi_reward else REW_COEFF_DEFAULT)
    if rew_coeff is not None:
        assert isinstance(rew_coeff, dict)
        assert set(rew_coeff.keys()).issubset(set(base.keys()) | {"collision"}), "unknown reward coefficient"
        base.update(rew_coeff)
    return {k: float(v) for k, v in base.items()}


def make_config(num_envs, model, controller="raw", raw_control_zero_middle=True, dim_mode="3D", sim_freq=200.,
                sim_steps=2, obs_repr="xyz_vxyz_R_omega", ep_time=7, room_size=10, init_random_state=False,
                rew_coeff=None, sense_noise=None, device_index=0, dtype="float32", seed=0, noise_mode="philox",
                auto_reset=True, n_agents=1, per_env_goal=False, env_id_offset=0, multi_reward=False, collision_k=1.0,
                goal=(0., 0., 2.), box=2.0, spawn_z_min=0.25, max_init_vel=1., max_init_omega=2 * np.pi, formation=None,
                per_env_model=False, resample_goal=False, excite=False, t2w_std=0.005, t2t_std=0.0005, pitch_max=0.5,
                roll_max=0.5, yaw_max=3.14):
    """Pure host function (no GPU): ctor arguments -> QsConfig.  Raises the reference's ValueErrors."""
    cfg = _lib.QsConfig()
    cfg.abi_version = _lib.QS_ABI_VERSION
    cfg.struct_size = C.sizeof(_lib.QsConfig)
    cfg.n_envs = int(num_envs)
    cfg.env_id_offset = int(env_id_offset)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cfg.n_agents = int(n_agents)
    cfg.device = int(device_index)
    cfg.dtype = {"float32": _lib.QS_F32, "float64": _lib.QS_F64}[dtype]
    if controller not in CONTROLLERS:
        raise ValueError("QuadEnv: unknown controller %s" % controller)
    cfg.controller = CONTROLLERS[controller]
    cfg.zero_middle = int(bool(raw_control_zero_middle))
    if obs_repr not in _lib.OBS_REPRS:
        raise ValueError("QuadEnv: obs_repr %s is not one of the working representations %s" % (obs_repr, _lib.OBS_REPRS))
    cfg.obs_repr = _lib.OBS_REPRS.index(obs_repr)
    cfg.noise_mode = {"philox": _lib.QS_NOISE_PHILOX, "injected": _lib.QS_NOISE_INJECTED}[noise_mode]
    # False / True (quadrotor.py:1100-1115), or "pitch_roll_restricted": the law of QuadrotorDynamics.pitch_roll_restricted_random_state
    # (:243-259), which the reference env never selects itself
    cfg.init_random_state = (_lib.QS_INIT_PITCH_ROLL_RESTRICTED if init_random_state == "pitch_roll_restricted"
                             else (_lib.QS_INIT_RANDOM if init_random_state else _lib.QS_INIT_DEFAULT))
    cfg.pitch_max, cfg.roll_max, cfg.yaw_max = float(pitch_max), float(roll_max), float(yaw_max)       # quadrotor.py:243
    cfg.t2w_std, cfg.t2t_std = float(t2w_std), float(t2t_std)                                         # quadrotor.py:658
    cfg.auto_reset = int(bool(auto_reset))
    if dim_mode not in ("1D", "2D", "3D"):
        raise ValueError('QuadEnv: Unknown dimensionality mode %s' % dim_mode)       # quadrotor.py:131
    cfg.dim_mode = int(dim_mode[0])
    cfg.sim_steps = int(sim_steps)
    dt = 1.0 / sim_freq                                                               # quadrotor.py:790
    cfg.ep_len = int(ep_time / (dt * sim_steps))                                      # quadrotor.py:792
    cfg.per_env_goal = int(bool(per_env_goal) or bool(resample_goal) or bool(excite))
    cfg.resample_goal, cfg.excite = int(bool(resample_goal)), int(bool(excite))         # quadrotor.py:1078-1081, :957-962
    cfg.dt = dt
    cfg.svd_period = 0.5                                                              # quadrotor.py:105
    for k in range(3):
        cfg.goal[k] = float(goal[k])
    cfg.spawn_box = float(box)                                                        # quadrotor.py:728
    cfg.spawn_z_min = float(spawn_z_min)                                              # quadrotor.py:1094
    lo = [-room_size, -room_size, 0.]
    hi = [room_size, room_size, room_size]                                            # quadrotor.py:723
    for k in range(3):
        cfg.room_box[0][k] = float(lo[k])
        cfg.room_box[1][k] = float(hi[k])
    cfg.max_init_vel = float(max_init_vel)                                            # quadrotor.py:717
    cfg.max_init_omega = float(max_init_omega)                                        # quadrotor.py:718
    cfg.collision_k = float(collision_k)
    cfg.per_env_model = int(bool(per_env_model))
    if formation is not None:                                                         # swarm: goal of drone a = formation[a]
        f = np.asarray(formation, dtype=np.float64).reshape(-1, 3)
        if f.shape[0] != int(n_agents) or per_env_goal:
            raise ValueError("formation must have one goal per agent and excludes per-env goals")
        cfg.use_formation = 1
        for a in range(f.shape[0]):
            for k in range(3):
                cfg.formation[a][k] = float(f[a, k])
    cfg.model = model.to_qs()
    rc = resolve_rew_coeff(rew_coeff, multi_reward)
    for k in ("pos", "effort", "action_change", "crash", "orient", "yaw", "rot", "attitude", "spin", "vel"):
        setattr(cfg.reward, k, rc[k])
    cfg.reward.log_pos = int(bool(multi_reward))
    if multi_reward:
        cfg.reward.pos_offset, cfg.reward.pos_log_weight = rc["pos_offset"], rc["pos_log_weight"]
        cfg.reward.pos_linear_weight = rc["pos_linear_weight"]
    cfg.reward.collision = rc.get("collision", 0.0)
    sn = resolve_sense_noise(sense_noise)
    cfg.sense.enabled = int(not sn.bypass)
    for k in ("pos_norm_std", "pos_unif_range", "vel_norm_std", "vel_unif_range", "quat_norm_std", "quat_unif_range",
              "gyro_norm_std", "gyro_noise_density", "gyro_random_walk", "gyro_bias_correlation_time",
              "acc_static_noise_std", "acc_dynamic_noise_ratio"):
        setattr(cfg.sense, k, float(getattr(sn, k)))
    return cfg, rc, sn


class QuadVecEnv(object):
    """N independent quadrotor envs (or N / n_agents swarms) on one GPU.

    step(actions) -> (obs [N,O], reward [N], done [N] bool) as torch tensors on the device.  The returned
    tensors are reused by the next call.  With auto_reset (default) an env that finishes an episode is re-spawned
    inside the same kernel: the step returns the terminal reward/done and the first observation of the next
    episode (standard vec-env semantics; the reference leaves reset() to the caller, quadrotor.py:1260-1305).
    """

    def __init__(self, num_envs, dynamics_params="DefaultQuad", dynamics_change=None, controller=None,
                 raw_control=True, raw_control_zero_middle=True, dim_mode="3D", sim_freq=200., sim_steps=2,
                 obs_repr="xyz_vxyz_R_omega", ep_time=7, room_size=10, init_random_state=False, rew_coeff=None,
                 sense_noise=None, gravity=GRAV, device="cuda:0", dtype="float32", seed=0, noise_mode="philox",
                 auto_reset=True, n_agents=1, goals=None, env_id_offset=0, multi_reward=False, collision_k=1.0,
                 info=False, goal=(0., 0., 2.), box=2.0, state=None, spawn_z_min=0.25, max_init_vel=1.,
                 max_init_omega=2 * np.pi, formation=None, per_env_dynamics=None, dyn_sampler_1=None, dyn_sampler_2=None,
                 host_mapped=False, resample_goal=False, excite=False, dynamics_simplification=False, terminal_obs=False,
                 t2w_std=0.005, t2t_std=0.0005, pitch_max=0.5, roll_max=0.5, yaw_max=3.14):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("QuadVecEnv runs on a CUDA device only (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("QuadVecEnv runs on a CUDA device only")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.tdtype = {"float32": torch.float32, "float64": torch.float64}[dtype]
        if controller is None:                                                    # quadrotor.py:876-886
            if raw_control:
                controller = {"1D": "vertical", "2D": "vertplane", "3D": "raw"}.get(dim_mode, "raw")
            else:
                controller = "mellinger"
        self.controller_name = controller
        self.dynamics_params = resolve_dynamics_params(dynamics_params, dynamics_change)
        self.model = ModelConstants(self.dynamics_params, gravity=gravity, dynamics_simplification=dynamics_simplification)
        self.cfg, self.rew_coeff, self.sense_noise = make_config(
            num_envs, self.model, controller=controller, raw_control_zero_middle=raw_control_zero_middle,
            dim_mode=dim_mode, sim_freq=sim_freq, sim_steps=sim_steps, obs_repr=obs_repr, ep_time=ep_time,
            room_size=room_size, init_random_state=init_random_state, rew_coeff=rew_coeff, sense_noise=sense_noise,
            device_index=dev_index, dtype=dtype, seed=seed, noise_mode=noise_mode, auto_reset=auto_reset,
            n_agents=n_agents, per_env_goal=goals is not None, env_id_offset=env_id_offset, multi_reward=multi_reward,
            collision_k=collision_k, goal=goal, box=box, spawn_z_min=spawn_z_min, max_init_vel=max_init_vel,
            max_init_omega=max_init_omega, formation=formation, per_env_model=per_env_dynamics is not None,
            resample_goal=resample_goal, excite=excite, t2w_std=t2w_std, t2t_std=t2t_std, pitch_max=pitch_max,
            roll_max=roll_max, yaw_max=yaw_max)
        self.layout = _lib.state_layout(self.cfg)
        self.num_envs = int(num_envs)
        self.n_agents = int(n_agents)
        self.obs_dim, self.act_dim = self.layout.obs_dim, self.layout.act_dim
        self.dt, self.sim_steps, self.ep_len = self.cfg.dt, int(sim_steps), self.cfg.ep_len
        self.obs_repr = obs_repr
        self.goal = np.array([self.cfg.goal[0], self.cfg.goal[1], self.cfg.goal[2]])
        N = self.num_envs
        # host_mapped: every buffer is PINNED HOST memory that the kernels read and write across PCIe (the device
        # address of a pinned allocation equals its host address under unified addressing).  For a handful of envs -
        # the single-env facades - this removes every copy: one launch + one stream sync per call, results and state
        # are plain host memory.  Not for throughput work.
        self.host_mapped = bool(host_mapped)
        self.mem = torch.device("cpu") if self.host_mapped else self.device
        self._stage = {}
        with torch.cuda.device(self.device):
            if state is None:
                self.state = self._alloc((self.layout.rows, N, 4), self.tdtype)
            else:       # a second handle over the same drones (e.g. another controller or dt): same layout required
                assert tuple(state.shape) == (self.layout.rows, N, 4) and state.dtype == self.tdtype and state.is_contiguous()
                assert state.device.type == self.mem.type
                self.state = state
            self.obs = self._alloc((N, self.obs_dim), self.tdtype)
            self.rew = self._alloc((N,), self.tdtype)
            self._done_u8 = self._alloc((N,), torch.uint8)
            self.done = self._done_u8.view(torch.bool)      # zero-copy: the kernel writes 0 / 1
            self.info = self._alloc((N, _lib.QS_INFO_DIM), self.tdtype) if info else None
        if state is None:
            self.set_field("rot", torch.eye(3, dtype=self.tdtype, device=self.mem).reshape(1, 9))
        if goals is None and (resample_goal or excite) and state is None:
            self.set_field("goals", np.asarray(goal, dtype=np.float64).reshape(1, 3))
        if goals is not None and state is None:
            self.set_field("goals", np.asarray(goals, dtype=np.float64).reshape(N, 3))
        self._h = C.c_void_p()
        _lib.check(self.lib.qs_create(C.byref(self.cfg), C.byref(self._h)))
        _lib.check(self.lib.qs_bind_state(self._h, self.state.data_ptr()), self._h)
        if info:
            _lib.check(self.lib.qs_bind_info(self._h, self.info.data_ptr()), self._h)
        # auto-reset hides the terminal state behind the next episode's first observation; terminal_obs=True keeps the
        # observation the reference's step() returns on its done step (rows of envs with done=1 are valid)
        self.terminal_obs = None
        if terminal_obs:
            self.terminal_obs = self._alloc((N, self.obs_dim), self.tdtype)
            _lib.check(self.lib.qs_bind_terminal_obs(self._h, self.terminal_obs.data_ptr()), self._h)
        self._noise_refs = None
        # spaces of ONE env, as the reference's controller / env report them (quadrotor.py:890, :898-936)
        from .quadrotor.spaces import Box
        lo, hi = self._action_bounds()
        self.action_space = Box(lo, hi, dtype=np.float32)
        rb = np.array([[-room_size, -room_size, 0.], [room_size, room_size, room_size]], dtype=np.float64)
        comp = {"xyz": (-(rb[1] - rb[0]), rb[1] - rb[0]), "vxyz": (-3. * np.ones(3), 3. * np.ones(3)), "R": (-np.ones(9), np.ones(9)),
                "omega": (-40. * np.ones(3), 40. * np.ones(3)), "h": (np.zeros(1), rb[1][2] * np.ones(1)),
                "acc": (-3. * GRAV * np.ones(3), 3. * GRAV * np.ones(3)), "act": (np.zeros(4), np.ones(4)),
                "quat": (-np.ones(4), np.ones(4)), "t2w": (np.zeros(1), 5. * np.ones(1)), "t2t": (np.zeros(1), np.ones(1))}
        comp["xyzr"], comp["vxyzr"] = comp["xyz"], comp["vxyz"]
        names = obs_repr.split("_")
        self.observation_space = Box(np.concatenate([comp[c][0] for c in names]), np.concatenate([comp[c][1] for c in names]),
                                     dtype=np.float32)
        self.models, self.model_params, self.model_index, self._n_models = None, None, None, 0
        if per_env_dynamics is not None:
            # dynamics randomisation (quadrotor.py:738-763, :1030-1056): `per_env_dynamics` independently sampled models
            # (base sampler -> dynamics_change -> sampler 1 -> sampler 2 -> limits), dealt to the envs at random
            self._dyn_spec = (dynamics_params, dynamics_change, dyn_sampler_1, dyn_sampler_2, gravity)
            self.resample_dynamics(n_models=self.num_envs if per_env_dynamics == "all" else int(per_env_dynamics))

    def _action_bounds(self):
        """quadrotor_control.py action_space() of the selected controller."""
        c, t2w = self.controller_name, self.model.thrust_to_weight
        if c in ("raw", "vertical", "vertplane"):
            d = {"raw": 4, "vertical": 1, "vertplane": 2}[c]
            return (-np.ones(d) if self.cfg.zero_middle else np.zeros(d)), np.ones(d)
        if c == "shifted":
            return -np.ones(4), (t2w - 1.0) * np.ones(4)
        if c == "velocity_yaw":
            hi = np.array([20., 20., 20., 4 * np.pi])
            return -hi, hi
        hi = np.array([t2w - 1.0, 10 * np.pi, 10 * np.pi, 2 * np.pi])       # omega_thrust / mellinger :213-221, :514-522
        return np.array([-1.0, -hi[1], -hi[2], -hi[3]]), hi

    # ------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.qs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _alloc(self, shape, dtype):
        if self.host_mapped:
            return self.torch.zeros(shape, dtype=dtype, pin_memory=True)
        return self.torch.zeros(shape, dtype=dtype, device=self.device)

    def _in(self, name, x, dtype, shape):
        """A kernel input as a contiguous tensor the device can address: a device tensor, or (host_mapped) a persistent
        pinned staging buffer the value is copied into."""
        t = self.torch
        v = t.as_tensor(x, dtype=dtype, device=self.mem) if not isinstance(x, np.ndarray) else t.as_tensor(np.ascontiguousarray(x)).to(dtype=dtype, device=self.mem)
        if tuple(v.shape) != tuple(shape):
            v = v.reshape(shape)
        if not self.host_mapped:
            return v.contiguous()
        buf = self._stage.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = self._stage[name] = t.empty(shape, dtype=dtype, pin_memory=True)
        buf.copy_(v)
        return buf

    def _done_call(self):
        """host_mapped: results live in host memory, valid once the stream has drained."""
        if self.host_mapped:
            self.torch.cuda.current_stream(self.device).synchronize()

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    # ---- logical fields of the structure-of-arrays state (include/quadsim.h "State buffer") ---------
    # The per-env state is the word sequence [pos3 vel3 omega3 R9 tick ret rot_damp4 cmds_damp4 | optional rows],
    # cut into 4-word vectors; word w of env i lives at state[w // 4, i, w % 4].
    def _fields(self):
        L = self.layout
        return {"pos": (0, 3), "vel": (3, 3), "omega": (6, 3), "rot": (9, 9), "tick": (18, 1), "ep_return": (19, 1),
                "thrust_rot_damp": (20, 4), "thrust_cmds_damp": (24, 4), "ou_state": (L.row_ou * 4, 4),
                "prev_action": (L.row_prev_action * 4, 4), "gyro_bias": (L.row_gyro_bias * 4, 3),
                "goals": (L.row_goal * 4, 3)}

    def get_field(self, name):
        """Snapshot (copy) of one logical field: [N, count] (rot: [N,3,3]; tick/ep_return: [N])."""
        first, count = self._fields()[name]
        if first < 0:
            raise AttributeError("state has no %s row in this configuration" % name)
        idx = self.torch.arange(first, first + count, device=self.mem)
        v = self.state[idx // 4, :, idx % 4].transpose(0, 1).contiguous()
        if name == "rot":
            return v.reshape(self.num_envs, 3, 3)
        if count == 1:
            return v.reshape(self.num_envs)
        return v

    pos = property(lambda self: self.get_field("pos"))
    vel = property(lambda self: self.get_field("vel"))
    omega = property(lambda self: self.get_field("omega"))
    rot = property(lambda self: self.get_field("rot"))
    tick = property(lambda self: self.get_field("tick"))
    ep_return = property(lambda self: self.get_field("ep_return"))
    thrust_rot_damp = property(lambda self: self.get_field("thrust_rot_damp"))
    thrust_cmds_damp = property(lambda self: self.get_field("thrust_cmds_damp"))
    ou_state = property(lambda self: self.get_field("ou_state"))
    prev_action = property(lambda self: self.get_field("prev_action"))
    gyro_bias = property(lambda self: self.get_field("gyro_bias"))
    goals = property(lambda self: self.get_field("goals"))

    def set_field(self, name, value, mask=None):
        """Write one logical field for all envs (value broadcast over N) or for the envs selected by mask."""
        t = self.torch
        first, count = self._fields()[name]
        if first < 0:
            raise AttributeError("state has no %s row in this configuration" % name)
        v = t.as_tensor(value, dtype=self.tdtype, device=self.mem).reshape(-1, count).expand(self.num_envs, count)
        idx = t.arange(first, first + count, device=self.mem)
        if mask is not None:
            m = t.as_tensor(mask, device=self.mem).bool()
            v = t.where(m[:, None], v, self.state[idx // 4, :, idx % 4].transpose(0, 1))
        self.state[idx // 4, :, idx % 4] = v.transpose(0, 1)

    def set_state(self, pos, vel, rot, omega, mask=None):
        """QuadrotorDynamics.set_state + reset() (quadrotor.py:213-224, :438-440) for all (or masked) envs."""
        t = self.torch
        om = t.as_tensor(omega, dtype=self.tdtype, device=self.mem).float().to(self.tdtype)      # float32 cast, :223
        self.set_field("pos", pos, mask)
        self.set_field("vel", vel, mask)
        self.set_field("rot", t.as_tensor(rot, dtype=self.tdtype, device=self.mem).reshape(-1, 9), mask)
        self.set_field("omega", om, mask)
        self.set_field("thrust_rot_damp", [0., 0., 0., 0.], mask)
        self.set_field("thrust_cmds_damp", [0., 0., 0., 0.], mask)

    # ---- per-env dynamics ------------------------------------------------------
    def set_models(self, params_list, index=None):
        """Per-env model constants: `params_list` = K parameter dicts, `index` [N] = model of every env (default: env i
        flies model i % K).  Host cost is O(K) (QuadLink + Jacobian per model); the [rows, N, 4] buffer is gathered
        on the device."""
        t = self.torch
        if not self.cfg.per_env_model:
            raise RuntimeError("create the env with per_env_dynamics=... to use per-env models")
        from .quadrotor.model import ModelConstants
        consts = [p if isinstance(p, ModelConstants) else ModelConstants(p, gravity=self.model.gravity) for p in params_list]
        table = np.stack([m.rows(self.dt) for m in consts], axis=1)                       # [rows, K, 4]
        self.model_params, self.model_consts = list(params_list), consts
        self._bind_model_table(table, index)

    def _bind_model_table(self, table, index=None):
        """table: float64 [QS_MODEL_ROWS, K, 4]; index [N] (default env i -> model i % K, or the identity when K == N)."""
        t = self.torch
        K = table.shape[1]
        tab = t.as_tensor(table, dtype=self.tdtype, device=self.mem)
        if index is None and K == self.num_envs:
            idx, cols = t.arange(K, device=self.mem), tab
        else:
            if index is None:
                index = np.arange(self.num_envs) % K
            idx = t.as_tensor(np.asarray(index), device=self.mem, dtype=t.long)
            cols = tab[:, idx, :]
        self.models = self._in("models", cols, self.tdtype, (table.shape[0], self.num_envs, 4))
        self.model_index = idx
        # rows 8 / 9 (include/quadsim.h): linearity, C_drag, C_roll | vel_damp, damp_omega_quadratic.  Only when EVERY model
        # is plain there may the step take the kernels that compile those terms out (checked on the float64 host table)
        tb = np.asarray(table, dtype=np.float64)
        self.models_plain = bool(np.all(tb[8, :, 0] == 1.0) and not np.any(tb[8, :, 1:3]) and not np.any(tb[9, :, 0:2]))
        _lib.check(self.lib.qs_bind_models(self._h, self.models.data_ptr(), _lib.QS_MODELS_PLAIN if self.models_plain else 0),
                   self._h)

    def resample_dynamics(self, n_models=None, seed=None):
        """Draw a fresh set of models with the reference's sampler chain and deal them to the envs (call between
        episodes, like the reference's resample_dynamics: a model change mid-episode is not meaningful)."""
        import copy
        dynamics_params, dynamics_change, s1, s2, _ = self._dyn_spec
        K = int(n_models or self._n_models)
        self._n_models = K
        if seed is not None:
            np.random.seed(seed)
        if dynamics_params == "RandomQuad" and dynamics_change is None and s1 is None and s2 is None and K > 256:
            # many fully random models (up to one per env): the vectorised host pipeline, O(ms) instead of 0.4 ms each
            from .quadrotor import batched_models as bm
            self.model_batch = bm.sample_randomquad_batch(K)
            self.model_params = None
            self._bind_model_table(bm.model_rows_batch(self.model_batch, self.dt),
                                   None if K == self.num_envs else np.random.randint(0, K, size=self.num_envs))
            return self.model_batch
        out = []
        for _ in range(K):
            p = resolve_dynamics_params(dynamics_params, dynamics_change, check=False)
            for spec in (s1, s2):
                if spec is not None:
                    kw = copy.deepcopy(spec)
                    cls = kw.pop("class") if "class" in kw else kw.pop("type")
                    p = getattr(quad_rand, cls)(params=p, **kw).sample(p)
            out.append(quad_rand.check_quad_param_limits(p))
        self.set_models(out, index=np.random.randint(0, K, size=self.num_envs))
        return out

    # ---- noise injection (test mode) -----------------------------------------------------------------
    def set_noise(self, ou_z=None, obs_z=None, reset_z=None, goal_z=None):
        """QS_NOISE_INJECTED: unit draws for the next reset()/step() calls; shapes [N,sim_steps,4], [N,QS_OBS_Z], [N,QS_RESET_Z]
        (+ goal_z [N,4] for the excite / resample_goal options)."""
        t = self.torch

        def dev(name, x, shape):
            if x is None:
                return None
            return self._in(name, np.asarray(x), self.tdtype, shape)

        N = self.num_envs
        refs = (dev("ou_z", ou_z, (N, self.sim_steps, 4)), dev("obs_z", obs_z, (N, _lib.QS_OBS_Z)),
                dev("reset_z", reset_z, (N, _lib.QS_RESET_Z)))
        self._noise_refs = refs      # keep the device buffers alive until replaced
        ptr = lambda v: None if v is None else v.data_ptr()
        _lib.check(self.lib.qs_set_noise_injection(self._h, ptr(refs[0]), ptr(refs[1]), ptr(refs[2])), self._h)
        if self.cfg.resample_goal or self.cfg.excite:
            gz = dev("goal_z", goal_z, (N, _lib.QS_GOAL_Z))
            self._noise_refs = refs + (gz,)
            _lib.check(self.lib.qs_set_goal_injection(self._h, ptr(gz)), self._h)

    # ---- the env API -----------------------------------------------------------------------------------
    def reset(self, mask=None):
        m = None
        if mask is not None:
            m = self._in("mask", self.torch.as_tensor(mask).to(self.torch.uint8), self.torch.uint8, (self.num_envs,))
        _lib.check(self.lib.qs_reset(self._h, None if m is None else m.data_ptr(), self.obs.data_ptr(), self._stream()),
                   self._h)
        self._done_call()
        return self.obs

    def observe(self, accelerometer=None):
        """A fresh noisy observation of the current state without stepping (state_<obs_repr>(env) / add_noise)."""
        acc = None if accelerometer is None else self._in("acc_in", accelerometer, self.tdtype, (self.num_envs, 3))
        _lib.check(self.lib.qs_observe(self._h, None if acc is None else acc.data_ptr(), self.obs.data_ptr(), self._stream()),
                   self._h)
        self._done_call()
        return self.obs

    def _actions(self, actions, k=None):
        if actions is None:
            return None
        want = (self.num_envs, self.act_dim) if k is None else (k, self.num_envs, self.act_dim)
        return self._in("actions", actions, self.tdtype, want)

    def step(self, actions=None):
        a = self._actions(actions)
        _lib.check(self.lib.qs_step(self._h, None if a is None else a.data_ptr(), self.obs.data_ptr(), self.rew.data_ptr(),
                                    self._done_u8.data_ptr(), self._stream()), self._h)
        self._done_call()
        return self.obs, self.rew, self.done

    def capture_steps(self, n_steps, actions=None):
        """CUDA graph of `n_steps` consecutive env steps reading `actions` ([N, A] device tensor the caller refills
        between replays, or None for the in-kernel Mellinger controller).  The clock lives on the device, so every
        replay continues the simulation with fresh noise; launches carry the PDL attribute, so consecutive steps
        overlap launch latency (the caller loop of quadrotor.py:1278-1305 without per-step host work).
        Returns the torch.cuda.CUDAGraph; outputs land in self.obs / self.rew / self.done."""
        t = self.torch
        if self.host_mapped:
            raise RuntimeError("capture_steps: host_mapped envs synchronise the stream on every call and cannot be captured")
        if actions is not None:
            # the graph bakes the POINTER in: it must be the caller's own tensor, not a converted temporary
            if not (isinstance(actions, t.Tensor) and actions.is_cuda and actions.device == self.device and
                    actions.dtype == self.tdtype and actions.is_contiguous() and
                    tuple(actions.shape) == (self.num_envs, self.act_dim)):
                raise ValueError("capture_steps: actions must be a contiguous %s tensor of shape (%d, %d) on %s"
                                 % (self.tdtype, self.num_envs, self.act_dim, self.device))
            self._graph_actions = actions                # keep it alive as long as the env
        a = actions
        stream = t.cuda.Stream(device=self.device)
        stream.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(stream):
            self.step(a)                                  # warm-up outside the capture (lazy module loading)
        stream.synchronize()
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g, stream=stream):
            for _ in range(int(n_steps)):
                self.step(a)
        return g

    def step_k(self, k, actions=None):
        """k fused env steps in one launch; returns (last obs, summed reward, any-done)."""
        a = self._actions(actions, k)
        _lib.check(self.lib.qs_step_k(self._h, int(k), None if a is None else a.data_ptr(), self.obs.data_ptr(),
                                      self.rew.data_ptr(), self._done_u8.data_ptr(), self._stream()), self._h)
        self._done_call()
        return self.obs, self.rew, self.done

    def step_host(self, actions_host, obs_host, rew_host, done_host):
        """End-to-end step on HOST tensors (pinned for full PCIe speed): H2D, fused step, D2H, synchronous.  Ordered after
        the work already queued on torch's current stream (reset(), step(), set_field() ...)."""
        for x in (obs_host, rew_host, done_host):
            assert x.device.type == "cpu" and x.is_contiguous()
        ap = None
        if actions_host is not None:
            assert actions_host.device.type == "cpu" and actions_host.is_contiguous() and actions_host.dtype == self.tdtype
            ap = actions_host.data_ptr()
        assert obs_host.dtype == self.tdtype and rew_host.dtype == self.tdtype and done_host.dtype == self.torch.uint8
        _lib.check(self.lib.qs_step_host(self._h, ap, obs_host.data_ptr(), rew_host.data_ptr(), done_host.data_ptr(),
                                         self._stream()), self._h)
        return obs_host, rew_host, done_host

    def dynamics_step(self, actions, controller=None, n_substeps=None):
        """controller.step_func / QuadrotorDynamics.step on the bound state; controller=None: raw motor commands."""
        cid = -1 if controller is None else CONTROLLERS[controller]
        adim = 4 if cid < 0 else [4, 1, 2, 4, 4, 4, 4][cid]
        a = None
        if actions is not None:
            a = self._in("dyn_actions", actions, self.tdtype, (self.num_envs, adim))
        n = self.sim_steps if n_substeps is None else int(n_substeps)
        _lib.check(self.lib.qs_dynamics_step(self._h, cid, None if a is None else a.data_ptr(), n, self._stream()), self._h)
        self._done_call()

    def affine_fg(self, s, dt, want_F=True, want_G=True):
        """QuadrotorDynamics.F(s, dt) / G(s) (quadrotor.py:461-518) for n states s [n, 21] = xyz, Vxyz, R (row-major), omega,
        goal: returns (F [n, 21], G [n, 21, C]) device tensors, C = 1 / 2 / 4 controls for dim_mode 1D / 2D / 3D."""
        t = self.torch
        x = t.as_tensor(np.asarray(s) if not isinstance(s, t.Tensor) else s, dtype=self.tdtype, device=self.mem).reshape(-1, 21).contiguous()
        n = int(x.shape[0])
        C_ = {1: 1, 2: 2, 3: 4}[int(self.cfg.dim_mode)]
        F = t.empty((n, 21), dtype=self.tdtype, device=self.mem, pin_memory=self.host_mapped) if want_F else None
        G = t.empty((n, 21, C_), dtype=self.tdtype, device=self.mem, pin_memory=self.host_mapped) if want_G else None
        if self.host_mapped:
            x = x.pin_memory()
        _lib.check(self.lib.qs_affine_fg(self._h, x.data_ptr(), float(dt), None if F is None else F.data_ptr(),
                                         None if G is None else G.data_ptr(), n, self._stream()), self._h)
        self._done_call()
        self._affine_in = x                            # keep the input alive until the stream has consumed it
        return F, G

    # ---- bookkeeping -----------------------------------------------------------------------------------
    def stats(self, clear=False):
        out = (C.c_double * 3)()
        _lib.check(self.lib.qs_get_stats(self._h, C.byref(out), int(clear), self._stream()), self._h)
        return np.array([out[0], out[1], out[2]])

    def stats_tensor(self):
        """The device-resident double[3] {sum ret, sum ret^2, episodes} as a torch tensor (for NCCL all-reduce)."""
        p = C.c_void_p()
        _lib.check(self.lib.qs_stats_device_ptr(self._h, C.byref(p)), self._h)
        from .stats import wrap_device_doubles
        return wrap_device_doubles(p.value, 3, self.device)

    def check_finite(self):
        _lib.check(self.lib.qs_check(self._h, self._stream()), self._h)

    @property
    def clock(self):
        t, s = C.c_double(), C.c_uint64()
        _lib.check(self.lib.qs_get_clock(self._h, C.byref(t), C.byref(s), self._stream()), self._h)
        return t.value, s.value

    @clock.setter
    def clock(self, value):
        _lib.check(self.lib.qs_set_clock(self._h, float(value[0]), int(value[1]), self._stream()), self._h)

    @property
    def launch_count(self):
        n = C.c_uint64()
        _lib.check(self.lib.qs_launch_count(self._h, C.byref(n)), self._h)
        return n.value

    @property
    def kernel_variant(self):
        return self.lib.qs_last_variant(self._h).decode()

    @property
    def host_path(self):
        return self.lib.qs_last_host_path(self._h).decode()

    def state_dict(self):
        """Checkpoint: the state tensor plus the global clocks (SURVEY.md section 5 "Checkpoint / resume")."""
        t, s = self.clock
        d = {"state": self.state.clone(), "svd_timer": t, "step_count": s}
        if self.models is not None:                       # per-env dynamics: the model constants are part of the state
            d["models"] = self.models.clone()
        return d

    def load_state_dict(self, d):
        self.state.copy_(d["state"])
        self.clock = (d["svd_timer"], d["step_count"])
        if "models" in d and self.models is not None:
            self.models.copy_(d["models"])
