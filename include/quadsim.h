/* quadsim.h - C-ABI of libquadsim.so: the B200-native (sm_100a) batched quadrotor env-step path.
 *
 * The reference (amolchanov86/gym_art) is pure Python and has no FFI; its seam for this path is
 * Python duck typing:
 *     QuadrotorEnv.reset()/step(action)            gym_art/quadrotor/quadrotor.py:1149-1156 (-> _reset :1059, _step :942)
 *     controller.step_func(dynamics, action, ...)  gym_art/quadrotor/quadrotor.py:966-970, quadrotor_control.py
 *     QuadrotorDynamics.step(thrust_cmds, dt)      gym_art/quadrotor/quadrotor.py:261-262 (step1 :273-436)
 *     SensorNoise.add_noise(...)                   gym_art/quadrotor/sensor_noise.py:100-159
 *     compute_reward_weighted(...)                 gym_art/quadrotor/quadrotor.py:544-638
 * Each entry point below names the reference interface it replaces.  The binding a maintainer of the
 * reference would add (ctypes) is shown in INTEGRATION.md and implemented in gym_art_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  Every function returns QS_OK (0) or a negative
 *     QsStatus; nothing throws across the ABI; the text of the last failure is qs_last_error().
 *   - every buffer is owned by the caller (normally torch tensors on the handle's device).  "real" is
 *     float for QS_F32 handles and double for QS_F64 handles.  The state and model buffers must be 32-byte aligned,
 *     actions and observations aligned to 4 reals (16 bytes f32 / 32 bytes f64; cudaMalloc / torch allocations are).
 *   - all work is enqueued on the caller's stream; the library never synchronises the device except in
 *     the *_host entry points, which are synchronous by definition.
 *   - a handle is bound to one device and is not thread-safe; use one handle per GPU.
 */
#ifndef QUADSIM_H
#define QUADSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QS_ABI_VERSION 8

typedef struct QsHandle QsHandle;
typedef void* qs_stream_t; /* cudaStream_t */

typedef enum QsStatus {
  QS_OK = 0,
  QS_EINVAL = -1,       /* bad argument / inconsistent config */
  QS_ECUDA = -2,        /* a CUDA runtime call failed */
  QS_ENONFINITE = -3,   /* a non-finite reward was produced (reference: ValueError, quadrotor.py:633-636) */
  QS_ESTATE = -4,       /* state / noise buffers not bound */
  QS_EUNSUPPORTED = -5, /* valid in the reference but not built here */
  QS_EABI = -6          /* QsConfig.abi_version / struct_size mismatch */
} QsStatus;

typedef enum QsDtype { QS_F32 = 0, QS_F64 = 1 } QsDtype;

/* quadrotor_control.py: RawControl :65, VerticalControl :102, VertPlaneControl :146, ShiftedMotorControl :11,
 * OmegaThrustControl :208, VelocityYawControl :238, NonlinearPositionController (Mellinger) :286 */
typedef enum QsController {
  QS_CTRL_RAW = 0,
  QS_CTRL_VERTICAL = 1,
  QS_CTRL_VERTPLANE = 2,
  QS_CTRL_SHIFTED = 3,
  QS_CTRL_OMEGA_THRUST = 4,
  QS_CTRL_VELOCITY_YAW = 5,
  QS_CTRL_MELLINGER = 6,
  QS_CTRL_COUNT = 7
} QsController;

/* get_state.py: the six observation builders that run in the reference as shipped (:5 :134 :147 :219 :236 :249) and
 * the six that run once the module is given the names it uses without importing them (`normal`, `R2quat`: quaternion
 * builders :276 :289 :307, thrust-to-weight / torque-to-thrust builders :325 :345 :363).  The remaining builders of that
 * file cannot run at all (invalid np.array calls, undefined locals, components unknown to make_observation_space). */
typedef enum QsObsRepr {
  QS_OBS_XYZ_VXYZ_R_OMEGA = 0,          /* 18 */
  QS_OBS_XYZ_VXYZ_R_OMEGA_H = 1,        /* 19 */
  QS_OBS_XYZR_VXYZR_R_OMEGA = 2,        /* 18 */
  QS_OBS_XYZR_VXYZR_R_OMEGA_H = 3,      /* 19 */
  QS_OBS_XYZ_VXYZ_R_OMEGA_ACC_ACT = 4,  /* 25 */
  QS_OBS_XYZ_VXYZ_R_OMEGA_ACT = 5,      /* 22 */
  QS_OBS_XYZ_VXYZ_QUAT_OMEGA = 6,       /* 13: R2quat(R) (quad_utils.py:101-109) x small-angle noise quaternion */
  QS_OBS_XYZR_VXYZR_QUAT_OMEGA = 7,     /* 13 */
  QS_OBS_XYZR_VXYZR_QUAT_OMEGA_H = 8,   /* 14 */
  QS_OBS_XYZR_VXYZR_R_OMEGA_T2W = 9,    /* 19: + noisy thrust_to_weight clipped to [1.5, 10] and scaled to [0, 1] */
  QS_OBS_XYZ_VXYZ_R_OMEGA_T2W = 10,     /* 19 */
  QS_OBS_XYZ_VXYZ_R_OMEGA_T2W_T2T = 11, /* 20: + noisy torque_to_thrust clipped to [0.005, 1], scaled to [0, 1] */
  QS_OBS_COUNT = 12
} QsObsRepr;

/* QsConfig.init_random_state: which law a reset draws velocity, attitude and body rates from */
typedef enum QsInitMode {
  QS_INIT_DEFAULT = 0, /* at rest, yawed towards the goal (quadrotor.py:1118-1126) */
  QS_INIT_RANDOM = 1,  /* init_random_state=True: QuadrotorDynamics.random_state in 3D (:227-239, :1113), the 1D / 2D
                        * laws of :1101-1110 otherwise */
  QS_INIT_PITCH_ROLL_RESTRICTED = 2 /* QuadrotorDynamics.pitch_roll_restricted_random_state (:243-259): Euler angles
                        * U(-pitch_max, pitch_max), U(-roll_max, roll_max), U(-yaw_max, yaw_max), R = Rz Ry Rx (3D only) */
} QsInitMode;

typedef enum QsNoiseMode {
  QS_NOISE_PHILOX = 0,  /* on-device Philox4x32-10, counter = (global env id, step, slot), key = seed */
  QS_NOISE_INJECTED = 1 /* unit draws supplied by the host (qs_set_noise_injection): parity tests */
} QsNoiseMode;

/* Injected-noise slot layout (unit normals "n", unit uniforms in [-1,1) "u").
 * obs_z  [n_envs][QS_OBS_Z]   : one SensorNoise.add_noise call, sensor_noise.py:116-157 */
#define QS_OBS_Z 33
#define QS_Z_NPOS 0
#define QS_Z_UPOS 3
#define QS_Z_NVEL 6
#define QS_Z_UVEL 9
#define QS_Z_NGYRO 12    /* white gyro noise, or the (folded) bias-walk normal when gyro_norm_std != 0 */
#define QS_Z_NGYRO_RW 15 /* gyro_random_walk normal, only when gyro_norm_std != 0 (sensor_noise.py:170) */
#define QS_Z_NTHETA 18
#define QS_Z_UTHETA 21
#define QS_Z_NACC_S 24
#define QS_Z_NACC_D 27
#define QS_Z_NT2W 30     /* unit normal of the thrust-to-weight observation noise (get_state.py:334); [31] torque-to-thrust */
#define QS_Z_NT2T 31
/* reset_z [n_envs][QS_RESET_Z] : QuadrotorEnv._reset, quadrotor.py:1087-1126 (+ :227-239, quad_utils.py:47-57,139) */
#define QS_RESET_Z 51
#define QS_RZ_UPOS 0    /* spawn position = goal + box * u              (:1087) */
#define QS_RZ_YAW 3     /* injected: accepted yaw / pi ; Philox: yaw = atan2(-y,-x) + (pi/3) u   (:1124-1126) */
#define QS_RZ_UVEL 4    /* init_random_state: velocity direction        (:230) */
#define QS_RZ_UVELMAG 7 /*                    velocity magnitude        (:231) */
#define QS_RZ_UOM 8     /*                    omega direction           (:234) */
#define QS_RZ_UOMMAG 11 /*                    omega magnitude           (:235) */
#define QS_RZ_NUP 12    /*                    "up" Gaussian 3-vector    (quad_utils.py:49);
                         * QS_INIT_PITCH_ROLL_RESTRICTED: unit uniforms of pitch, roll, yaw (quadrotor.py:254-256);
                         * 1D: [QS_RZ_UVELMAG] upward speed; 2D: [QS_RZ_UOMMAG] pitch rate, [QS_RZ_UVEL..] velocity,
                         * [QS_RZ_YAW] pitch angle - unit symmetric draws u standing for rand() = (1 + u) / 2 (:1101-1110) */
#define QS_RZ_NFWD 15   /*                    accepted "fwd" Gaussian   (quad_utils.py:50-52) */
#define QS_RZ_OBS 18    /* obs_z of the observation returned by the reset (:1143) */
/* ou_z   [n_envs][sim_steps][4] : OUNoise.noise() normals, one per substep (quad_utils.py:197-201) */
/* goal_z [n_envs][QS_GOAL_Z] (qs_set_goal_injection): unit uniforms of the moving-goal options -
 *   [0..2] excite goal x, y, z (quadrotor.py:959-961), [3] resample_goal height (quadrotor.py:1079) */
#define QS_GOAL_Z 4

/* Optional per-env diagnostics, info [n_envs][QS_INFO_DIM] (reference: info['obs_comp'], quadrotor.py:994-1006) */
#define QS_INFO_DIM 48
#define QS_INFO_ACC 0            /* world-frame acceleration, dynamics.acc */
#define QS_INFO_ACCELEROMETER 3  /* dynamics.accelerometer */
#define QS_INFO_TORQUE 6         /* dynamics.torque */
#define QS_INFO_OMEGA_DOT 9      /* dynamics.omega_dot */
#define QS_INFO_CMDS 12          /* controller.action: the 4 normalized motor commands */
#define QS_INFO_CRASHED 16
#define QS_INFO_COLLIDED 17
#define QS_INFO_COSTS 18         /* 10 weighted costs in the reference's summation order (quadrotor.py:593-604) */
#define QS_INFO_EP_RETURN 28     /* running return of the episode the step belonged to */
#define QS_INFO_TICK 29          /* tick after the step (before any auto-reset) */
#define QS_INFO_REWARD 30        /* reward of the (last) step */
#define QS_INFO_DONE 31          /* done flag of the (last) step */
#define QS_INFO_COSTS_RAW 32     /* the same 10 costs before weighting: rew_info 'rewraw_*' (quadrotor.py:619-630) */

/* Per-env model buffer (dynamics randomisation, quadrotor_randomization.py + quadrotor.py:1030-1056 resample_dynamics):
 * real[QS_MODEL_ROWS][n_envs][4], read-only for the kernels, one 16-byte vector per env per row:
 *   0 thrust_max[4]            1 torque_max[k] * prop_ccw[k]     2 prop_pos[k].y     3 -prop_pos[k].x     4 prop_pos[k].z
 *   5 inertia.xyz, mass        6 1/inertia.xyz, 1/mass
 *   7 tau_up, tau_down (= min(1, 4 dt / (damp_time + 1e-6)), quadrotor.py:284-288), ou_sigma, arm
 *   8 linearity, C_drag, C_roll, thrust_to_weight      9 vel_damp, damp_omega_quadratic, torque_to_thrust, 0
 *   10..13 rows of the inverse Jacobian (quadrotor_control.py:192-211) */
#define QS_MODEL_ROWS 14

/* QuadrotorDynamics constants (quadrotor.py:142-208) + the controllers' inverse Jacobian (quadrotor_control.py:192-211) */
typedef struct QsModel {
  double mass;
  double inertia[3];
  double thrust_max[4];
  double torque_max[4];
  double prop_pos[4][3];
  double prop_ccw[4];
  double linearity, C_drag, C_roll;
  double damp_time_up, damp_time_down;
  double vel_damp, damp_omega_quadratic;
  double ou_theta, ou_sigma; /* OUNoise(4, sigma = 0.2 * thrust_noise_ratio), quadrotor.py:198 */
  double arm, omega_max;
  double gravity;            /* self.gravity: accelerometer only (:436); the integrator uses GRAV = 9.81 (:426) */
  double thrust_to_weight;
  double torque_to_thrust;   /* t2t observation component only (get_state.py:379) */
  double Jinv[4][4];
} QsModel;

/* rew_coeff (quadrotor.py:799-806); log_pos selects the quadrotor_multi variant (quadrotor_multi.py:554, :810-818) */
typedef struct QsReward {
  double pos, effort, action_change, crash, orient, yaw, rot, attitude, spin, vel;
  double pos_offset, pos_log_weight, pos_linear_weight;
  double collision; /* swarm only (ours): cost per step while a drone is in contact */
  int32_t log_pos;
  int32_t _pad;
} QsReward;

/* SensorNoise ctor arguments (sensor_noise.py:58-63); enabled = !bypass */
typedef struct QsSensorNoise {
  int32_t enabled;
  int32_t _pad;
  double pos_norm_std, pos_unif_range, vel_norm_std, vel_unif_range;
  double quat_norm_std, quat_unif_range;
  double gyro_norm_std, gyro_noise_density, gyro_random_walk, gyro_bias_correlation_time;
  double acc_static_noise_std, acc_dynamic_noise_ratio;
} QsSensorNoise;

/* QuadrotorEnv ctor arguments (quadrotor.py:653-658) flattened */
typedef struct QsConfig {
  int32_t abi_version; /* QS_ABI_VERSION */
  int32_t struct_size; /* sizeof(QsConfig) */
  int64_t n_envs;      /* drones on this handle (= swarm envs * n_agents) */
  int64_t env_id_offset; /* global id of local env 0: Philox streams are invariant to the GPU count */
  uint64_t seed;
  int32_t n_agents;    /* 1 = independent envs; M > 1 = swarms of M consecutive drones (M in 2,4,8,16,32) */
  int32_t device;
  int32_t dtype;       /* QsDtype */
  int32_t controller;  /* QsController */
  int32_t zero_middle; /* raw_control_zero_middle */
  int32_t obs_repr;    /* QsObsRepr */
  int32_t noise_mode;  /* QsNoiseMode */
  int32_t init_random_state; /* QsInitMode */
  int32_t auto_reset;  /* fold reset() into step(): on done return the post-reset observation */
  int32_t dim_mode;    /* 1, 2 or 3 (quadrotor.py:1089-1092, :1120-1121) */
  int32_t sim_steps;
  int32_t ep_len;
  int32_t per_env_goal; /* goals live in the state buffer (row given by QsStateLayout.row_goal) */
  int32_t resample_goal; /* goal = (0, 0, U(0.5, 2)) at every reset (quadrotor.py:1078-1079); needs per_env_goal */
  int32_t excite;      /* goal = (U(-.5,.5), U(-.5,.5), U(1.5,2.5)) every 5th tick (quadrotor.py:957-962); needs per_env_goal */
  int32_t _pad;
  double dt;           /* 1 / sim_freq */
  double svd_period;   /* since_last_svd_limit = 0.5 s (quadrotor.py:105) */
  double goal[3];      /* (0, 0, 2) (quadrotor.py:1081) */
  double spawn_box;    /* self.box = 2 (quadrotor.py:728) */
  double spawn_z_min;  /* spawn height floor 0.25 (quadrotor.py:1094) */
  double room_box[2][3];
  double max_init_vel, max_init_omega; /* quadrotor.py:717-718 */
  double collision_k;  /* swarm: contact when |p_i - p_j| < collision_k * 2 * arm (per-env models: collision_k * (arm_i + arm_j)) */
  double pitch_max, roll_max, yaw_max; /* QS_INIT_PITCH_ROLL_RESTRICTED (quadrotor.py:243: 0.5, 0.5, 3.14) */
  double t2w_std, t2t_std; /* t2w / t2t observation noise (quadrotor.py:658: 0.005, 0.0005) */
  int32_t use_formation; /* swarm: drone a of every swarm flies to formation[a] (no goal row in the state buffer) */
  int32_t per_env_model; /* dynamics randomisation: model constants come from qs_bind_models, not from `model` */
  double formation[32][3];
  QsModel model;
  QsReward reward;
  QsSensorNoise sense;
} QsConfig;

/* State buffer: real[rows][n_envs][4], one 16-byte (f32) vector per env per row, rows contiguous over envs.
 *   row 0: pos.xyz, vel.x        row 1: vel.yz, omega.xy      row 2: omega.z, R00 R01 R02
 *   row 3: R10 R11 R12 R20       row 4: R21 R22, tick, ep_return
 *   row 5: thrust_rot_damp[4]    row 6: thrust_cmds_damp[4]
 *   optional rows (index or -1): OU state (iff ou_sigma > 0 or per_env_model), previous action (iff the observation
 *   contains `act` or the action_change reward coefficient is non-zero), gyro bias (iff gyro_norm_std != 0),
 *   goal (iff per_env_goal, resample_goal or excite). */
typedef struct QsStateLayout {
  int32_t rows;
  int32_t row_ou, row_prev_action, row_gyro_bias, row_goal;
  int32_t obs_dim, act_dim;
  int32_t _pad;
} QsStateLayout;

int qs_abi_version(void);

/* Pure function of the config; no GPU needed. */
int qs_state_layout(const QsConfig* cfg, QsStateLayout* out);

/* replaces QuadrotorEnv.__init__ (quadrotor.py:653-827) minus the host-side parameter pipeline */
int qs_create(const QsConfig* cfg, QsHandle** out);
int qs_destroy(QsHandle* h);
const char* qs_last_error(const QsHandle* h); /* h may be NULL: error of the last failed qs_create */

/* state: device buffer real[rows][n_envs][4]; the library never frees it */
int qs_bind_state(QsHandle* h, void* state);
/* QsConfig.per_env_model handles: device buffer real[QS_MODEL_ROWS][n_envs][4]; may be re-bound or rewritten between
 * steps (the reference resamples dynamics between episodes); the library never frees it */
#define QS_MODELS_PLAIN 1 /* the caller asserts that EVERY bound model has linearity == 1 and C_drag = C_roll = vel_damp =
                           * damp_omega_quadratic = 0 (rows 8 and 9), which lets the step take the kernels specialised for
                           * that case.  Without the flag the run-time-parameterised kernel is used (always correct). */
int qs_bind_models(QsHandle* h, const void* models, int32_t flags);
/* optional (auto_reset handles): real[n_envs][obs_dim].  A step that ends an env's episode returns the FIRST observation
 * of the next episode in `obs`; the observation of the terminal state - what the reference's step() returns on its done
 * step (quadrotor.py:988) - is written to this buffer, rows of envs that did not finish are left untouched.
 * NULL to disable. */
int qs_bind_terminal_obs(QsHandle* h, void* terminal_obs);
/* optional diagnostics buffer real[n_envs][QS_INFO_DIM], aligned to 4 reals (rows leave as vector stores); NULL to disable */
int qs_bind_info(QsHandle* h, void* info);
/* QS_NOISE_INJECTED: device buffers of unit draws consumed by the next reset/step calls (any may be NULL = zeros) */
int qs_set_noise_injection(QsHandle* h, const void* ou_z, const void* obs_z, const void* reset_z);

/* QS_NOISE_INJECTED + excite / resample_goal: device buffer real[n_envs][QS_GOAL_Z] (NULL = zeros) */
int qs_set_goal_injection(QsHandle* h, const void* goal_z);

/* replaces QuadrotorEnv.reset() (quadrotor.py:1059-1144) for the envs whose mask byte is non-zero
 * (mask NULL = all).  obs: real[n_envs][obs_dim], rows of unmasked envs are left untouched. */
int qs_reset(QsHandle* h, const uint8_t* mask, void* obs, qs_stream_t stream);

/* replaces state_<obs_repr>(env) (get_state.py) / SensorNoise.add_noise (sensor_noise.py:100-159): a fresh noisy
 * observation of the bound state without stepping (one add_noise call; injected mode reads the reset-observation slots
 * reset_z[QS_RZ_OBS..]).  accelerometer: optional clean body-frame readings real[n_envs][3] (NULL = (0, 0, 9.81), the
 * value set_state leaves behind, quadrotor.py:221).  obs: real[n_envs][obs_dim]. */
int qs_observe(QsHandle* h, const void* accelerometer, void* obs, qs_stream_t stream);

/* replaces QuadrotorEnv.step() (quadrotor.py:942-1028): controller -> sim_steps x step1 -> collision ->
 * crash/reward/tick/done -> (auto-reset) -> observation, one fused kernel.
 * actions: real[n_envs][act_dim] (may be NULL for QS_CTRL_MELLINGER: the reference ignores the action and the
 * effort cost then sees zeros); obs: real[n_envs][obs_dim]; rew: real[n_envs]; done: uint8[n_envs]. */
int qs_step(QsHandle* h, const void* actions, void* obs, void* rew, uint8_t* done, qs_stream_t stream);

/* K consecutive env steps in one launch with the state held in registers (caller loop quadrotor.py:1278-1305).
 * actions: real[K][n_envs][act_dim] or NULL (Mellinger / zero action); obs: observation after the last step;
 * rew: sum of the K rewards; done: OR of the K done flags. */
int qs_step_k(QsHandle* h, int k, const void* actions, void* obs, void* rew, uint8_t* done, qs_stream_t stream);

/* replaces controller.step_func(dynamics, action, goal, dt) / QuadrotorDynamics.step(cmds, dt)
 * (quadrotor_control.py, quadrotor.py:261): controller mixing + n_substeps x step1 on the bound state, no
 * reward/obs/tick.  controller = -1 passes `actions` straight through as normalized motor commands. */
int qs_dynamics_step(QsHandle* h, int controller, const void* actions, int n_substeps, qs_stream_t stream);

/* replaces QuadrotorDynamics.F(s, dt) / G(s) (quadrotor.py:461-518): the affine representation s' = dt (F(s) + G(s) u) of
 * the dynamics, evaluated for n caller-supplied states s [n][21] = xyz, Vxyz, R (row-major), omega, goal with the handle's
 * model (every state with model 0 of the model buffer on per_env_model handles).  F: real[n][21]; G: real[n][21][C], C = 1 /
 * 2 / 4 controls for dim_mode 1D / 2D / 3D (quadrotor.py:121-129).  Either output may be NULL.  Does not touch the bound
 * state. */
int qs_affine_fg(QsHandle* h, const void* s, double dt, void* F, void* G, int64_t n, qs_stream_t stream);

/* Same contract as qs_step but with HOST buffers; returns after the results are in host memory.  The step is ordered
 * after everything already queued on `stream` (resets, steps, state writes).  Pinned buffers (cudaHostAlloc /
 * cudaHostRegister, e.g. torch pin_memory()) aligned to 4 reals are read and written by the step kernel directly across
 * PCIe in one launch on `stream`; pageable or unaligned buffers go through a chunked H2D / kernel / D2H pipeline over
 * internal streams that wait for `stream` first. */
int qs_step_host(QsHandle* h, const void* actions_host, void* obs_host, void* rew_host, uint8_t* done_host, qs_stream_t stream);

/* Finished-episode statistics accumulated on the device: {sum of returns, sum of squared returns, episodes}.
 * qs_stats_device_ptr exposes the double[3] device buffer so the caller can all-reduce it over NCCL. */
int qs_get_stats(QsHandle* h, double out[3], int clear, qs_stream_t stream);
int qs_stats_device_ptr(QsHandle* h, double** out);

/* Sticky device-side flag: QS_ENONFINITE once any env produced a non-finite reward.  Synchronises the stream. */
int qs_check(QsHandle* h, qs_stream_t stream);

/* Global bookkeeping that is part of a checkpoint: the re-orthonormalisation timer (quadrotor.py:104,381-386;
 * a global phase, never reset by reset()) and the Philox step counter.  Both live on the device (the last CTA
 * of every launch advances them), so a sequence of qs_step calls can be captured in a CUDA graph and replayed.
 * These two calls copy across PCIe and synchronise the stream. */
int qs_get_clock(QsHandle* h, double* svd_timer, uint64_t* step_count, qs_stream_t stream);
int qs_set_clock(QsHandle* h, double svd_timer, uint64_t step_count, qs_stream_t stream);

/* Number of kernels this handle has launched (bench.py reports it as gpu_launches). */
int qs_launch_count(const QsHandle* h, uint64_t* out);
/* Name of the kernel variant the last step launched ("tma-raw", "tma-generic-drag", "generic-philox", ...): diagnostics. */
const char* qs_last_variant(const QsHandle* h);
/* How the last qs_step_host moved its data: "zero-copy" (pinned buffers, read / written by the kernel itself across
 * PCIe) or "staged" (pageable buffers: chunked cudaMemcpyAsync pipeline). */
const char* qs_last_host_path(const QsHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* QUADSIM_H */
