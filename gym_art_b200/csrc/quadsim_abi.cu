// quadsim_abi.cu - the C-ABI of libquadsim.so (include/quadsim.h): config validation, QsConfig -> kernel
// parameter block, variant dispatch, clocks, host-buffer pipeline.  No torch, no Python.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>

#include "../../include/quadsim.h"
#include "qs_device.cuh"

namespace {

thread_local std::string g_create_error;

const int kObsDims[QS_OBS_COUNT] = {18, 19, 18, 19, 25, 22, 13, 13, 14, 19, 19, 20};
const int kActDims[QS_CTRL_COUNT] = {4, 1, 2, 4, 4, 4, 4};
const int kHostStreams = 4;

}  // namespace

struct QsHandle {
  QsConfig cfg;
  QsStateLayout lay;
  void* state;
  void* info;
  void* term_obs;
  const void* models;
  int models_plain;          // caller asserted QS_MODELS_PLAIN for the bound model buffer
  const void* ou_z;
  const void* obs_z;
  const void* reset_z;
  const void* goal_z;
  double* d_stats;
  int* d_status;
  qs::Clock* d_clock;        // device-resident {svd_timer, step}: steps stay CUDA-graph capturable
  unsigned int* d_ticket;    // [0]: clock ticket; [1 + 2k], [2 + 2k]: launch ticket and tile counter of launch slot k
  uint64_t launches;
  int pdl;
  std::string err;
  const char* last_variant;
  // qs_step_host scratch (library-owned)
  cudaStream_t hs[kHostStreams];
  int host_streams_made;     // staged path resources, created one by one (a failed setup resumes where it stopped)
  int host_zero_copy;        // A/B knob QUADSIM_HOST_ZEROCOPY (default 1)
  int tma_generic;           // A/B knob QUADSIM_TMA_GENERIC (default 1)
  int ow_threads, ow_extra, host_ctas, prefetch, spec_draw;   // A/B knobs QUADSIM_OW_THREADS / QUADSIM_PREFETCH / QUADSIM_SPEC_DRAW (see Params)
#ifdef QS_PHASE_TIMING
  unsigned long long* timing;   // QUADSIM_TIMING_PTR: stamp buffer of the experiment build (tools/build_timing.sh)
#endif
  cudaEvent_t order_ev;      // qs_step_host: orders the library's internal streams after the caller's stream
  const char* last_host_path;
  void *h_act, *h_obs, *h_rew;
  uint8_t* h_done;
};

namespace {

int fail(QsHandle* h, int code, const char* fmt, const char* a = "") {
  char buf[512];
  snprintf(buf, sizeof(buf), fmt, a);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define QS_CUDA(h, call)                                                        \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) return fail((h), QS_ECUDA, #call ": %s", cudaGetErrorString(e_)); \
  } while (0)

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// one 4-wide vector of `real`: what every vector load / store and bulk copy of the kernels assumes
size_t vec_align(const QsHandle* h) { return h->cfg.dtype == QS_F64 ? 32 : 16; }
bool vec_aligned(const QsHandle* h, const void* p) { return ((uintptr_t)p) % vec_align(h) == 0; }

// device-side alias of a pinned host pointer (false for pageable memory or a pointer the kernels cannot use with vector
// accesses: the caller then takes the staged path)
bool device_alias(const QsHandle* h, const void* host, void** dev, bool need_vec) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
  if (need_vec && !vec_aligned(h, a.devicePointer)) return false;
  *dev = a.devicePointer;
  return true;
}

int validate(const QsConfig* c, std::string* why) {
  if (!c) { *why = "cfg is NULL"; return QS_EINVAL; }
  if (c->abi_version != QS_ABI_VERSION || c->struct_size != (int32_t)sizeof(QsConfig)) {
    *why = "QsConfig abi_version/struct_size mismatch"; return QS_EABI;
  }
  if (c->n_envs <= 0) { *why = "n_envs must be positive"; return QS_EINVAL; }
  if (c->dtype != QS_F32 && c->dtype != QS_F64) { *why = "dtype must be QS_F32 or QS_F64"; return QS_EINVAL; }
  if (c->controller < 0 || c->controller >= QS_CTRL_COUNT) { *why = "unknown controller"; return QS_EINVAL; }
  if (c->obs_repr < 0 || c->obs_repr >= QS_OBS_COUNT) { *why = "unknown obs_repr"; return QS_EINVAL; }
  if (c->noise_mode != QS_NOISE_PHILOX && c->noise_mode != QS_NOISE_INJECTED) { *why = "unknown noise_mode"; return QS_EINVAL; }
  if (c->dim_mode < 1 || c->dim_mode > 3) { *why = "QuadEnv: Unknown dimensionality mode"; return QS_EINVAL; }   // quadrotor.py:131
  if (c->init_random_state < QS_INIT_DEFAULT || c->init_random_state > QS_INIT_PITCH_ROLL_RESTRICTED) { *why = "unknown init_random_state (QsInitMode)"; return QS_EINVAL; }
  if (c->init_random_state == QS_INIT_PITCH_ROLL_RESTRICTED && c->dim_mode != 3) { *why = "QS_INIT_PITCH_ROLL_RESTRICTED is a 3D law"; return QS_EINVAL; }
  if (c->sim_steps < 1 || c->sim_steps > 64) { *why = "sim_steps must be in [1, 64]"; return QS_EINVAL; }
  if (c->ep_len < 0 || c->ep_len > (1 << 23)) { *why = "ep_len out of range"; return QS_EINVAL; }
  if (!(c->dt > 0.0)) { *why = "dt must be positive"; return QS_EINVAL; }
  if (c->n_agents < 1 || c->n_agents > 32 || !is_pow2(c->n_agents)) { *why = "n_agents must be 1, 2, 4, 8, 16 or 32"; return QS_EINVAL; }
  if (c->n_envs % c->n_agents != 0) { *why = "n_envs must be a multiple of n_agents"; return QS_EINVAL; }
  if ((c->resample_goal || c->excite) && !c->per_env_goal) { *why = "resample_goal / excite need per_env_goal (a goal row)"; return QS_EINVAL; }
  if (c->use_formation && c->per_env_goal) { *why = "use_formation and per_env_goal are mutually exclusive"; return QS_EINVAL; }
  if (!(c->model.mass > 0.0) || !(c->model.inertia[0] > 0.0) || !(c->model.inertia[1] > 0.0) || !(c->model.inertia[2] > 0.0)) {
    *why = "model mass/inertia must be positive"; return QS_EINVAL;
  }
  return QS_OK;
}

bool needs_prev_action(const QsConfig* c) {
  return c->obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_ACC_ACT || c->obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_ACT ||
         c->reward.action_change != 0.0;
}

void layout_of(const QsConfig* c, QsStateLayout* o) {
  int rows = 7;
  o->row_ou = (c->model.ou_sigma > 0.0 || c->per_env_model) ? rows++ : -1;
  o->row_prev_action = needs_prev_action(c) ? rows++ : -1;
  o->row_gyro_bias = (c->sense.enabled && c->sense.gyro_norm_std != 0.0) ? rows++ : -1;
  o->row_goal = c->per_env_goal ? rows++ : -1;
  o->rows = rows;
  o->obs_dim = kObsDims[c->obs_repr];
  o->act_dim = kActDims[c->controller];
  o->_pad = 0;
}

// which rarely used feature groups (qs::Feat) the configuration needs; 0 = the specialised fast subset
int feat_of(const QsHandle* h) {
  const QsConfig& c = h->cfg;
  const QsModel& m = c.model;
  const QsSensorNoise& s = c.sense;
  int f = qs::F_NONE;
  if (m.linearity != 1.0 || m.C_drag != 0.0 || m.C_roll != 0.0 || m.vel_damp != 0.0 || m.damp_omega_quadratic != 0.0) f |= qs::F_DRAG;
  if (s.enabled && (s.pos_unif_range != 0.0 || s.vel_unif_range != 0.0 || s.quat_norm_std != 0.0 || s.quat_unif_range != 0.0 ||
                    s.gyro_norm_std != 0.0)) f |= qs::F_NOISEX;
  if (c.dim_mode != 3 || c.per_env_goal) f |= qs::F_MISC;
  if (h->info != nullptr) f |= qs::F_INFO;
  if (c.obs_repr >= QS_OBS_XYZ_VXYZ_QUAT_OMEGA || c.init_random_state == QS_INIT_PITCH_ROLL_RESTRICTED) f |= qs::F_MISC;
  // swarms (and the log-distance reward that comes with them) and per-env models have specialised kernels for the
  // common step shape only
  const bool plain = c.sim_steps == 2 && h->lay.obs_dim == 18 && h->lay.row_prev_action < 0 &&
                     c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA;
  if ((c.n_agents > 1 || c.reward.log_pos || c.per_env_model) && !plain) f |= qs::F_MISC;
  if (c.per_env_model && (c.reward.log_pos || c.n_agents > 1)) f |= qs::F_MISC;   // per-env models in swarms: the catch-all twin
  // the specialised per-env-model kernels compile out linearity / rotor drag / damping (rows 8 and 9 of the model buffer):
  // they are taken only when the caller has asserted that every bound model is plain in that sense
  if (c.per_env_model && !h->models_plain) f |= qs::F_DRAG;
  // the contact stage, formation goals and the log-distance reward live in the specialised swarm kernels (fast subset) and in
  // F_MISC: a swarm that needs any other feature group takes F_MISC along
  if ((c.n_agents > 1 || c.reward.log_pos) && f != qs::F_NONE) f |= qs::F_MISC;
  return f;
}

template <typename R>
void fill_params(const QsHandle* h, qs::Params<R>* pp) {
  qs::Params<R>& p = *pp;
  memset(&p, 0, sizeof(p));
  const QsConfig& c = h->cfg;
  const QsModel& m = c.model;
  p.state = (R*)h->state;
  p.info = (R*)h->info;
  p.models = (const R*)h->models;
  p.term_obs = (R*)h->term_obs;
#ifdef QS_PHASE_TIMING
  p.timing = h->timing;
#endif
  p.ou_z = (const R*)h->ou_z;
  p.obs_z = (const R*)h->obs_z;
  p.reset_z = (const R*)h->reset_z;
  p.goal_z = (const R*)h->goal_z;
  p.resample_goal = c.resample_goal ? 1 : 0; p.excite = c.excite ? 1 : 0;
  p.stats = h->d_stats;
  p.status = h->d_status;
  p.n = c.n_envs;
  p.begin = 0; p.end = c.n_envs;
  p.env_id_offset = c.env_id_offset;
  p.seed_lo = (unsigned)(c.seed & 0xffffffffu);
  p.seed_hi = (unsigned)(c.seed >> 32);
  p.clock = h->d_clock; p.clock_ticket = h->d_ticket; p.clock_advance = 1;
  p.launch_ticket = h->d_ticket + 1; p.tile_counter = h->d_ticket + 2;   // launch slot 0
  p.svd_period = c.svd_period;
  p.dt_d = c.dt;
  p.row_ou = h->lay.row_ou; p.row_act = h->lay.row_prev_action; p.row_gyro = h->lay.row_gyro_bias; p.row_goal = h->lay.row_goal;
  p.obs_dim = h->lay.obs_dim; p.act_dim = h->lay.act_dim;
  p.k_steps = 1; p.sim_steps = c.sim_steps; p.ep_len = c.ep_len;
  p.controller = c.controller; p.zero_middle = c.zero_middle; p.n_agents = c.n_agents; p.dim_mode = c.dim_mode;
  p.obs_rel = (c.obs_repr == QS_OBS_XYZR_VXYZR_R_OMEGA || c.obs_repr == QS_OBS_XYZR_VXYZR_R_OMEGA_H ||
               c.obs_repr == QS_OBS_XYZR_VXYZR_QUAT_OMEGA || c.obs_repr == QS_OBS_XYZR_VXYZR_QUAT_OMEGA_H ||
               c.obs_repr == QS_OBS_XYZR_VXYZR_R_OMEGA_T2W);
  p.obs_h = (c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_H || c.obs_repr == QS_OBS_XYZR_VXYZR_R_OMEGA_H ||
             c.obs_repr == QS_OBS_XYZR_VXYZR_QUAT_OMEGA_H);
  p.obs_quat = (c.obs_repr == QS_OBS_XYZ_VXYZ_QUAT_OMEGA || c.obs_repr == QS_OBS_XYZR_VXYZR_QUAT_OMEGA ||
                c.obs_repr == QS_OBS_XYZR_VXYZR_QUAT_OMEGA_H);
  p.obs_t2w = (c.obs_repr == QS_OBS_XYZR_VXYZR_R_OMEGA_T2W || c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_T2W ||
               c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_T2W_T2T);
  p.obs_t2t = (c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_T2W_T2T);
  p.obs_acc = (c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_ACC_ACT);
  p.obs_act = (c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_ACC_ACT || c.obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_ACT);
  p.init_random = c.init_random_state; p.auto_reset = c.auto_reset;
  p.sense_on = c.sense.enabled ? 1 : 0;
  p.ou_on = (m.ou_sigma > 0.0 || c.per_env_model) ? 1 : 0;
  p.gyro_bias_mode = (c.sense.enabled && c.sense.gyro_norm_std != 0.0) ? 1 : 0;
  p.dyn_controller = -1; p.dyn_substeps = c.sim_steps;
  p.inject = (c.noise_mode == QS_NOISE_INJECTED) ? 1 : 0;
  p.feat = feat_of(h);
  p.tma_generic = h->tma_generic;
  p.tma_rows = h->lay.rows;
  p.row_bytes = (long long)c.n_envs * 4 * (long long)sizeof(R);
  p.pdl = h->pdl;
  p.ow_threads = h->ow_threads; p.ow_extra = h->ow_extra; p.prefetch = 0;
#ifdef QS_DEBUG_BOUNDS
  { const char* e = getenv("QUADSIM_BOUNDS_SELFTEST"); p.bounds_selftest = e ? atoi(e) : 0; }
#endif
  p.spec_draw = h->spec_draw;
  // model (quadrotor.py:142-208, :284-288)
  p.dt = (R)c.dt;
  const double eps = 1e-6;
  double tau_up = 4.0 * c.dt / (m.damp_time_up + eps), tau_down = 4.0 * c.dt / (m.damp_time_down + eps);
  if (tau_up > 1.0) tau_up = 1.0;
  if (tau_down > 1.0) tau_down = 1.0;
  for (int k = 0; k < 4; ++k) {
    p.thrust_max[k] = (R)m.thrust_max[k];
    p.tz[k] = (R)(m.torque_max[k] * m.prop_ccw[k]);
    p.pcx[k] = (R)(m.prop_pos[k][1]);    // prop_pos x z_hat = (y, -x, 0)   (quadrotor.py:182)
    p.pcy[k] = (R)(-m.prop_pos[k][0]);
    p.ppx[k] = (R)m.prop_pos[k][0]; p.ppy[k] = (R)m.prop_pos[k][1]; p.ppz[k] = (R)m.prop_pos[k][2];
    p.ccw[k] = (R)m.prop_ccw[k];
    for (int j = 0; j < 4; ++j) p.jinv[k][j] = (R)m.Jinv[k][j];
  }
  for (int k = 0; k < 3; ++k) {
    p.inertia[k] = (R)m.inertia[k];
    p.inv_inertia[k] = (R)(1.0 / m.inertia[k]);
    p.room_lo[k] = (R)c.room_box[0][k]; p.room_hi[k] = (R)c.room_box[1][k];
    p.goal[k] = (R)c.goal[k];
  }
  p.mass = (R)m.mass; p.inv_mass = (R)(1.0 / m.mass);
  p.lin = (R)m.linearity; p.c_drag = (R)m.C_drag; p.c_roll = (R)m.C_roll;
  p.tau_up = (R)tau_up; p.tau_down = (R)tau_down;
  p.vel_damp = (R)m.vel_damp; p.damp_omega_q = (R)m.damp_omega_quadratic;
  p.ou_theta = (R)m.ou_theta; p.ou_sigma = (R)m.ou_sigma;
  p.arm = (R)m.arm; p.omega_max = (R)m.omega_max; p.gravity = (R)m.gravity; p.t2w = (R)m.thrust_to_weight;
  p.t2t = (R)m.torque_to_thrust; p.t2w_half_std = (R)(c.t2w_std / 2.0); p.t2t_half_std = (R)(c.t2t_std / 2.0);
  p.pitch_max = (R)c.pitch_max; p.roll_max = (R)c.roll_max; p.yaw_max = (R)c.yaw_max;
  p.spawn_box = (R)c.spawn_box; p.spawn_z_min = (R)c.spawn_z_min; p.max_init_vel = (R)c.max_init_vel; p.max_init_omega = (R)c.max_init_omega;
  p.coll_dist = (R)(c.collision_k * 2.0 * m.arm); p.coll_k = (R)c.collision_k;
  p.formation_on = (c.use_formation && c.n_agents > 1) ? 1 : 0;
  for (int a = 0; a < 32; ++a)
    for (int k = 0; k < 3; ++k) p.formation[a][k] = (R)c.formation[a][k];
  const QsReward& r = c.reward;
  p.c_pos = (R)r.pos; p.c_effort = (R)r.effort; p.c_act_change = (R)r.action_change; p.c_crash = (R)r.crash;
  p.c_orient = (R)r.orient; p.c_yaw = (R)r.yaw; p.c_rot = (R)r.rot; p.c_attitude = (R)r.attitude; p.c_spin = (R)r.spin;
  p.c_vel = (R)r.vel; p.c_pos_offset = (R)r.pos_offset; p.c_pos_log_w = (R)r.pos_log_weight; p.c_pos_lin_w = (R)r.pos_linear_weight;
  p.c_collision = (R)r.collision; p.log_pos = r.log_pos;
  const QsSensorNoise& s = c.sense;
  p.sn_pos_std = (R)s.pos_norm_std; p.sn_pos_unif = (R)s.pos_unif_range; p.sn_vel_std = (R)s.vel_norm_std;
  p.sn_vel_unif = (R)s.vel_unif_range; p.sn_quat_std = (R)s.quat_norm_std; p.sn_quat_unif = (R)s.quat_unif_range;
  p.sn_gyro_density = (R)s.gyro_noise_density; p.sn_gyro_rw = (R)s.gyro_random_walk;
  p.sn_acc_static = (R)s.acc_static_noise_std; p.sn_acc_dyn = (R)s.acc_dynamic_noise_ratio;
  if (p.gyro_bias_mode) {   // sensor_noise.py:165-167, the walk folded over n add_noise calls
    double T = s.gyro_bias_correlation_time;
    double sg = s.gyro_noise_density / sqrt(c.dt);
    double sb = sqrt(-(sg * sg) * (T / 2.0) * (exp(-2.0 * c.dt / T) - 1.0));
    double pi1 = exp(-c.dt / T);
    p.gb_pi1 = (R)pi1; p.gb_sig1 = (R)sb;
    p.gb_pi3 = (R)(pi1 * pi1 * pi1);
    p.gb_sig3 = (R)(sb * sqrt(1.0 + pi1 * pi1 + pi1 * pi1 * pi1 * pi1));
  }
}

template <typename R>
int do_step(QsHandle* h, int k, const void* actions, void* obs, void* rew, uint8_t* done, cudaStream_t st,
            long long begin, long long end, int advance, int slot = 0, bool device_buffers = true) {
  qs::Params<R> p;
  fill_params<R>(h, &p);
  p.k_steps = k;
  if (device_buffers) p.prefetch = h->prefetch;   // the L2 prefetch hint is for device memory (not for pinned host aliases)
  else { p.host_io = 1; p.host_ctas = h->host_ctas; }
  p.clock_advance = advance;
  if (k > 1 && (h->cfg.n_agents > 1 || h->cfg.reward.log_pos || h->cfg.per_env_model)) p.feat |= qs::F_MISC;   // specialised kernels: k == 1 only
  if (!advance) p.pdl = 0;                       // chunk launches of the host pipeline follow copies on side streams
  p.launch_ticket = h->d_ticket + 1 + 2 * slot;
  p.tile_counter = h->d_ticket + 2 + 2 * slot;
  p.actions = (const R*)actions; p.obs = (R*)obs; p.rew = (R*)rew; p.done = done;
  p.begin = begin; p.end = end;
  cudaError_t e = qs::launch_step<R>(p, st, &h->last_variant);
  if (e != cudaSuccess) {
    // a launch that did not start leaves the ticket / tile counters untouched, but one that was rejected mid-pipeline
    // (chunked host path) must not leave a partial count behind: the clock would silently stop advancing
    cudaGetLastError();
    cudaMemsetAsync(h->d_ticket, 0, (1 + 2 * kHostStreams) * sizeof(unsigned int), st);
    return fail(h, QS_ECUDA, "env_step launch: %s", cudaGetErrorString(e));
  }
  h->launches++;
  return QS_OK;
}

int check_ready(QsHandle* h, bool need_actions, const void* actions) {
  if (!h) return QS_EINVAL;
  if (!h->state) return fail(h, QS_ESTATE, "state buffer not bound (qs_bind_state)");
  if (h->cfg.per_env_model && !h->models) return fail(h, QS_ESTATE, "model buffer not bound (qs_bind_models)");
  if (need_actions && !actions && h->cfg.controller != QS_CTRL_MELLINGER)
    return fail(h, QS_EINVAL, "actions may be NULL only for QS_CTRL_MELLINGER");
  return QS_OK;
}

}  // namespace

extern "C" {

int qs_abi_version(void) { return QS_ABI_VERSION; }

int qs_state_layout(const QsConfig* cfg, QsStateLayout* out) {
  std::string why;
  int rc = validate(cfg, &why);
  if (rc != QS_OK) { g_create_error = why; return rc; }
  if (!out) return QS_EINVAL;
  layout_of(cfg, out);
  return QS_OK;
}

int qs_create(const QsConfig* cfg, QsHandle** out) {
  std::string why;
  int rc = validate(cfg, &why);
  if (rc != QS_OK) { g_create_error = why; return rc; }
  if (!out) { g_create_error = "out is NULL"; return QS_EINVAL; }
  QsHandle* h = new (std::nothrow) QsHandle();
  if (!h) { g_create_error = "out of host memory"; return QS_EINVAL; }
  h->cfg = *cfg;
  layout_of(cfg, &h->lay);
  h->state = h->info = nullptr;
  h->models = nullptr;
  h->models_plain = 0;
  h->term_obs = nullptr;
  h->ou_z = h->obs_z = h->reset_z = nullptr;
  h->goal_z = nullptr;
  h->d_stats = nullptr; h->d_status = nullptr; h->d_clock = nullptr; h->d_ticket = nullptr;
  h->launches = 0;
  { const char* e = getenv("QUADSIM_PDL"); h->pdl = e ? atoi(e) : 1; }   // A/B knob: programmatic dependent launch, default on
  h->last_variant = "";
  h->host_streams_made = 0;
  h->order_ev = nullptr;
  { const char* e = getenv("QUADSIM_HOST_ZEROCOPY"); h->host_zero_copy = e ? atoi(e) : 1; }
  { const char* e = getenv("QUADSIM_TMA_GENERIC"); h->tma_generic = e ? atoi(e) : 1; }
  { const char* e = getenv("QUADSIM_OW_THREADS"); h->ow_threads = e ? atoi(e) : 64; }
  { const char* e = getenv("QUADSIM_OW_EXTRA"); h->ow_extra = e ? atoi(e) : 0; }
  { const char* e = getenv("QUADSIM_HOST_CTAS"); h->host_ctas = e ? atoi(e) : 2; }   // measured: 134 -> 123 us per 65,536-env step
  { const char* e = getenv("QUADSIM_PREFETCH"); h->prefetch = e ? atoi(e) : 3; }
  { const char* e = getenv("QUADSIM_SPEC_DRAW"); h->spec_draw = e ? atoi(e) : 1; }
#ifdef QS_PHASE_TIMING
  { const char* e = getenv("QUADSIM_TIMING_PTR"); h->timing = e ? (unsigned long long*)strtoull(e, nullptr, 0) : nullptr; }
#endif
  h->last_host_path = "";
  h->h_act = h->h_obs = h->h_rew = nullptr; h->h_done = nullptr;
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_stats, 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_status, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->d_stats, 0, 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(h->d_status, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_clock, sizeof(qs::Clock));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_ticket, (1 + 2 * kHostStreams) * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->d_clock, 0, sizeof(qs::Clock));
  if (e == cudaSuccess) e = cudaMemset(h->d_ticket, 0, (1 + 2 * kHostStreams) * sizeof(unsigned int));
  if (e != cudaSuccess) {
    g_create_error = std::string("qs_create: ") + cudaGetErrorString(e);
    if (h->d_stats) cudaFree(h->d_stats);
    if (h->d_status) cudaFree(h->d_status);
    if (h->d_clock) cudaFree(h->d_clock);
    if (h->d_ticket) cudaFree(h->d_ticket);
    delete h;
    return QS_ECUDA;
  }
  *out = h;
  return QS_OK;
}

int qs_destroy(QsHandle* h) {
  if (!h) return QS_EINVAL;
  cudaSetDevice(h->cfg.device);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  for (int k = 0; k < h->host_streams_made; ++k) cudaStreamDestroy(h->hs[k]);
  if (h->h_act) cudaFree(h->h_act);
  if (h->h_obs) cudaFree(h->h_obs);
  if (h->h_rew) cudaFree(h->h_rew);
  if (h->h_done) cudaFree(h->h_done);
  cudaFree(h->d_stats);
  cudaFree(h->d_status);
  cudaFree(h->d_clock);
  cudaFree(h->d_ticket);
  delete h;
  return QS_OK;
}

const char* qs_last_error(const QsHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int qs_bind_state(QsHandle* h, void* state) {
  if (!h) return QS_EINVAL;
  if (!state) return fail(h, QS_EINVAL, "state is NULL");
  if (((uintptr_t)state) % 32 != 0) return fail(h, QS_EINVAL, "state must be 32-byte aligned");
  h->state = state;
  return QS_OK;
}

int qs_bind_models(QsHandle* h, const void* models, int32_t flags) {
  if (!h) return QS_EINVAL;
  if (!h->cfg.per_env_model) return fail(h, QS_EINVAL, "handle was not created with per_env_model");
  if (!models) return fail(h, QS_EINVAL, "models is NULL");
  if (((uintptr_t)models) % 32 != 0) return fail(h, QS_EINVAL, "models must be 32-byte aligned");
  if (flags & ~QS_MODELS_PLAIN) return fail(h, QS_EINVAL, "unknown qs_bind_models flag");
  h->models = models;
  h->models_plain = (flags & QS_MODELS_PLAIN) ? 1 : 0;
  return QS_OK;
}

int qs_bind_terminal_obs(QsHandle* h, void* terminal_obs) {
  if (!h) return QS_EINVAL;
  h->term_obs = terminal_obs;
  return QS_OK;
}

int qs_bind_info(QsHandle* h, void* info) {
  if (!h) return QS_EINVAL;
  if (info && !vec_aligned(h, info)) return fail(h, QS_EINVAL, "info must be aligned to 4 reals (16 bytes f32 / 32 bytes f64): vector stores");
  h->info = info;
  return QS_OK;
}

int qs_set_noise_injection(QsHandle* h, const void* ou_z, const void* obs_z, const void* reset_z) {
  if (!h) return QS_EINVAL;
  if (h->cfg.noise_mode != QS_NOISE_INJECTED) return fail(h, QS_EINVAL, "handle was not created with QS_NOISE_INJECTED");
  h->ou_z = ou_z; h->obs_z = obs_z; h->reset_z = reset_z;
  return QS_OK;
}

int qs_set_goal_injection(QsHandle* h, const void* goal_z) {
  if (!h) return QS_EINVAL;
  if (h->cfg.noise_mode != QS_NOISE_INJECTED) return fail(h, QS_EINVAL, "handle was not created with QS_NOISE_INJECTED");
  h->goal_z = goal_z;
  return QS_OK;
}

int qs_reset(QsHandle* h, const uint8_t* mask, void* obs, qs_stream_t stream) {
  int rc = check_ready(h, false, nullptr);
  if (rc != QS_OK) return rc;
  if (!obs) return fail(h, QS_EINVAL, "obs is NULL");
  if (!vec_aligned(h, obs)) return fail(h, QS_EINVAL, "obs must be aligned to 4 reals (16 bytes f32 / 32 bytes f64)");
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaError_t e;
  if (h->cfg.dtype == QS_F32) {
    qs::Params<float> p; fill_params<float>(h, &p);
    p.reset_mask = mask; p.obs = (float*)obs; p.begin = 0; p.end = p.n;
    e = qs::launch_reset<float>(p, (cudaStream_t)stream);
  } else {
    qs::Params<double> p; fill_params<double>(h, &p);
    p.reset_mask = mask; p.obs = (double*)obs; p.begin = 0; p.end = p.n;
    e = qs::launch_reset<double>(p, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return fail(h, QS_ECUDA, "env_reset launch: %s", cudaGetErrorString(e));
  h->launches++;
  return QS_OK;
}

int qs_observe(QsHandle* h, const void* accelerometer, void* obs, qs_stream_t stream) {
  int rc = check_ready(h, false, nullptr);
  if (rc != QS_OK) return rc;
  if (!obs) return fail(h, QS_EINVAL, "obs is NULL");
  if (!vec_aligned(h, obs)) return fail(h, QS_EINVAL, "obs must be aligned to 4 reals (16 bytes f32 / 32 bytes f64)");
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaError_t e;
  if (h->cfg.dtype == QS_F32) {
    qs::Params<float> p; fill_params<float>(h, &p);
    p.acc_in = (const float*)accelerometer; p.obs = (float*)obs; p.begin = 0; p.end = p.n;
    e = qs::launch_observe<float>(p, (cudaStream_t)stream);
  } else {
    qs::Params<double> p; fill_params<double>(h, &p);
    p.acc_in = (const double*)accelerometer; p.obs = (double*)obs; p.begin = 0; p.end = p.n;
    e = qs::launch_observe<double>(p, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return fail(h, QS_ECUDA, "env_observe launch: %s", cudaGetErrorString(e));
  h->launches++;
  return QS_OK;
}

int qs_step_k(QsHandle* h, int k, const void* actions, void* obs, void* rew, uint8_t* done, qs_stream_t stream) {
  int rc = check_ready(h, true, actions);
  if (rc != QS_OK) return rc;
  if (k < 1) return fail(h, QS_EINVAL, "k must be >= 1");
  if (!obs || !rew || !done) return fail(h, QS_EINVAL, "obs/rew/done must not be NULL");
  if (!vec_aligned(h, obs) || (actions && !vec_aligned(h, actions)))
    return fail(h, QS_EINVAL, "actions and obs must be aligned to 4 reals (16 bytes f32 / 32 bytes f64): vector loads, bulk stores");
  if (k > 1 && h->cfg.noise_mode == QS_NOISE_INJECTED) return fail(h, QS_EUNSUPPORTED, "qs_step_k with injected noise needs k == 1");
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  if (h->cfg.dtype == QS_F32) rc = do_step<float>(h, k, actions, obs, rew, done, (cudaStream_t)stream, 0, h->cfg.n_envs, 1);
  else rc = do_step<double>(h, k, actions, obs, rew, done, (cudaStream_t)stream, 0, h->cfg.n_envs, 1);
  return rc;
}

int qs_step(QsHandle* h, const void* actions, void* obs, void* rew, uint8_t* done, qs_stream_t stream) {
  return qs_step_k(h, 1, actions, obs, rew, done, stream);
}

int qs_dynamics_step(QsHandle* h, int controller, const void* actions, int n_substeps, qs_stream_t stream) {
  int rc = check_ready(h, false, nullptr);
  if (rc != QS_OK) return rc;
  if (controller < -1 || controller >= QS_CTRL_COUNT) return fail(h, QS_EINVAL, "unknown controller");
  if (n_substeps < 0) return fail(h, QS_EINVAL, "n_substeps must be >= 0");
  if (!actions && controller != QS_CTRL_MELLINGER) return fail(h, QS_EINVAL, "actions is NULL");
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaError_t e;
  if (h->cfg.dtype == QS_F32) {
    qs::Params<float> p; fill_params<float>(h, &p);
    p.actions = (const float*)actions; p.dyn_controller = controller; p.dyn_substeps = n_substeps;
    p.act_dim = controller < 0 ? 4 : kActDims[controller]; p.begin = 0; p.end = p.n;
    e = qs::launch_dynamics<float>(p, (cudaStream_t)stream);
  } else {
    qs::Params<double> p; fill_params<double>(h, &p);
    p.actions = (const double*)actions; p.dyn_controller = controller; p.dyn_substeps = n_substeps;
    p.act_dim = controller < 0 ? 4 : kActDims[controller]; p.begin = 0; p.end = p.n;
    e = qs::launch_dynamics<double>(p, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return fail(h, QS_ECUDA, "dynamics_step launch: %s", cudaGetErrorString(e));
  h->launches++;
  return QS_OK;
}

int qs_affine_fg(QsHandle* h, const void* s, double dt, void* F, void* G, int64_t n, qs_stream_t stream) {
  if (!h) return QS_EINVAL;
  if (!s) return fail(h, QS_EINVAL, "s is NULL");
  if (n < 0) return fail(h, QS_EINVAL, "n must be >= 0");
  if (!(dt > 0.0)) return fail(h, QS_EINVAL, "dt must be positive");
  if (h->cfg.per_env_model && !h->models) return fail(h, QS_ESTATE, "model buffer not bound (qs_bind_models)");
  if (n == 0 || (!F && !G)) return QS_OK;
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  const int C = h->cfg.dim_mode == 1 ? 1 : (h->cfg.dim_mode == 2 ? 2 : 4);           // control_mx, quadrotor.py:121-129
  cudaError_t e;
  if (h->cfg.dtype == QS_F32) {
    qs::Params<float> p; fill_params<float>(h, &p);
    qs::AffineArgs<float> a = {(const float*)s, (float*)F, (float*)G, (long long)n, (float)dt, C};
    e = qs::launch_affine<float>(p, a, (cudaStream_t)stream);
  } else {
    qs::Params<double> p; fill_params<double>(h, &p);
    qs::AffineArgs<double> a = {(const double*)s, (double*)F, (double*)G, (long long)n, dt, C};
    e = qs::launch_affine<double>(p, a, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return fail(h, QS_ECUDA, "affine_fg launch: %s", cudaGetErrorString(e));
  h->launches++;
  return QS_OK;
}

int qs_step_host(QsHandle* h, const void* actions_host, void* obs_host, void* rew_host, uint8_t* done_host, qs_stream_t stream) {
  int rc = check_ready(h, true, actions_host);
  if (rc != QS_OK) return rc;
  if (!obs_host || !rew_host || !done_host) return fail(h, QS_EINVAL, "obs/rew/done must not be NULL");
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t user = (cudaStream_t)stream;
  const long long n = h->cfg.n_envs;
  const size_t rs = h->cfg.dtype == QS_F32 ? sizeof(float) : sizeof(double);
  const int od = h->lay.obs_dim, ad = h->lay.act_dim;
  // Pinned (cudaHostAlloc / cudaHostRegister) buffers are addressable from the device: one launch ON THE CALLER'S STREAM
  // (so it is ordered after the resets, steps and state writes the caller queued there) reads the actions and writes
  // obs / reward / done straight across PCIe (coalesced 16-byte vectors), so there is no staging copy, no copy-engine
  // hand-off and a single stream synchronisation per step.
  if (h->host_zero_copy) {
    void* d_act = nullptr; void* d_obs = nullptr; void* d_rew = nullptr; void* d_done = nullptr;
    if ((!actions_host || device_alias(h, actions_host, &d_act, true)) && device_alias(h, obs_host, &d_obs, true) &&
        device_alias(h, rew_host, &d_rew, false) && device_alias(h, done_host, &d_done, false)) {
      if (h->cfg.dtype == QS_F32) rc = do_step<float>(h, 1, d_act, d_obs, d_rew, (uint8_t*)d_done, user, 0, n, 1, 0, false);
      else rc = do_step<double>(h, 1, d_act, d_obs, d_rew, (uint8_t*)d_done, user, 0, n, 1, 0, false);
      if (rc != QS_OK) return rc;
      QS_CUDA(h, cudaStreamSynchronize(user));
      h->last_host_path = "zero-copy";
      return QS_OK;
    }
  }
  h->last_host_path = "staged";
  // resources are created one by one, so a setup that failed half-way resumes instead of leaking what it had made
  while (h->host_streams_made < kHostStreams) {
    QS_CUDA(h, cudaStreamCreateWithFlags(&h->hs[h->host_streams_made], cudaStreamNonBlocking));
    h->host_streams_made++;
  }
  if (!h->order_ev) QS_CUDA(h, cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
  if (!h->h_act) QS_CUDA(h, cudaMalloc(&h->h_act, (size_t)n * ad * rs));
  if (!h->h_obs) QS_CUDA(h, cudaMalloc(&h->h_obs, (size_t)n * od * rs));
  if (!h->h_rew) QS_CUDA(h, cudaMalloc(&h->h_rew, (size_t)n * rs));
  if (!h->h_done) QS_CUDA(h, cudaMalloc((void**)&h->h_done, (size_t)n));
  // the internal streams start after everything the caller has queued on its stream so far
  QS_CUDA(h, cudaEventRecord(h->order_ev, user));
  // chunks of whole CTAs (and whole warps, so swarm groups never straddle a chunk), one stream each:
  // H2D(actions) -> fused step -> D2H(obs, rew, done); copies of one chunk overlap the kernel of another
  long long per = (n + kHostStreams - 1) / kHostStreams;
  per = ((per + qs::BLOCK - 1) / qs::BLOCK) * qs::BLOCK;
  int used = 0;
  for (int k = 0; k < kHostStreams; ++k) {
    long long b = (long long)k * per, e = b + per < n ? b + per : n;
    if (b >= n) break;
    used = k + 1;
    cudaStream_t st = h->hs[k];
    QS_CUDA(h, cudaStreamWaitEvent(st, h->order_ev, 0));
    if (actions_host)
      QS_CUDA(h, cudaMemcpyAsync((char*)h->h_act + (size_t)b * ad * rs, (const char*)actions_host + (size_t)b * ad * rs,
                                 (size_t)(e - b) * ad * rs, cudaMemcpyHostToDevice, st));
    const void* act_dev = actions_host ? h->h_act : nullptr;
    if (h->cfg.dtype == QS_F32) rc = do_step<float>(h, 1, act_dev, h->h_obs, h->h_rew, h->h_done, st, b, e, 0, k);
    else rc = do_step<double>(h, 1, act_dev, h->h_obs, h->h_rew, h->h_done, st, b, e, 0, k);
    if (rc != QS_OK) return rc;
    QS_CUDA(h, cudaMemcpyAsync((char*)obs_host + (size_t)b * od * rs, (char*)h->h_obs + (size_t)b * od * rs,
                               (size_t)(e - b) * od * rs, cudaMemcpyDeviceToHost, st));
    QS_CUDA(h, cudaMemcpyAsync((char*)rew_host + (size_t)b * rs, (char*)h->h_rew + (size_t)b * rs, (size_t)(e - b) * rs,
                               cudaMemcpyDeviceToHost, st));
    QS_CUDA(h, cudaMemcpyAsync(done_host + b, h->h_done + b, (size_t)(e - b), cudaMemcpyDeviceToHost, st));
  }
  for (int k = 0; k < used; ++k) QS_CUDA(h, cudaStreamSynchronize(h->hs[k]));
  // the chunk launches only read the clock; advance it once for the whole step (on the caller's stream: every chunk has
  // completed, and whatever the caller queues next is ordered after the advance)
  cudaError_t ce;
  if (h->cfg.dtype == QS_F32) { qs::Params<float> p; fill_params<float>(h, &p); ce = qs::launch_clock_advance<float>(p, user); }
  else { qs::Params<double> p; fill_params<double>(h, &p); ce = qs::launch_clock_advance<double>(p, user); }
  if (ce != cudaSuccess) return fail(h, QS_ECUDA, "clock_advance launch: %s", cudaGetErrorString(ce));
  h->launches++;
  QS_CUDA(h, cudaStreamSynchronize(user));
  return QS_OK;
}

int qs_get_stats(QsHandle* h, double out[3], int clear, qs_stream_t stream) {
  if (!h || !out) return QS_EINVAL;
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  QS_CUDA(h, cudaMemcpyAsync(out, h->d_stats, 3 * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (clear) QS_CUDA(h, cudaMemsetAsync(h->d_stats, 0, 3 * sizeof(double), (cudaStream_t)stream));
  QS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return QS_OK;
}

int qs_stats_device_ptr(QsHandle* h, double** out) {
  if (!h || !out) return QS_EINVAL;
  *out = h->d_stats;
  return QS_OK;
}

int qs_check(QsHandle* h, qs_stream_t stream) {
  if (!h) return QS_EINVAL;
  int flag = 0;
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  QS_CUDA(h, cudaMemcpyAsync(&flag, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  QS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  if (flag & 2) {            // QS_DEBUG_BOUNDS builds only: a kernel-side index / pointer check failed (qs::BoundsCode mask)
    char code[32];
    snprintf(code, sizeof(code), "0x%x", (unsigned)flag >> 8);
    return fail(h, QS_ECUDA, "QS_DEBUG_BOUNDS: bounds check failed, code mask %s", code);
  }
  if (flag) return fail(h, QS_ENONFINITE, "QuadEnv: reward is Nan");   // quadrotor.py:636
  return QS_OK;
}

int qs_get_clock(QsHandle* h, double* svd_timer, uint64_t* step_count, qs_stream_t stream) {
  if (!h) return QS_EINVAL;
  qs::Clock c;
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  QS_CUDA(h, cudaMemcpyAsync(&c, h->d_clock, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  QS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  if (svd_timer) *svd_timer = c.svd_timer;
  if (step_count) *step_count = c.step;
  return QS_OK;
}

int qs_set_clock(QsHandle* h, double svd_timer, uint64_t step_count, qs_stream_t stream) {
  if (!h) return QS_EINVAL;
  qs::Clock c;
  c.svd_timer = svd_timer;
  c.step = step_count;
  QS_CUDA(h, cudaSetDevice(h->cfg.device));
  QS_CUDA(h, cudaMemcpyAsync(h->d_clock, &c, sizeof(c), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  QS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return QS_OK;
}

int qs_launch_count(const QsHandle* h, uint64_t* out) {
  if (!h || !out) return QS_EINVAL;
  *out = h->launches;
  return QS_OK;
}

const char* qs_last_variant(const QsHandle* h) { return h ? h->last_variant : ""; }

const char* qs_last_host_path(const QsHandle* h) { return h ? h->last_host_path : ""; }

}  // extern "C"
