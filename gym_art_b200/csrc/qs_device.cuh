// qs_device.cuh - device side of libquadsim: one thread per env (drone), the whole env step fused in registers.
//
// Reference semantics (amolchanov86/gym_art, paths under gym_art/quadrotor/):
//   QuadrotorEnv._step            quadrotor.py:942-1028      -> env_step_kernel
//   QuadrotorDynamics.step1       quadrotor.py:273-436       -> substep()
//   controllers                   quadrotor_control.py       -> controller_cmds()
//   compute_reward_weighted       quadrotor.py:544-638       -> compute_reward()
//   SensorNoise.add_noise         sensor_noise.py:100-159    -> add_noise()
//   state_<obs_repr>              get_state.py:5-260         -> build_obs()
//   QuadrotorEnv._reset           quadrotor.py:1059-1144     -> reset_env()
// Written for sm_100a only.  No tensor cores: the per-env maths is 3x3 / 4x4, not a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qs {

constexpr int BLOCK = 128;           // threads per CTA = envs per CTA
constexpr int MAX_OBS = 25;
constexpr int OBS_Z = 33, RESET_Z = 51, INFO_DIM = 48, MODEL_ROWS = 14;
constexpr int OBS_SCRATCH = 32;     // words build_obs leaves behind: 18 core | h | acc3 | act4 | quat4 | t2w | t2t
enum InitMode { INIT_DEFAULT = 0, INIT_RANDOM = 1, INIT_PITCH_ROLL_RESTRICTED = 2 };   // QsInitMode

enum Ctrl { CTRL_RAW = 0, CTRL_VERTICAL, CTRL_VERTPLANE, CTRL_SHIFTED, CTRL_OMEGA_THRUST, CTRL_VELOCITY_YAW, CTRL_MELLINGER };

// Global bookkeeping that lives on the device so that a sequence of steps can be captured in a CUDA graph and
// replayed: the re-orthonormalisation timer (quadrotor.py:104, :381-386 - a global phase that reset() never
// touches) and the Philox step index.
struct Clock {
  double svd_timer;
  unsigned long long step;
};

// QS_DEBUG_BOUNDS build (tools/build_bounds.sh -> _exp/libquadsim_bounds.so; compute-sanitizer is not available on the GPU
// pool): every index / pointer the kernels derive - env ranges, optional state rows, the bulk-copy ring, the observation
// tiles, the tile scheduler, tickets - is checked against its bound BEFORE the access.  A failed check does not trap (the
// access that follows might): it ORs bit 1 and its code into the handle's status word, which qs_check reports.
#ifdef QS_DEBUG_BOUNDS
#define QS_CHECK(p, cond, code) do { if (!(cond)) atomicOr((p).status, 2 | ((code) << 8)); } while (0)
__device__ __forceinline__ unsigned int dyn_smem_bytes() { unsigned int v; asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v)); return v; }
#else
#define QS_CHECK(p, cond, code) do { } while (0)
#endif
enum BoundsCode { BC_ENV = 1, BC_ROW = 2, BC_MODEL = 4, BC_RING = 8, BC_SRC = 16, BC_OBS_SMEM = 32, BC_OBS_GLOBAL = 64, BC_TILE = 128,
                  BC_TICKET = 256, BC_INFO = 512, BC_ACTION = 1024, BC_SELFTEST = 2048 };

// ---------------------------------------------------------------------------------------------------------
// kernel parameter block (passed by value as a __grid_constant__; lives in the constant bank)
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct Params {
  // buffers
  R* state;                 // [rows][n][4]
  const R* actions;         // [k][n][act_dim] or null
  R* obs;                   // [n][obs_dim]
  R* rew;                   // [n]
  uint8_t* done;            // [n]
  R* info;                  // [n][INFO_DIM] or null
  R* term_obs;              // [n][obs_dim] or null: observation of the terminal state of envs that auto-reset this step
  const R* models;          // per-env model constants [MODEL_ROWS][n][4] or null (quadsim.h "Per-env model buffer")
  const uint8_t* reset_mask;
  const R* ou_z;            // injected: [n][sim_steps][4]
  const R* obs_z;           // injected: [n][OBS_Z]
  const R* reset_z;         // injected: [n][RESET_Z]
  const R* goal_z;          // injected: [n][4] moving-goal draws (excite xyz, resample_goal z)
  const R* acc_in;          // qs_observe: clean accelerometer readings [n][3] or null
  unsigned long long* timing;   // QS_PHASE_TIMING experiment builds only
  double* stats;            // [3]
  int* status;              // sticky non-finite flag
  long long n;              // envs on the handle = row stride of the state buffer
  long long begin, end;     // env range [begin, end) this launch covers (chunked host pipeline)
  long long env_id_offset;
  unsigned int seed_lo, seed_hi;
  // clocks
  double svd_period, dt_d;
  Clock* clock;             // device-resident global clock (read by every CTA, advanced by the last one)
  unsigned int* clock_ticket;
  unsigned int* tile_counter;  // dynamic tile scheduler of the TMA kernel (zero between launches)
  unsigned int* launch_ticket; // per-launch-slot ticket (the host pipeline runs several chunk launches concurrently)
  int clock_advance;        // 0: read only (chunked host pipeline advances it with clock_advance_kernel)
  // layout / config
  int row_ou, row_act, row_gyro, row_goal;
  int obs_dim, act_dim;
  int k_steps, sim_steps, ep_len;
  int controller, zero_middle, n_agents, dim_mode;
  int obs_rel, obs_h, obs_acc, obs_act;
  int obs_quat, obs_t2w, obs_t2t;     // get_state.py:276-382 (quaternion instead of R; thrust-to-weight / torque-to-thrust tail)
  int init_random /* InitMode */, auto_reset, sense_on, ou_on, gyro_bias_mode;
  int resample_goal, excite;          // moving goals (quadrotor.py:1078-1081, :957-962); the goal row is then written back
  int dyn_controller, dyn_substeps;   // qs_dynamics_step
  int inject, feat;                   // variant selection (host): noise source, Feat mask the configuration needs
  int tma_generic;                    // fp32 generic configuration through the TMA harness (A/B knob QUADSIM_TMA_GENERIC)
  int tma_rows;                       // rows of the state buffer (fast subset: 7 + OU + previous action)
  long long row_bytes;                // byte stride between state rows = n * 4 * sizeof(real)
  int pdl;                            // step kernel launched with programmatic dependent launch (see env_step_tma_kernel)
  int ow_threads;                     // threads per CTA of a one-wave launch (32, 64 or 128)
  int obs_stages;                     // pipelined launch: observation tiles in shared memory (2, or 1 when that buys a resident CTA)
  int ow_extra;                       // A/B knob QUADSIM_OW_EXTRA: CTA slots per SM a one-wave launch leaves to its dependent launch
  int host_io;                        // the launch reads / writes pinned HOST buffers across PCIe (qs_step_host zero-copy)
  int host_ctas;                      // host_io launches: resident CTAs per SM (0 = as many as fit)
  int prefetch;                       // L2 prefetch of the CTA's first tile ahead of griddepcontrol.wait: 0 off, 1 bulk, 3 lines
  int spec_draw;                      // draw the step's noise ahead of griddepcontrol.wait from a speculative clock read
  int bounds_selftest;                // QS_DEBUG_BOUNDS builds: raise BC_SELFTEST (proves the reporting path end to end)
  // model
  R dt;
  R thrust_max[4], tz[4], pcx[4], pcy[4], ppx[4], ppy[4], ppz[4], ccw[4];
  R inertia[3], inv_inertia[3];
  R mass, inv_mass, lin, c_drag, c_roll, tau_up, tau_down, vel_damp, damp_omega_q;
  R ou_theta, ou_sigma, arm, omega_max, gravity, t2w, t2t;
  R t2w_half_std, t2t_half_std;       // t2w_std / 2, t2t_std / 2 (get_state.py:334, :379)
  R pitch_max, roll_max, yaw_max;     // INIT_PITCH_ROLL_RESTRICTED (quadrotor.py:243)
  R jinv[4][4];
  R room_lo[3], room_hi[3];
  R goal[3], spawn_box, spawn_z_min, max_init_vel, max_init_omega, coll_dist, coll_k;
  int formation_on;
  R formation[32][3];                  // swarm: goal of drone a of every swarm
  // reward
  R c_pos, c_effort, c_act_change, c_crash, c_orient, c_yaw, c_rot, c_attitude, c_spin, c_vel;
  R c_pos_offset, c_pos_log_w, c_pos_lin_w, c_collision;
  int log_pos;
  // sensor noise
  R sn_pos_std, sn_pos_unif, sn_vel_std, sn_vel_unif, sn_quat_std, sn_quat_unif;
  R sn_gyro_density, sn_gyro_rw, sn_acc_static, sn_acc_dyn;
  R gb_pi1, gb_sig1, gb_pi3, gb_sig3;   // gyro-bias walk folded over 1 (reset) / 3 (step) add_noise calls
};

// compile-time specialisation knobs; -1 = decided at run time from Params.
// FEAT: which groups of rarely used features the instantiation carries (the host computes the mask a configuration needs,
// feat_of() in quadsim_abi.cu, and picks the narrowest kernel that covers it):
enum Feat {
  F_NONE = 0,
  F_DRAG = 1,     // rotor drag / rolling moment block, motor linearity != 1, velocity and omega^2 damping (quadrotor.py:302-349, :403, :431)
  F_NOISEX = 2,   // uniform position / velocity noise, quaternion noise, gyro-bias random walk (sensor_noise.py:116-170)
  F_MISC = 4,     // dim_mode 1D / 2D, per-env and moving goals, swarms outside the common shape, quaternion / t2w / t2t
                  // observations, pitch-roll-restricted spawn
  F_INFO = 8,     // diagnostics buffer (qs_bind_info): 48 words per env step, every raw and weighted reward term
  F_ALL = 15
};
template <int CTRL_, int OU_, int SENSE_, bool INJECT_, int FEAT_, int SIM_ = -1, bool PLAIN_ = false, bool SWARM_ = false,
          bool PEM_ = false>
struct Traits {
  static_assert(FEAT_ >= 0 && FEAT_ <= F_ALL, "FEAT is a Feat mask");
  static constexpr bool PEM = PEM_;     // per-env model constants (dynamics randomisation), read from the model buffer
  static constexpr int FEAT = FEAT_;
  static constexpr bool DRAG = (FEAT_ & F_DRAG) != 0, NOISEX = (FEAT_ & F_NOISEX) != 0, MISC = (FEAT_ & F_MISC) != 0;
  static constexpr bool INFO = (FEAT_ & F_INFO) != 0;
  static constexpr bool SWARM = SWARM_ || MISC;    // n_agents > 1 handled: contact stage, formation goals, log-distance reward
  static constexpr int CTRL = CTRL_;
  static constexpr int SIM = SIM_;      // compile-time sim_steps (substep loop unrolled, OU draws hoisted), -1 = run time
  static constexpr bool PLAIN = PLAIN_; // one env step per launch, 18-word observation, no previous-action row
  static constexpr int OU = OU_;
  static constexpr int SENSE = SENSE_;
  static constexpr bool INJECT = INJECT_;
  // MUFU-based sqrt / rsqrt / small-angle sincos forms of the fp32 build (~2 ulp; the fp64 build never takes them).  The
  // per-step parity tests run the injected-noise twins of the same instantiations.
#if defined(QS_EXP_NO_FASTM)
  static constexpr bool FASTM = false;
#elif defined(QS_EXP_GEN_NO_FASTM)
  static constexpr bool FASTM = FEAT_ == F_NONE;
#else
  static constexpr bool FASTM = true;
#endif
  static constexpr bool RARE = FEAT_ != F_NONE;
};

// replay of `since_last_svd += dt; if since_last_svd > limit: ...; since_last_svd = 0` in double (quadrotor.py:381-386)
__device__ __forceinline__ bool svd_tick(double& t, double dt, double period) {
  t += dt;
  const bool fire = t > period;
  if (fire) t = 0.0;
  return fire;
}

// last CTA of the launch advances the device clock: every CTA read it at its start, and a CTA only takes a ticket
// after its last use of the value, so nobody can observe the update within this launch
template <typename R>
__device__ __forceinline__ void finish_clock(const Params<R>& p, double svd_t_end, unsigned long long steps) {
  if (threadIdx.x == 0 && p.clock_advance) {
    const unsigned int t = atomicAdd(p.clock_ticket, 1u);
    if (t == gridDim.x - 1) {
      *p.clock_ticket = 0u;
      p.clock->svd_timer = svd_t_end;
      p.clock->step += steps;
    }
  }
}

// same ticket scheme for launches that must not advance the clock (chunks of the host pipeline) but still have to
// leave the work counter of the tile scheduler at zero.  Called by ONE thread of the CTA, after every thread of the
// CTA has read the clock.
template <typename R>
__device__ __forceinline__ void finish_launch_one(const Params<R>& p, double svd_t_end, unsigned long long steps) {
  const unsigned int t = atomicAdd(p.launch_ticket, 1u);
  if (t == gridDim.x - 1) {
    *p.launch_ticket = 0u;
    *p.tile_counter = 0u;
    if (p.clock_advance) { p.clock->svd_timer = svd_t_end; p.clock->step += steps; }
  }
}
template <typename R>
__device__ __forceinline__ void finish_launch(const Params<R>& p, double svd_t_end, unsigned long long steps) {
  if (threadIdx.x == 0) finish_launch_one(p, svd_t_end, steps);
}

// Programmatic dependent launch (PDL): the step kernel lets the NEXT launch on the stream become resident and run
// its prologue while this one is still computing, and itself waits for the PREVIOUS launch to complete (with its
// memory flushed) before it touches global memory.  Both are no-ops for a launch made without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------
// scalar maths
// ---------------------------------------------------------------------------------------------------------
template <typename R> struct Mth;
template <> struct Mth<float> {
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
  static __device__ __forceinline__ float acos(float x) { return acosf(x); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
  static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ bool finite(float x) { return isfinite(x); }
  // raw MUFU forms: no denormal pre/post-scaling (callers guarantee normal, positive arguments)
  static __device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static __device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  // Box-Muller pair from two 32-bit words (Philox mode only; distribution parity, not bit parity)
  static __device__ __forceinline__ void normal2(uint32_t a, uint32_t b, float* z0, float* z1) {
    float u1 = (float)((a >> 8) + 1u) * 5.9604644775390625e-08f;   // (0, 1], never denormal
    float t = lg2_ftz(u1) * -1.3862943611198906f;                  // -2 ln(u1) >= 0
    float r = t * rsqrt_ftz(fmaxf(t, 1e-30f));                     // sqrt(t), exact 0 at t == 0
    float ang = (float)(b >> 8) * 3.7450704e-07f;                  // 2 pi * 2^-24 * [0, 2^24)
    float s, c;
    __sincosf(ang, &s, &c);
    *z0 = r * c;
    *z1 = r * s;
  }
  static __device__ __forceinline__ float unif(uint32_t a) { return (float)(int32_t)a * 4.656612873077393e-10f; }  // [-1, 1]
};
template <> struct Mth<double> {
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
  static __device__ __forceinline__ double acos(double x) { return ::acos(x); }
  static __device__ __forceinline__ double log(double x) { return ::log(x); }
  static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
  static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ bool finite(double x) { return isfinite(x); }
  static __device__ __forceinline__ void normal2(uint32_t a, uint32_t b, double* z0, double* z1) {
    double u1 = ((double)a + 1.0) * 2.3283064365386963e-10;        // (0, 1]
    double u2 = (double)b * 2.3283064365386963e-10;                // [0, 1)
    double r = ::sqrt(-2.0 * ::log(u1));
    double s, c;
    ::sincospi(2.0 * u2, &s, &c);
    *z0 = r * c;
    *z1 = r * s;
  }
  static __device__ __forceinline__ double unif(uint32_t a) { return (double)(int32_t)a * 4.656612873077393e-10; }
};

// np.clip semantics: a NaN argument stays NaN (fminf/fmaxf would return the bound and hide a diverged env from the
// non-finite reward check, quadrotor.py:633-636).  fp32: FMNMX.NAN, one instruction per bound like the plain form.
__device__ __forceinline__ float clampr(float x, float lo, float hi) {
  float y;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(y), "f"(hi));
  return y;
}
__device__ __forceinline__ double clampr(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Fast fp32 forms used by the throughput variants (Traits::RARE == false): MUFU.RSQ / MUFU.RCP based, ~2 ulp, no
// denormal / special-case branches.  The per-step parity tests run these same variants (injected-noise twins).
template <typename R, bool FAST> __device__ __forceinline__ R q_rsqrt(R x) {
  if (FAST && sizeof(R) == 4) return (R)Mth<float>::rsqrt_ftz((float)x);   // callers pass normal positive values
  return R(1) / Mth<R>::sqrt(x);
}
template <typename R, bool FAST> __device__ __forceinline__ R q_sqrt(R x) {       // x >= 0
  if (FAST && sizeof(R) == 4) { float xf = (float)x; return (R)(xf * Mth<float>::rsqrt_ftz(fmaxf(xf, 1e-30f))); }
  return Mth<R>::sqrt(x);
}
template <typename R, bool FAST> __device__ __forceinline__ R q_rcp(R x) {
  if (FAST && sizeof(R) == 4) return (R)__frcp_rn((float)x);
  return R(1) / x;
}
// sin and cos of a small angle (the half rotation angle of one substep, <= omega_max*sqrt(3)*dt/2 ~ 0.17 rad at the
// default dt): Taylor polynomials to x^9 / x^8, relative truncation error < 3e-8 for |x| <= 0.8, far below
// the fp32 rounding of the matrix update; larger angles take the library routine.
template <typename R, bool FAST> __device__ __forceinline__ void q_sincos_small(R x, R* s, R* c) {
  if (FAST && sizeof(R) == 4) {
    float xf = (float)x;
    if (fabsf(xf) <= 0.8f) {
      float x2 = xf * xf;
      float sp = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -1.6666667e-1f), 1.0f);
      float cp = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
      *s = (R)(xf * sp);
      *c = (R)cp;
      return;
    }
  }
  Mth<R>::sincos(x, s, c);
}

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no RNG state in HBM
// ---------------------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Philox slot map (8 bits): obs groups 0..9 (step) / 16..25 (reset obs), t2w/t2t observation noise 11 (step) / 41 (reset),
// reset draws 26..38, OU substeps 64+s
enum Slot { SLOT_OBS = 0, SLOT_T2W = 11, SLOT_EXCITE = 12, SLOT_RESET_OBS = 16, SLOT_RESET_POS = 26, SLOT_RESET_VEL = 27, SLOT_RESET_OM = 28,
            SLOT_RESET_UP = 29, SLOT_RESET_FWD = 30, SLOT_RESET_GOAL = 40, SLOT_RESET_T2W = 41,
            SLOT_OU = 64 };

template <typename R>
struct Rng {
  uint32_t k0, k1, e0, e1, s0, s1;
  __device__ __forceinline__ void init(const Params<R>& p, long long i, unsigned long long step) {
    unsigned long long gid = (unsigned long long)(p.env_id_offset + i);
    k0 = p.seed_lo; k1 = p.seed_hi;
    e0 = (uint32_t)gid; e1 = (uint32_t)(gid >> 32);
    s0 = (uint32_t)step; s1 = (uint32_t)(step >> 32) & 0x00FFFFFFu;
  }
  __device__ __forceinline__ U4 raw(int slot) const {
    U4 c; c.x = e0; c.y = e1; c.z = s0; c.w = s1 | ((uint32_t)slot << 24);
    return philox4x32_10(c, k0, k1);
  }
  __device__ __forceinline__ void normal4(int slot, R* z) const {
    U4 u = raw(slot);
    Mth<R>::normal2(u.x, u.y, &z[0], &z[1]);
    Mth<R>::normal2(u.z, u.w, &z[2], &z[3]);
  }
  __device__ __forceinline__ void unif4(int slot, R* z) const {
    U4 u = raw(slot);
    z[0] = Mth<R>::unif(u.x); z[1] = Mth<R>::unif(u.y); z[2] = Mth<R>::unif(u.z); z[3] = Mth<R>::unif(u.w);
  }
};

// ---------------------------------------------------------------------------------------------------------
// Model constants as one thread sees them.  One model per handle: straight out of the constant bank (Params).
// Per-env models (dynamics randomisation, quadrotor_randomization.py): read-only loads from the model buffer, row
// layout in quadsim.h; nothing is cached in registers (the fp64 build has none to spare), L1 serves the re-reads.
// ---------------------------------------------------------------------------------------------------------
template <typename R> struct alignas(4 * sizeof(R)) V4;
template <typename R> __device__ __forceinline__ V4<R> ldg4(const R* ptr);
template <typename R, bool PEM> struct Mdl;
template <typename R> struct Mdl<R, false> {
  const Params<R>& p;
  __device__ __forceinline__ Mdl(const Params<R>& p_, long long) : p(p_) {}
  __device__ __forceinline__ R thrust_max(int k) const { return p.thrust_max[k]; }
  __device__ __forceinline__ R tz(int k) const { return p.tz[k]; }
  __device__ __forceinline__ R pcx(int k) const { return p.pcx[k]; }
  __device__ __forceinline__ R pcy(int k) const { return p.pcy[k]; }
  __device__ __forceinline__ R ppx(int k) const { return p.ppx[k]; }
  __device__ __forceinline__ R ppy(int k) const { return p.ppy[k]; }
  __device__ __forceinline__ R ppz(int k) const { return p.ppz[k]; }
  __device__ __forceinline__ R inertia(int k) const { return p.inertia[k]; }
  __device__ __forceinline__ R inv_inertia(int k) const { return p.inv_inertia[k]; }
  __device__ __forceinline__ R mass() const { return p.mass; }
  __device__ __forceinline__ R inv_mass() const { return p.inv_mass; }
  __device__ __forceinline__ R tau_up() const { return p.tau_up; }
  __device__ __forceinline__ R tau_down() const { return p.tau_down; }
  __device__ __forceinline__ R ou_sigma() const { return p.ou_sigma; }
  __device__ __forceinline__ R arm() const { return p.arm; }
  __device__ __forceinline__ R lin() const { return p.lin; }
  __device__ __forceinline__ R c_drag() const { return p.c_drag; }
  __device__ __forceinline__ R c_roll() const { return p.c_roll; }
  __device__ __forceinline__ R t2w() const { return p.t2w; }
  __device__ __forceinline__ R t2t() const { return p.t2t; }
  __device__ __forceinline__ R vel_damp() const { return p.vel_damp; }
  __device__ __forceinline__ R damp_omega_q() const { return p.damp_omega_q; }
  __device__ __forceinline__ R jinv(int r, int c) const { return p.jinv[r][c]; }
};
template <typename R> struct Mdl<R, true> {
  const R* base;          // element 0 of this env's vector in row 0
  long long stride;       // elements between rows
  __device__ __forceinline__ Mdl(const Params<R>& p, long long i) : base(p.models + i * 4), stride(p.n * 4) { QS_CHECK(p, i >= 0 && i < p.n, BC_MODEL); }
  // one 16-byte (fp64: two 16-byte) read-only vector load per row; the component index is a compile-time constant at every
  // call site, so the select folds away and loads of the same row are merged by the compiler (four scalar loads per row before)
  __device__ __forceinline__ R at(int row, int c) const {
    const V4<R> v = ldg4(base + row * stride);
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
  }
  __device__ __forceinline__ R thrust_max(int k) const { return at(0, k); }
  __device__ __forceinline__ R tz(int k) const { return at(1, k); }
  __device__ __forceinline__ R pcx(int k) const { return at(2, k); }
  __device__ __forceinline__ R pcy(int k) const { return at(3, k); }
  __device__ __forceinline__ R ppx(int k) const { return -at(3, k); }
  __device__ __forceinline__ R ppy(int k) const { return at(2, k); }
  __device__ __forceinline__ R ppz(int k) const { return at(4, k); }
  __device__ __forceinline__ R inertia(int k) const { return at(5, k); }
  __device__ __forceinline__ R inv_inertia(int k) const { return at(6, k); }
  __device__ __forceinline__ R mass() const { return at(5, 3); }
  __device__ __forceinline__ R inv_mass() const { return at(6, 3); }
  __device__ __forceinline__ R tau_up() const { return at(7, 0); }
  __device__ __forceinline__ R tau_down() const { return at(7, 1); }
  __device__ __forceinline__ R ou_sigma() const { return at(7, 2); }
  __device__ __forceinline__ R arm() const { return at(7, 3); }
  __device__ __forceinline__ R lin() const { return at(8, 0); }
  __device__ __forceinline__ R c_drag() const { return at(8, 1); }
  __device__ __forceinline__ R c_roll() const { return at(8, 2); }
  __device__ __forceinline__ R t2w() const { return at(8, 3); }
  __device__ __forceinline__ R vel_damp() const { return at(9, 0); }
  __device__ __forceinline__ R damp_omega_q() const { return at(9, 1); }
  __device__ __forceinline__ R t2t() const { return at(9, 2); }
  __device__ __forceinline__ R jinv(int r, int c) const { return at(10 + r, c); }
};

// ---------------------------------------------------------------------------------------------------------
// per-env state in registers
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct St {
  R px, py, pz, vx, vy, vz, wx, wy, wz;
  R r[9];                 // rotation, row-major
  R rd[4], cd[4], ou[4];  // thrust_rot_damp, thrust_cmds_damp, OU state
  R tick, ret;
  R pa[4];                // previous action (actions[0] of the last step)
  R gb[3];                // gyro bias
  R gx, gy, gz;           // goal
  // not persisted
  R ax, ay, az;           // world acceleration of the last substep (dynamics.acc)
  R tqx, tqy, tqz;        // dynamics.torque
  R odx, ody, odz;        // dynamics.omega_dot
  bool w_is_f32;          // set_state stored omega as a float32 array (quadrotor.py:223), see substep()
  bool fresh;             // accelerometer still holds its set_state value (0,0,GRAV) (quadrotor.py:221)
};

template <typename R> struct alignas(4 * sizeof(R)) V4 { R x, y, z, w; };
template <> __device__ __forceinline__ V4<float> ldg4<float>(const float* ptr) {
  const float4 q = __ldg(reinterpret_cast<const float4*>(ptr));
  V4<float> v; v.x = q.x; v.y = q.y; v.z = q.z; v.w = q.w;
  return v;
}
template <> __device__ __forceinline__ V4<double> ldg4<double>(const double* ptr) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(ptr)), b = __ldg(reinterpret_cast<const double2*>(ptr) + 1);
  V4<double> v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
  return v;
}

template <typename R> __device__ __forceinline__ V4<R> ld4(const R* base, long long n, int row, long long i) {
  return *reinterpret_cast<const V4<R>*>(base + ((long long)row * n + i) * 4);
}
template <typename R> __device__ __forceinline__ void st4(R* base, long long n, int row, long long i, R x, R y, R z, R w) {
  V4<R> v; v.x = x; v.y = y; v.z = z; v.w = w;
  *reinterpret_cast<V4<R>*>(base + ((long long)row * n + i) * 4) = v;
}

template <typename R, typename Q>
__device__ __forceinline__ void unpack_core(St<R>& s, const Q& q0, const Q& q1, const Q& q2, const Q& q3, const Q& q4,
                                            const Q& q5, const Q& q6) {
  s.px = q0.x; s.py = q0.y; s.pz = q0.z; s.vx = q0.w;
  s.vy = q1.x; s.vz = q1.y; s.wx = q1.z; s.wy = q1.w;
  s.wz = q2.x; s.r[0] = q2.y; s.r[1] = q2.z; s.r[2] = q2.w;
  s.r[3] = q3.x; s.r[4] = q3.y; s.r[5] = q3.z; s.r[6] = q3.w;
  s.r[7] = q4.x; s.r[8] = q4.y; s.tick = q4.z; s.ret = q4.w;
  s.rd[0] = q5.x; s.rd[1] = q5.y; s.rd[2] = q5.z; s.rd[3] = q5.w;
  s.cd[0] = q6.x; s.cd[1] = q6.y; s.cd[2] = q6.z; s.cd[3] = q6.w;
}

template <typename R>
__device__ __forceinline__ void default_core(St<R>& s) {   // lanes beyond the last env: harmless identity state
  s.px = s.py = s.pz = s.vx = s.vy = s.vz = s.wx = s.wy = s.wz = R(0);
#pragma unroll
  for (int k = 0; k < 9; ++k) s.r[k] = (k % 4 == 0) ? R(1) : R(0);
#pragma unroll
  for (int k = 0; k < 4; ++k) { s.rd[k] = R(0); s.cd[k] = R(0); }
  s.tick = R(0); s.ret = R(0);
}

template <typename R, bool SWARM = true>
__device__ __forceinline__ void init_optional(St<R>& s, const Params<R>& p) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { s.ou[k] = R(0); s.pa[k] = R(0); }
  s.gb[0] = s.gb[1] = s.gb[2] = R(0);
  s.gx = p.goal[0]; s.gy = p.goal[1]; s.gz = p.goal[2];
  if (SWARM) {
    if (p.formation_on) {                     // launches start on a multiple of BLOCK, and n_agents divides 32
      const int a = (int)threadIdx.x & (p.n_agents - 1);
      s.gx = p.formation[a][0]; s.gy = p.formation[a][1]; s.gz = p.formation[a][2];
    }
  }
}

template <typename R>
__device__ __forceinline__ void finish_load(St<R>& s) {
  s.ax = s.ay = s.az = R(0);
  s.tqx = s.tqy = s.tqz = s.odx = s.ody = s.odz = R(0);
  // straight after set_state()/reset() (tick 0, motor filter zeroed, :438-440) omega is still the float32 array of :223
  s.w_is_f32 = (s.tick == R(0)) && (s.rd[0] == R(0)) && (s.rd[1] == R(0)) && (s.rd[2] == R(0)) && (s.rd[3] == R(0)) &&
               (s.cd[0] == R(0)) && (s.cd[1] == R(0)) && (s.cd[2] == R(0)) && (s.cd[3] == R(0));
  s.fresh = false;
}

template <typename R, class T>
__device__ __forceinline__ void load_state(St<R>& s, const Params<R>& p, long long i, bool active) {
  QS_CHECK(p, !active || (i >= p.begin && i < p.end && p.end <= p.n), BC_ENV);
  QS_CHECK(p, p.row_ou < p.tma_rows && p.row_act < p.tma_rows && p.row_gyro < p.tma_rows && p.row_goal < p.tma_rows && p.tma_rows >= 7, BC_ROW);
  if (active) {
    const char* ptr = reinterpret_cast<const char*>(p.state + i * 4);   // row 0; rows are row_bytes apart
    const long long rb = p.row_bytes;
    V4<R> q0 = *reinterpret_cast<const V4<R>*>(ptr); ptr += rb;
    V4<R> q1 = *reinterpret_cast<const V4<R>*>(ptr); ptr += rb;
    V4<R> q2 = *reinterpret_cast<const V4<R>*>(ptr); ptr += rb;
    V4<R> q3 = *reinterpret_cast<const V4<R>*>(ptr); ptr += rb;
    V4<R> q4 = *reinterpret_cast<const V4<R>*>(ptr); ptr += rb;
    V4<R> q5 = *reinterpret_cast<const V4<R>*>(ptr); ptr += rb;
    V4<R> q6 = *reinterpret_cast<const V4<R>*>(ptr);
    unpack_core(s, q0, q1, q2, q3, q4, q5, q6);
  } else {
    default_core(s);
  }
  init_optional<R, T::SWARM>(s, p);
  const bool ou_on = T::OU >= 0 ? (T::OU != 0) : (p.ou_on != 0);
  if (active) {
    if (ou_on) { V4<R> q = ld4(p.state, p.n, p.row_ou, i); s.ou[0] = q.x; s.ou[1] = q.y; s.ou[2] = q.z; s.ou[3] = q.w; }
    if (!T::PLAIN && p.row_act >= 0) { V4<R> q = ld4(p.state, p.n, p.row_act, i); s.pa[0] = q.x; s.pa[1] = q.y; s.pa[2] = q.z; s.pa[3] = q.w; }
    if (T::NOISEX) { if (p.row_gyro >= 0) { V4<R> q = ld4(p.state, p.n, p.row_gyro, i); s.gb[0] = q.x; s.gb[1] = q.y; s.gb[2] = q.z; } }
    if (T::MISC) { if (p.row_goal >= 0) { V4<R> q = ld4(p.state, p.n, p.row_goal, i); s.gx = q.x; s.gy = q.y; s.gz = q.z; } }
  }
  finish_load(s);
}

// CS: streaming (evict-first) store - one-wave launches only (see env_step_tma_kernel)
template <typename R, bool CS = false> __device__ __forceinline__ void st4p(char*& ptr, long long stride, R x, R y, R z, R w) {
  V4<R> v; v.x = x; v.y = y; v.z = z; v.w = w;
  if (CS && sizeof(R) == 4) __stcs(reinterpret_cast<float4*>(ptr), make_float4((float)x, (float)y, (float)z, (float)w));
  else *reinterpret_cast<V4<R>*>(ptr) = v;
  ptr += stride;
}

template <typename R, class T, bool CS = false>
__device__ __forceinline__ void store_state(const St<R>& s, const Params<R>& p, long long i) {
#ifdef QS_EXP_OLD_ADDR
  st4(p.state, p.n, 0, i, s.px, s.py, s.pz, s.vx);
  st4(p.state, p.n, 1, i, s.vy, s.vz, s.wx, s.wy);
  st4(p.state, p.n, 2, i, s.wz, s.r[0], s.r[1], s.r[2]);
  st4(p.state, p.n, 3, i, s.r[3], s.r[4], s.r[5], s.r[6]);
  st4(p.state, p.n, 4, i, s.r[7], s.r[8], s.tick, s.ret);
  st4(p.state, p.n, 5, i, s.rd[0], s.rd[1], s.rd[2], s.rd[3]);
  st4(p.state, p.n, 6, i, s.cd[0], s.cd[1], s.cd[2], s.cd[3]);
#else
  QS_CHECK(p, i >= p.begin && i < p.end && p.end <= p.n && p.row_bytes == p.n * 4 * (long long)sizeof(R), BC_ENV);
  char* ptr = reinterpret_cast<char*>(p.state + i * 4);                 // row 0; rows are row_bytes apart
  const long long rb = p.row_bytes;
  st4p<R, CS>(ptr, rb, s.px, s.py, s.pz, s.vx);
  st4p<R, CS>(ptr, rb, s.vy, s.vz, s.wx, s.wy);
  st4p<R, CS>(ptr, rb, s.wz, s.r[0], s.r[1], s.r[2]);
  st4p<R, CS>(ptr, rb, s.r[3], s.r[4], s.r[5], s.r[6]);
  st4p<R, CS>(ptr, rb, s.r[7], s.r[8], s.tick, s.ret);
  st4p<R, CS>(ptr, rb, s.rd[0], s.rd[1], s.rd[2], s.rd[3]);
  st4p<R, CS>(ptr, rb, s.cd[0], s.cd[1], s.cd[2], s.cd[3]);
#endif
  const bool ou_on = T::OU >= 0 ? (T::OU != 0) : (p.ou_on != 0);
  if (ou_on) st4(p.state, p.n, p.row_ou, i, s.ou[0], s.ou[1], s.ou[2], s.ou[3]);
  if (!T::PLAIN && p.row_act >= 0) st4(p.state, p.n, p.row_act, i, s.pa[0], s.pa[1], s.pa[2], s.pa[3]);
  if (T::NOISEX) { if (p.row_gyro >= 0) st4(p.state, p.n, p.row_gyro, i, s.gb[0], s.gb[1], s.gb[2], R(0)); }
  // the goal row is read-only for the step path unless the goal moves
  if (T::MISC) { if (p.row_goal >= 0 && (p.resample_goal || p.excite)) st4(p.state, p.n, p.row_goal, i, s.gx, s.gy, s.gz, R(0)); }
}

// ---------------------------------------------------------------------------------------------------------
// polar re-orthonormalisation: replaces u,s,v = svd(R); R = u @ v (quadrotor.py:384-385).
// R is within rounding of a rotation, so Newton's iteration X <- (X + X^-T)/2 reaches the same polar factor
// to machine precision in one or two steps; four are run so a caller-supplied sloppy R also converges.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void polar_orthonormalize(R* r) {
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    R c00 = r[4] * r[8] - r[5] * r[7], c01 = r[5] * r[6] - r[3] * r[8], c02 = r[3] * r[7] - r[4] * r[6];
    R c10 = r[2] * r[7] - r[1] * r[8], c11 = r[0] * r[8] - r[2] * r[6], c12 = r[1] * r[6] - r[0] * r[7];
    R c20 = r[1] * r[5] - r[2] * r[4], c21 = r[2] * r[3] - r[0] * r[5], c22 = r[0] * r[4] - r[1] * r[3];
    R det = r[0] * c00 + r[1] * c01 + r[2] * c02;
    R h = R(0.5) / det;
    r[0] = R(0.5) * r[0] + h * c00; r[1] = R(0.5) * r[1] + h * c01; r[2] = R(0.5) * r[2] + h * c02;
    r[3] = R(0.5) * r[3] + h * c10; r[4] = R(0.5) * r[4] + h * c11; r[5] = R(0.5) * r[5] + h * c12;
    r[6] = R(0.5) * r[6] + h * c20; r[7] = R(0.5) * r[7] + h * c21; r[8] = R(0.5) * r[8] + h * c22;
  }
}

// ---------------------------------------------------------------------------------------------------------
// QuadrotorDynamics.step1 (quadrotor.py:273-436).  z: 4 unit normals for the OU thrust noise.
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__device__ __forceinline__ void substep(St<R>& s, const R* cmd, const Params<R>& p, const Mdl<R, T::PEM>& m, const R* z,
                                        bool fire_svd) {
  const bool ou_on = T::OU >= 0 ? (T::OU != 0) : (p.ou_on != 0);
  const R dt = p.dt;
  R tsum = R(0), tqx = R(0), tqy = R(0), tqz = R(0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    R c = clampr(cmd[k], R(0), R(1));                                   // :279
    R tau = (c < s.cd[k]) ? m.tau_down() : m.tau_up();                      // :284-288 (tau pre-clipped to 1 on the host)
    R rt = q_sqrt<R, T::FASTM>(c);                                      // :294
    s.rd[k] = tau * (rt - s.rd[k]) + s.rd[k];                           // :295
    R cd = s.rd[k] * s.rd[k];                                           // :296
    if (ou_on) {
      s.ou[k] = s.ou[k] + (p.ou_theta * (-s.ou[k]) + m.ou_sigma() * z[k]);   // quad_utils.py:197-201 (mu = 0)
      cd = cd + c * s.ou[k];                                            // :299
    }
    cd = clampr(cd, R(0), R(1));                                        // :300
    s.cd[k] = cd;
    R t;
    if (T::DRAG) t = m.thrust_max(k) * ((R(1) - m.lin()) * cd * cd + m.lin() * cd);   // :302, :140
    else t = m.thrust_max(k) * cd;                                      // linearity == 1
    tqx += m.pcx(k) * t;                                                // :304, :310
    tqy += m.pcy(k) * t;
    tqz += m.tz(k) * cd;                                                // :307
    tsum += t;                                                          // :362
  }
  R fdx = R(0), fdy = R(0), fdz = R(0);                                 // rotor drag force (body frame)
  if (T::DRAG) {
    if (m.c_drag() != R(0) || m.c_roll() != R(0)) {                         // :318-349
      R vbx = s.r[0] * s.vx + s.r[3] * s.vy + s.r[6] * s.vz;            // rot.T @ vel
      R vby = s.r[1] * s.vx + s.r[4] * s.vy + s.r[7] * s.vz;
      R vbz = s.r[2] * s.vx + s.r[5] * s.vy + s.r[8] * s.vz;
      R dtx = R(0), dty = R(0), dtz = R(0), rtx = R(0), rty = R(0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        R vrx = vbx + (s.wy * m.ppz(k) - s.wz * m.ppy(k));              // :323 omega x prop_pos
        R vry = vby + (s.wz * m.ppx(k) - s.wx * m.ppz(k));              // z component projected out (:325)
        R sq = q_sqrt<R, T::FASTM>(s.cd[k]);
        R fx = -m.c_drag() * sq * vrx, fy = -m.c_drag() * sq * vry;         // :328
        fdx += fx; fdy += fy;
        dtx += fy * m.ppz(k);                                           // :331 f x prop_pos (f.z = 0)
        dty += -fx * m.ppz(k);
        dtz += fx * m.ppy(k) - fy * m.ppx(k);
        rtx += -m.c_roll() * p.ccw[k] * sq * vrx;                         // :334
        rty += -m.c_roll() * p.ccw[k] * sq * vry;
      }
      R vtx = dtx + rtx, vty = dty + rty, vtz = dtz;                    // :336
      R iwx = s.wx * m.inertia(0), iwy = s.wy * m.inertia(1), iwz = s.wz * m.inertia(2);
      if (T::FASTM && sizeof(R) == 4) {
        // fp32: norms by MUFU.RSQ (x * rsqrt(x)), the two rescalings by the reciprocal norm already at hand - no IEEE
        // square roots or divisions (they were 280 of this kernel's 1,880 instructions per env step)
        const R hid = R(0.5) / dt;
        R vel_norm = q_sqrt<R, true>(vbx * vbx + vby * vby + vbz * vbz);  // :339
        R rdf2 = fdx * fdx + fdy * fdy;                                   // :340
        R irdf = q_rsqrt<R, true>(Mth<R>::max(rdf2, R(1e-30)));
        R rdf = rdf2 * irdf;
        R rdf_clip = clampr(rdf, R(0), vel_norm * m.mass() * hid);        // :341
        if (rdf > R(1e-6)) { R sc = rdf_clip * irdf; fdx = fdx * sc; fdy = fdy * sc; }   // :342-343
        R rvt2 = vtx * vtx + vty * vty + vtz * vtz;                       // :346
        R irvt = q_rsqrt<R, true>(Mth<R>::max(rvt2, R(1e-30)));
        R rvt = rvt2 * irvt;
        R rvt_clip = clampr(rvt, R(0), q_sqrt<R, true>(iwx * iwx + iwy * iwy + iwz * iwz) * hid);   // :347
        if (rvt > R(1e-6)) { R sc = rvt_clip * irvt; vtx = vtx * sc; vty = vty * sc; vtz = vtz * sc; }
      } else {
      R vel_norm = Mth<R>::sqrt(vbx * vbx + vby * vby + vbz * vbz);     // :339
      R rdf = Mth<R>::sqrt(fdx * fdx + fdy * fdy);                      // :340
      R rdf_clip = clampr(rdf, R(0), vel_norm * m.mass() / (R(2) * dt));  // :341
      if (rdf > R(1e-6)) { R sc = rdf_clip; fdx = (fdx / rdf) * sc; fdy = (fdy / rdf) * sc; }   // :342-343
      R rvt = Mth<R>::sqrt(vtx * vtx + vty * vty + vtz * vtz);          // :346
      R rvt_clip = clampr(rvt, R(0), Mth<R>::sqrt(iwx * iwx + iwy * iwy + iwz * iwz) / (R(2) * dt));   // :347
      if (rvt > R(1e-6)) { vtx = (vtx / rvt) * rvt_clip; vty = (vty / rvt) * rvt_clip; vtz = (vtz / rvt) * rvt_clip; }
      }
      tqx += vtx; tqy += vty; tqz += vtz;                               // :361
    }
  }
  s.tqx = tqx; s.tqy = tqy; s.tqz = tqz;
  // ---- attitude: Rodrigues with the world-frame rate, left-multiplied (:370-378) ----
  {
    R ox = s.r[0] * s.wx + s.r[1] * s.wy + s.r[2] * s.wz;
    R oy = s.r[3] * s.wx + s.r[4] * s.wy + s.r[5] * s.wz;
    R oz = s.r[6] * s.wx + s.r[7] * s.wy + s.r[8] * s.wz;
    R on2 = ox * ox + oy * oy + oz * oz;
    // reference: skipped when |omega| == 0 (:373).  The fast form needs no branch: with omega == 0 the axis is the
    // zero vector and the half angle is 0, so dR is exactly the identity and R comes out bit-identical.
    if ((T::FASTM && sizeof(R) == 4) || on2 != R(0)) {
      R inv, on;
      if (T::FASTM && sizeof(R) == 4) { inv = q_rsqrt<R, true>(Mth<R>::max(on2, R(1e-30))); on = on2 * inv; }
      else { on = Mth<R>::sqrt(on2); inv = R(1) / on; }
      R kx = ox * inv, ky = oy * inv, kz = oz * inv;
      // sin(a) and 1 - cos(a) from the half angle: no cancellation for the small per-substep angles
      R sh, ch;
      q_sincos_small<R, T::FASTM>(R(0.5) * (on * dt), &sh, &ch);
      R sa = R(2) * sh * ch, cm = R(2) * sh * sh;
      R kxx = kx * kx, kyy = ky * ky, kzz = kz * kz, kxy = kx * ky, kxz = kx * kz, kyz = ky * kz;
      R d00 = R(1) - cm * (kyy + kzz), d01 = cm * kxy - sa * kz, d02 = cm * kxz + sa * ky;
      R d10 = cm * kxy + sa * kz, d11 = R(1) - cm * (kxx + kzz), d12 = cm * kyz - sa * kx;
      R d20 = cm * kxz - sa * ky, d21 = cm * kyz + sa * kx, d22 = R(1) - cm * (kxx + kyy);
      R n0 = d00 * s.r[0] + d01 * s.r[3] + d02 * s.r[6], n1 = d00 * s.r[1] + d01 * s.r[4] + d02 * s.r[7];
      R n2 = d00 * s.r[2] + d01 * s.r[5] + d02 * s.r[8];
      R n3 = d10 * s.r[0] + d11 * s.r[3] + d12 * s.r[6], n4 = d10 * s.r[1] + d11 * s.r[4] + d12 * s.r[7];
      R n5 = d10 * s.r[2] + d11 * s.r[5] + d12 * s.r[8];
      R n6 = d20 * s.r[0] + d21 * s.r[3] + d22 * s.r[6], n7 = d20 * s.r[1] + d21 * s.r[4] + d22 * s.r[7];
      R n8 = d20 * s.r[2] + d21 * s.r[5] + d22 * s.r[8];
      s.r[0] = n0; s.r[1] = n1; s.r[2] = n2; s.r[3] = n3; s.r[4] = n4; s.r[5] = n5; s.r[6] = n6; s.r[7] = n7; s.r[8] = n8;
    }
  }
  if (fire_svd) polar_orthonormalize(s.r);                              // :381-386 (uniform across the grid)
  // ---- body rates (:398-405): old omega in the gyroscopic term ----
  {
    R iwx = m.inertia(0) * s.wx, iwy = m.inertia(1) * s.wy, iwz = m.inertia(2) * s.wz;
    R gx = (-s.wy) * iwz - (-s.wz) * iwy;                               // cross(-omega, I omega)
    R gy = (-s.wz) * iwx - (-s.wx) * iwz;
    R gz = (-s.wx) * iwy - (-s.wy) * iwx;
    s.odx = m.inv_inertia(0) * (gx + tqx);
    s.ody = m.inv_inertia(1) * (gy + tqy);
    s.odz = m.inv_inertia(2) * (gz + tqz);
    R fx = dt, fy = dt, fz = dt;
    if (T::DRAG) {
      R w2x = s.wx * s.wx, w2y = s.wy * s.wy, w2z = s.wz * s.wz;
      if (sizeof(R) == 8 && s.w_is_f32) {   // omega ** 2 on the float32 array set_state left behind (:223, :403)
        w2x = (R)((float)s.wx * (float)s.wx); w2y = (R)((float)s.wy * (float)s.wy); w2z = (R)((float)s.wz * (float)s.wz);
      }
      fx = (R(1) - clampr(m.damp_omega_q() * w2x, R(0), R(1))) * dt;      // :403-404
      fy = (R(1) - clampr(m.damp_omega_q() * w2y, R(0), R(1))) * dt;
      fz = (R(1) - clampr(m.damp_omega_q() * w2z, R(0), R(1))) * dt;
    }
    s.wx = clampr(s.wx + fx * s.odx, -p.omega_max, p.omega_max);        // :404-405
    s.wy = clampr(s.wy + fy * s.ody, -p.omega_max, p.omega_max);
    s.wz = clampr(s.wz + fz * s.odz, -p.omega_max, p.omega_max);
    s.w_is_f32 = false;
  }
  // ---- translation (:418-436): position with the old velocity, acceleration with the new R ----
  s.px = clampr(s.px + dt * s.vx, p.room_lo[0], p.room_hi[0]);
  s.py = clampr(s.py + dt * s.vy, p.room_lo[1], p.room_hi[1]);
  s.pz = clampr(s.pz + dt * s.vz, p.room_lo[2], p.room_hi[2]);
  R bx = fdx, by = fdy, bz = tsum + fdz;                                // thrust + rotor_drag_force (body)
  R ax, ay, az;
  if (T::DRAG) {
    ax = m.inv_mass() * (s.r[0] * bx + s.r[1] * by + s.r[2] * bz);
    ay = m.inv_mass() * (s.r[3] * bx + s.r[4] * by + s.r[5] * bz);
    az = m.inv_mass() * (s.r[6] * bx + s.r[7] * by + s.r[8] * bz) - R(9.81);
  } else {
    ax = m.inv_mass() * (s.r[2] * bz);
    ay = m.inv_mass() * (s.r[5] * bz);
    az = m.inv_mass() * (s.r[8] * bz) - R(9.81);                          // GRAV constant, not self.gravity (:426)
  }
  s.ax = ax; s.ay = ay; s.az = az;
  if (T::DRAG) {
    R vd = R(1) - m.vel_damp();
    s.vx = vd * s.vx + dt * ax; s.vy = vd * s.vy + dt * ay; s.vz = vd * s.vz + dt * az;   // :431
  } else {
    s.vx = s.vx + dt * ax; s.vy = s.vy + dt * ay; s.vz = s.vz + dt * az;
  }
  s.fresh = false;
}

// accelerometer = R^T (acc + (0,0,gravity)) (:436); (0,0,GRAV) straight after set_state (:221)
template <typename R>
__device__ __forceinline__ void accelerometer(const St<R>& s, const Params<R>& p, R* a) {
  if (s.fresh) { a[0] = R(0); a[1] = R(0); a[2] = R(9.81); return; }
  R x = s.ax, y = s.ay, z = s.az + p.gravity;
  a[0] = s.r[0] * x + s.r[3] * y + s.r[6] * z;
  a[1] = s.r[1] * x + s.r[4] * y + s.r[7] * z;
  a[2] = s.r[2] * x + s.r[5] * y + s.r[8] * z;
}

// ---------------------------------------------------------------------------------------------------------
// controllers: policy action -> 4 normalized motor commands (quadrotor_control.py)
// ---------------------------------------------------------------------------------------------------------
template <typename R, bool FAST = false> __device__ __forceinline__ void normalize3(R& x, R& y, R& z) {   // quad_utils.py:35-41
  if (FAST && sizeof(R) == 4) {
    R n2 = x * x + y * y + z * z;
    if (!(n2 < R(1e-10))) { R inv = q_rsqrt<R, true>(n2); x = x * inv; y = y * inv; z = z * inv; }
    return;
  }
  R n = Mth<R>::sqrt(x * x + y * y + z * z);
  if (!(n < R(0.00001))) { x = x / n; y = y / n; z = z / n; }
}

template <typename R, class M>
__device__ __forceinline__ void jinv_mix(const M& m, R d0, R d1, R d2, R d3, R* cmd) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    cmd[k] = clampr(m.jinv(k, 0) * d0 + m.jinv(k, 1) * d1 + m.jinv(k, 2) * d2 + m.jinv(k, 3) * d3, R(0), R(1));
}

template <typename R, bool FAST, class M>
__device__ __forceinline__ void attitude_controller(const St<R>& s, const Params<R>& p, const M& m, bool mellinger, const R* a,
                                                    R* cmd) {
  R adx, ady, adz, xcx, xcy, xcz, ewx, ewy, ewz, kp_a, kd_a, yaw_scale;
  if (mellinger) {                                                      // :315-362
    R tx = s.gx - s.px, ty = s.gy - s.py, tz = s.gz - s.pz;
    R n2 = tx * tx + ty * ty + tz * tz;
    if (!(n2 <= R(16))) {                                               // clamp_norm, quad_utils.py:63-67
      R f = FAST ? R(4) * q_rsqrt<R, FAST>(n2) : R(4) / Mth<R>::sqrt(n2);
      tx = f * tx; ty = f * ty; tz = f * tz;
    }
    R epx = -tx, epy = -ty, epz = -tz;                                  // :320
    adx = R(-4.5) * epx - R(3.5) * s.vx;                                // :323
    ady = R(-4.5) * epy - R(3.5) * s.vy;
    adz = R(-4.5) * epz - R(3.5) * s.vz + R(9.81);
    xcx = R(1); xcy = R(0); xcz = R(0);                                 // :333
    ewx = s.wx; ewy = s.wy; ewz = s.wz;
    kp_a = R(200); kd_a = R(50); yaw_scale = R(0.2);
  } else {                                                              // VelocityYawControl :250-281 (thrust_mag per :275)
    adx = R(-5) * (s.vx - a[0]); ady = R(-5) * (s.vy - a[1]); adz = R(-5) * (s.vz - a[2]) + R(9.81);
    xcx = s.r[0]; xcy = s.r[3]; xcz = s.r[6];
    ewx = s.wx - R(0); ewy = s.wy - R(0); ewz = s.wz - a[3];
    kp_a = R(100); kd_a = R(50); yaw_scale = R(1);
  }
  R zx = adx, zy = ady, zz = adz;
  normalize3<R, FAST>(zx, zy, zz);                                      // :338
  R yx = zy * xcz - zz * xcy, yy = zz * xcx - zx * xcz, yz = zx * xcy - zy * xcx;   // cross(zb, xc)
  normalize3<R, FAST>(yx, yy, yz);                                      // :339
  R xx = yy * zz - yz * zy, xy = yz * zx - yx * zz, xz = yx * zy - yy * zx;          // cross(yb, zb) :340
  // A = R_des^T R ; e_R = 0.5 * vee(A - A^T) (:344-346)
  R a21 = zx * s.r[1] + zy * s.r[4] + zz * s.r[7], a12 = yx * s.r[2] + yy * s.r[5] + yz * s.r[8];
  R a02 = xx * s.r[2] + xy * s.r[5] + xz * s.r[8], a20 = zx * s.r[0] + zy * s.r[3] + zz * s.r[6];
  R a10 = yx * s.r[0] + yy * s.r[3] + yz * s.r[6], a01 = xx * s.r[1] + xy * s.r[4] + xz * s.r[7];
  R erx = R(0.5) * (a21 - a12), ery = R(0.5) * (a02 - a20), erz = R(0.5) * (a10 - a01);
  erz *= yaw_scale;                                                     // :347
  R kwx = kd_a * ewx, kwy = kd_a * ewy, kwz = kd_a * ewz;
  if (sizeof(R) == 8 && mellinger && s.w_is_f32) {                      // float32 product on the f32 omega array (:350)
    kwx = (R)(50.0f * (float)ewx); kwy = (R)(50.0f * (float)ewy); kwz = (R)(50.0f * (float)ewz);
  }
  R dwx = -kp_a * erx - kwx, dwy = -kp_a * ery - kwy, dwz = -kp_a * erz - kwz;       // :350
  R thrust_mag = adx * s.r[2] + ady * s.r[5] + adz * s.r[8];            // :352
  jinv_mix<R>(m, thrust_mag, dwx, dwy, dwz, cmd);                          // :357-359
}

template <typename R, class T>
__device__ __forceinline__ void controller_cmds(const St<R>& s, const Params<R>& p, const Mdl<R, T::PEM>& m, int ctrl, const R* a,
                                                R* cmd) {
  if (ctrl == CTRL_RAW || ctrl == CTRL_VERTICAL || ctrl == CTRL_VERTPLANE) {   // :72-99, :116-136, :159-179
    R lo = p.zero_middle ? R(-1) : R(0);
    R b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      R v = p.zero_middle ? R(0.5) * (a[k] + R(1)) : R(1) * (a[k] + R(0));
      b[k] = clampr(v, lo, R(1));
    }
    if (ctrl == CTRL_RAW) { cmd[0] = b[0]; cmd[1] = b[1]; cmd[2] = b[2]; cmd[3] = b[3]; }
    else if (ctrl == CTRL_VERTICAL) { cmd[0] = cmd[1] = cmd[2] = cmd[3] = b[0]; }
    else { cmd[0] = cmd[1] = b[0]; cmd[2] = cmd[3] = b[1]; }
  } else if (ctrl == CTRL_SHIFTED) {                                    // :23-27
#pragma unroll
    for (int k = 0; k < 4; ++k) cmd[k] = clampr((a[k] + R(1)) / m.t2w(), R(0), R(1));
  } else if (ctrl == CTRL_OMEGA_THRUST) {                               // :225-234
    R dwx = R(-5) * (s.wx - a[1]), dwy = R(-5) * (s.wy - a[2]), dwz = R(-5) * (s.wz - a[3]);
    R acc_des = R(9.81) * (a[0] + R(1));
    jinv_mix<R>(m, acc_des, dwx, dwy, dwz, cmd);
  } else {
    attitude_controller<R, T::FASTM>(s, p, m, ctrl == CTRL_MELLINGER, a, cmd);
  }
}

// ---------------------------------------------------------------------------------------------------------
// reward (quadrotor.py:544-638; log variant quadrotor_multi.py:554)
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__device__ __forceinline__ R compute_reward(const St<R>& s, const Params<R>& p, const R* a, const R* a_prev,
                                            bool crashed, bool collided, R* costs) {
  const bool all = costs != nullptr;                                    // info buffer bound: every raw cost, as the reference
  R dx = s.gx - s.px, dy = s.gy - s.py, dz = s.gz - s.pz;
  R dist = q_sqrt<R, T::FASTM>(dx * dx + dy * dy + dz * dz);
  R r_pos = dist;                                                       // :547-549
  if (T::SWARM && p.log_pos) r_pos = p.c_pos_log_w * Mth<R>::log(dist + p.c_pos_offset) + p.c_pos_lin_w * dist;   // multi :554
  R c_pos = p.c_pos * r_pos;
  R r_eff = q_sqrt<R, T::FASTM>(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);   // :553-554
  R c_eff = p.c_effort * r_eff;
  R r_ac = R(0), c_ac = R(0);
  if (all || (!T::PLAIN && p.c_act_change != R(0))) {                   // :556-558
    R e0 = a[0] - a_prev[0], e1 = a[1] - a_prev[1], e2 = a[2] - a_prev[2], e3 = a[3] - a_prev[3];
    r_ac = Mth<R>::sqrt(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3);
    c_ac = p.c_act_change * r_ac;
  }
  R r_vel = R(0), c_vel = R(0);
  if (all || p.c_vel != R(0)) { r_vel = Mth<R>::sqrt(s.vx * s.vx + s.vy * s.vy + s.vz * s.vz); c_vel = p.c_vel * r_vel; }   // :562-563
  R c_or = p.c_orient * (-s.r[8]);                                      // :567-568
  R c_yaw = p.c_yaw * (-s.r[0]);                                        // :570-571
  R r_rot = R(0), c_rot = R(0), r_att = R(0), c_att = R(0);
  if (all || p.c_rot != R(0)) {                                         // :575-578
    r_rot = Mth<R>::acos(clampr(((s.r[0] + s.r[4] + s.r[8]) - R(1)) / R(2), R(-1), R(1)));
    c_rot = p.c_rot * r_rot;
  }
  if (all || p.c_attitude != R(0)) { r_att = Mth<R>::acos(clampr(s.r[8], R(-1), R(1))); c_att = p.c_attitude * r_att; }   // :580-581
  R r_spin = q_sqrt<R, T::FASTM>(s.wx * s.wx + s.wy * s.wy + s.wz * s.wz);                         // :585-586
  R c_spin = p.c_spin * r_spin;
  R c_crash = crashed ? p.c_crash : R(0);                               // :590-591
  R sum = ((c_pos + c_eff) + (c_crash + c_or)) + ((c_yaw + c_rot) + (c_att + c_spin));            // np.sum pairwise
  sum = (sum + c_ac) + c_vel;                                           // :593-604
  if (T::SWARM) { if (collided) sum += p.c_collision; }
  if (costs) {                                                          // rew_info order (:607-631): weighted, then raw
    costs[0] = c_pos; costs[1] = c_eff; costs[2] = c_crash; costs[3] = c_or; costs[4] = c_yaw;
    costs[5] = c_rot; costs[6] = c_att; costs[7] = c_spin; costs[8] = c_ac; costs[9] = c_vel;
    costs[10] = r_pos; costs[11] = r_eff; costs[12] = crashed ? R(1) : R(0); costs[13] = -s.r[8]; costs[14] = -s.r[0];
    costs[15] = r_rot; costs[16] = r_att; costs[17] = r_spin; costs[18] = r_ac; costs[19] = r_vel;
  }
  return -p.dt * sum;
}

// ---------------------------------------------------------------------------------------------------------
// noise sources: injected (host-supplied unit draws) or Philox
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
struct NoiseSrc {
  const Params<R>& p;
  long long i;
  Rng<R> rng;
  __device__ __forceinline__ NoiseSrc(const Params<R>& p_, long long i_) : p(p_), i(i_) {}
  __device__ __forceinline__ void begin_step(unsigned long long step) { if (!T::INJECT) rng.init(p, i, step); }

  __device__ __forceinline__ void ou(int sub, R* z) const {
    if (T::INJECT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) z[k] = p.ou_z ? p.ou_z[(i * p.sim_steps + sub) * 4 + k] : R(0);
    } else {
      rng.normal4(SLOT_OU + sub, z);
    }
  }
  // one 3-vector group of an add_noise call: g = slot/3 in the QS_Z_* layout
  __device__ __forceinline__ void obs3(bool reset_obs, int g, bool is_unif, R* z) const {
    if (T::INJECT) {
      const R* src = reset_obs ? p.reset_z : p.obs_z;
      long long base = reset_obs ? i * RESET_Z + 18 : i * OBS_Z;
#pragma unroll
      for (int k = 0; k < 3; ++k) z[k] = src ? src[base + g * 3 + k] : R(0);
    } else {
      R t[4];
      int slot = (reset_obs ? SLOT_RESET_OBS : SLOT_OBS) + g;
      if (g == 10) slot = reset_obs ? SLOT_RESET_T2W : SLOT_T2W;         // group 10: t2w, t2t (QS_Z_NT2W, QS_Z_NT2T)
      if (is_unif) rng.unif4(slot, t); else rng.normal4(slot, t);
      z[0] = t[0]; z[1] = t[1]; z[2] = t[2];
    }
  }
  __device__ __forceinline__ R reset_at(int k) const { return p.reset_z ? p.reset_z[i * RESET_Z + k] : R(0); }
  // moving-goal draws: excite (slots 0..2, one call per step) and resample_goal (slot 3, at reset)
  __device__ __forceinline__ void goal4(bool at_reset, R* u) const {
    if (T::INJECT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = p.goal_z ? p.goal_z[i * 4 + k] : R(0);
    } else {
      rng.unif4(at_reset ? SLOT_RESET_GOAL : SLOT_EXCITE, u);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------
// SensorNoise.add_noise (sensor_noise.py:100-159) + observation assembly (get_state.py)
// ---------------------------------------------------------------------------------------------------------
template <typename R, bool FAST>
__device__ __forceinline__ void rot2quat(const R* r, R* q) {            // sensor_noise.py:26-55
  R tr = r[0] + r[4] + r[8];
  if (FAST && sizeof(R) == 4) {
    // fp32: the four cases differ only in which diagonal combination goes under the root and where the entries land -
    // one MUFU.RSQ (S = 2 sqrt(t), 1 / S = rsqrt(t) / 2) and selects instead of four divergent branches with a square
    // root and three divisions each (a warp of 32 random attitudes takes all of them)
    const int c = tr > R(0) ? 0 : ((r[0] > r[4] && r[0] > r[8]) ? 1 : (r[4] > r[8] ? 2 : 3));
    const R t = c == 0 ? tr + R(1) : (c == 1 ? R(1) + r[0] - r[4] - r[8] : (c == 2 ? R(1) + r[4] - r[0] - r[8] : R(1) + r[8] - r[0] - r[4]));
    const R is = q_rsqrt<R, true>(t), h = R(0.5) * is, big = R(0.5) * (t * is);
    const R d0 = (r[7] - r[5]) * h, d1 = (r[2] - r[6]) * h, d2 = (r[3] - r[1]) * h;
    const R s01 = (r[1] + r[3]) * h, s02 = (r[2] + r[6]) * h, s12 = (r[5] + r[7]) * h;
    q[0] = c == 0 ? big : (c == 1 ? d0 : (c == 2 ? d1 : d2));
    q[1] = c == 0 ? d0 : (c == 1 ? big : (c == 2 ? s01 : s02));
    q[2] = c == 0 ? d1 : (c == 1 ? s01 : (c == 2 ? big : s12));
    q[3] = c == 0 ? d2 : (c == 1 ? s02 : (c == 2 ? s12 : big));
    return;
  }
  if (tr > R(0)) {
    if (FAST && sizeof(R) == 4) {                                       // the common branch: 1/S by MUFU.RSQ
      R t = tr + R(1);
      R is = q_rsqrt<R, true>(t);                                       // S = 2 sqrt(t), 1/S = is/2
      R h = R(0.5) * is;
      q[0] = R(0.5) * (t * is); q[1] = (r[7] - r[5]) * h; q[2] = (r[2] - r[6]) * h; q[3] = (r[3] - r[1]) * h;
      return;
    }
    R S = Mth<R>::sqrt(tr + R(1)) * R(2);
    q[0] = R(0.25) * S; q[1] = (r[7] - r[5]) / S; q[2] = (r[2] - r[6]) / S; q[3] = (r[3] - r[1]) / S;
  } else if (r[0] > r[4] && r[0] > r[8]) {
    R S = Mth<R>::sqrt(R(1) + r[0] - r[4] - r[8]) * R(2);
    q[0] = (r[7] - r[5]) / S; q[1] = R(0.25) * S; q[2] = (r[1] + r[3]) / S; q[3] = (r[2] + r[6]) / S;
  } else if (r[4] > r[8]) {
    R S = Mth<R>::sqrt(R(1) + r[4] - r[0] - r[8]) * R(2);
    q[0] = (r[2] - r[6]) / S; q[1] = (r[1] + r[3]) / S; q[2] = R(0.25) * S; q[3] = (r[5] + r[7]) / S;
  } else {
    R S = Mth<R>::sqrt(R(1) + r[8] - r[0] - r[4]) * R(2);
    q[0] = (r[3] - r[1]) / S; q[1] = (r[2] + r[6]) / S; q[2] = (r[5] + r[7]) / S; q[3] = R(0.25) * S;
  }
}

// q <- q x quat_from_small_angle(theta), theta = quat_norm_std * n + quat_unif_range * u (sensor_noise.py:9-21, :134-151,
// quad_utils.py:92-99)
template <typename R, class T>
__device__ __forceinline__ void noise_quat(const Params<R>& p, const NoiseSrc<R, T>& ns, bool reset_obs, R* q) {
  R zn[3] = {R(0), R(0), R(0)}, zu[3] = {R(0), R(0), R(0)};
  if (p.sn_quat_std != R(0)) ns.obs3(reset_obs, 6, false, zn);        // a zero scale needs no draw (every slot has its own counter)
  if (p.sn_quat_unif != R(0)) ns.obs3(reset_obs, 7, true, zu);
  R tx = p.sn_quat_std * zn[0] + p.sn_quat_unif * zu[0], ty = p.sn_quat_std * zn[1] + p.sn_quat_unif * zu[1];
  R tz = p.sn_quat_std * zn[2] + p.sn_quat_unif * zu[2];
  R qs = (tx * tx + ty * ty + tz * tz) / R(4);                          // sensor_noise.py:9-21
  R t0, f;
  if (qs < R(1)) { t0 = q_sqrt<R, T::FASTM>(R(1) - qs); f = R(0.5); }
  else { t0 = q_rsqrt<R, T::FASTM>(R(1) + qs); f = R(0.5) * t0; }
  R t1 = tx * f, t2 = ty * f, t3 = tz * f;
  if (T::FASTM && sizeof(R) == 4) {                                     // fp32: normalise by MUFU.RSQ (norm^2 is 1 + O(theta^2))
    R inv = q_rsqrt<R, true>(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3);
    t0 *= inv; t1 *= inv; t2 *= inv; t3 *= inv;
  } else {
    R tn = Mth<R>::sqrt(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3);
    t0 /= tn; t1 /= tn; t2 /= tn; t3 /= tn;
  }
  R n0 = q[0] * t0 - q[1] * t1 - q[2] * t2 - q[3] * t3;                 // quad_utils.py:92-99
  R n1 = q[0] * t1 + q[1] * t0 - q[2] * t3 + q[3] * t2;
  R n2 = q[0] * t2 + q[1] * t3 + q[2] * t0 - q[3] * t1;
  R n3 = q[0] * t3 - q[1] * t2 + q[2] * t1 + q[3] * t0;
  q[0] = n0; q[1] = n1; q[2] = n2; q[3] = n3;
}

template <typename R, class T>
__device__ __forceinline__ void build_obs(St<R>& s, const Params<R>& p, const Mdl<R, T::PEM>& m, const NoiseSrc<R, T>& ns,
                                          bool reset_obs, const R* act_prev, R* o, const R* zpre = nullptr) {
  const bool sense = T::SENSE >= 0 ? (T::SENSE != 0) : (p.sense_on != 0);
  R px = s.px, py = s.py, pz = s.pz, vx = s.vx, vy = s.vy, vz = s.vz, wx = s.wx, wy = s.wy, wz = s.wz;
  R rr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) rr[k] = s.r[k];
  R acc[3] = {R(0), R(0), R(0)};
  const bool obs_acc = !T::PLAIN && p.obs_acc, obs_rel = !T::PLAIN && p.obs_rel;
  const bool obs_quat = T::MISC && p.obs_quat;
  if (obs_acc) accelerometer(s, p, acc);
  R qo[4] = {R(1), R(0), R(0), R(0)};
  if (obs_quat) {                                                       // R2quat, quad_utils.py:101-109 (single branch)
    const R t = R(1) + s.r[0] + s.r[4] + s.r[8];
    if (T::FASTM && sizeof(R) == 4 && t > R(1e-12)) {                   // fp32: w = sqrt(t) / 2 and 1 / (4 w) from one MUFU.RSQ
      const R is = q_rsqrt<R, true>(t), i4 = R(0.5) * is;
      qo[0] = R(0.5) * (t * is); qo[1] = (s.r[7] - s.r[5]) * i4; qo[2] = (s.r[2] - s.r[6]) * i4; qo[3] = (s.r[3] - s.r[1]) * i4;
    } else {                                                            // (t <= 0: the reference's own 0 / 0 and NaN cases)
      R w = Mth<R>::sqrt(t) / R(2);
      R w4 = R(4) * w;
      qo[0] = w; qo[1] = (s.r[7] - s.r[5]) / w4; qo[2] = (s.r[2] - s.r[6]) / w4; qo[3] = (s.r[3] - s.r[1]) / w4;
    }
  }
  if (sense) {
    R z[3];
    if (zpre) { z[0] = zpre[0]; z[1] = zpre[1]; z[2] = zpre[2]; } else
    ns.obs3(reset_obs, 0, false, z);                                    // :116-118
    px += p.sn_pos_std * z[0]; py += p.sn_pos_std * z[1]; pz += p.sn_pos_std * z[2];
    if (T::NOISEX && p.sn_pos_unif != R(0)) {
      ns.obs3(reset_obs, 1, true, z);
      px += p.sn_pos_unif * z[0]; py += p.sn_pos_unif * z[1]; pz += p.sn_pos_unif * z[2];
    }
    if (zpre) { z[0] = zpre[3]; z[1] = zpre[4]; z[2] = zpre[5]; } else
    ns.obs3(reset_obs, 2, false, z);                                    // :122-124
    vx += p.sn_vel_std * z[0]; vy += p.sn_vel_std * z[1]; vz += p.sn_vel_std * z[2];
    if (T::NOISEX && p.sn_vel_unif != R(0)) {
      ns.obs3(reset_obs, 3, true, z);
      vx += p.sn_vel_unif * z[0]; vy += p.sn_vel_unif * z[1]; vz += p.sn_vel_unif * z[2];
    }
    if (zpre) { z[0] = zpre[6]; z[1] = zpre[7]; z[2] = zpre[8]; } else
    ns.obs3(reset_obs, 4, false, z);
    if (T::NOISEX && p.gyro_bias_mode) {                                  // :127-128, :162-170 (walk folded over the calls)
      R pi_ = reset_obs ? p.gb_pi1 : p.gb_pi3, sg = reset_obs ? p.gb_sig1 : p.gb_sig3;
      s.gb[0] = pi_ * s.gb[0] + sg * z[0]; s.gb[1] = pi_ * s.gb[1] + sg * z[1]; s.gb[2] = pi_ * s.gb[2] + sg * z[2];
      R z2[3];
      ns.obs3(reset_obs, 5, false, z2);
      wx = wx + s.gb[0] + p.sn_gyro_rw * z2[0]; wy = wy + s.gb[1] + p.sn_gyro_rw * z2[1]; wz = wz + s.gb[2] + p.sn_gyro_rw * z2[2];
    } else {
      wx += p.sn_gyro_density * z[0]; wy += p.sn_gyro_density * z[1]; wz += p.sn_gyro_density * z[2];   // :130-131
    }
    // rotation: R -> quaternion -> (x small-angle quaternion) -> R (:134-147).  Without quaternion noise (the default,
    // and the only case in the fast subset) the float64 reference's round trip returns R to 1e-15; an fp32 round trip
    // would only add rounding, so the fast kernels pass R through.
    const bool quat_noise = T::NOISEX && (p.sn_quat_std != R(0) || p.sn_quat_unif != R(0));
    if (sizeof(R) == 8 || !T::FASTM || quat_noise || obs_quat) {
      if (obs_quat) {                                                   // rot.shape == (4,): quat x noise quaternion (:148-151)
        if (quat_noise) noise_quat<R, T>(p, ns, reset_obs, qo);
      } else {
        R q[4];
        rot2quat<R, T::FASTM>(s.r, q);
        if (quat_noise) noise_quat<R, T>(p, ns, reset_obs, q);
        R qw = q[0], qx = q[1], qy = q[2], qz = q[3];                   // quad_utils.py:82-87
        rr[0] = R(1) - R(2) * qy * qy - R(2) * qz * qz; rr[1] = R(2) * qx * qy - R(2) * qz * qw; rr[2] = R(2) * qx * qz + R(2) * qy * qw;
        rr[3] = R(2) * qx * qy + R(2) * qz * qw; rr[4] = R(1) - R(2) * qx * qx - R(2) * qz * qz; rr[5] = R(2) * qy * qz - R(2) * qx * qw;
        rr[6] = R(2) * qx * qz - R(2) * qy * qw; rr[7] = R(2) * qy * qz + R(2) * qx * qw; rr[8] = R(1) - R(2) * qx * qx - R(2) * qy * qy;
      }
    }
    if (obs_acc) {                                                      // :156-157
      R zs[3], zd[3];
      ns.obs3(reset_obs, 8, false, zs);
      ns.obs3(reset_obs, 9, false, zd);
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[k] = acc[k] + p.sn_acc_static * zs[k] + acc[k] * (p.sn_acc_dyn * zd[k]);
    }
  }
  R ex = px - s.gx, ey = py - s.gy, ez = pz - s.gz;                     // get_state.py:15
  if (obs_rel) {                                                        // get_state.py:159-160 (true R)
    R tx = s.r[0] * ex + s.r[3] * ey + s.r[6] * ez, ty = s.r[1] * ex + s.r[4] * ey + s.r[7] * ez;
    R tz = s.r[2] * ex + s.r[5] * ey + s.r[8] * ez;
    ex = tx; ey = ty; ez = tz;
    tx = s.r[0] * vx + s.r[3] * vy + s.r[6] * vz; ty = s.r[1] * vx + s.r[4] * vy + s.r[7] * vz;
    tz = s.r[2] * vx + s.r[5] * vy + s.r[8] * vz;
    vx = tx; vy = ty; vz = tz;
  }
  o[0] = ex; o[1] = ey; o[2] = ez; o[3] = vx; o[4] = vy; o[5] = vz;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[6 + k] = rr[k];
  o[15] = wx; o[16] = wy; o[17] = wz;
  // tail (h | acc+act | act) is written by the caller from these scratch slots
  o[18] = pz;
  o[19] = acc[0]; o[20] = acc[1]; o[21] = acc[2];
  o[22] = act_prev[0]; o[23] = act_prev[1]; o[24] = act_prev[2]; o[25] = act_prev[3];
  if (T::MISC) {
    o[26] = qo[0]; o[27] = qo[1]; o[28] = qo[2]; o[29] = qo[3];
    o[30] = R(0); o[31] = R(0);
    if (p.obs_t2w) {            // get_state.py:334-337, :379-382: drawn by the builder itself, SensorNoise bypassed or not
      R z[3];
      ns.obs3(reset_obs, 10, false, z);
      R w = m.t2w(), sc = p.t2w_half_std * w;
      R nw = clampr(w + (sc < R(0) ? -sc : sc) * z[0], R(1.5), R(10));              // quadrotor.py:707-708
      o[30] = (T::FASTM && sizeof(R) == 4) ? (nw - R(1.5)) * R(1.0 / 8.5) : (nw - R(1.5)) / (R(10) - R(1.5));
      if (p.obs_t2t) {
        w = m.t2t(); sc = p.t2t_half_std * w;
        nw = clampr(w + (sc < R(0) ? -sc : sc) * z[1], R(0.005), R(1));             // quadrotor.py:711-712
        o[31] = (T::FASTM && sizeof(R) == 4) ? (nw - R(0.005)) * R(1.0 / 0.995) : (nw - R(0.005)) / (R(1) - R(0.005));
      }
    }
  }
}

// writes one env's observation row into the CTA's shared tile in final [obs_dim] order
template <typename R, bool PLAIN = false, bool EXT = true>
__device__ __forceinline__ void obs_to_tile(const Params<R>& p, const R* o, R* row) {
  if (PLAIN) {
#pragma unroll
    for (int k = 0; k < 18; ++k) row[k] = o[k];
    return;
  }
  int n = 18;
  if (EXT && p.obs_quat) {      // xyz vxyz quat omega: the quaternion takes the place of the nine matrix entries
#pragma unroll
    for (int k = 0; k < 6; ++k) row[k] = o[k];
    row[6] = o[26]; row[7] = o[27]; row[8] = o[28]; row[9] = o[29];
    row[10] = o[15]; row[11] = o[16]; row[12] = o[17];
    n = 13;
  } else {
#pragma unroll
    for (int k = 0; k < 18; ++k) row[k] = o[k];
  }
  if (p.obs_h) row[n++] = o[18];
  if (p.obs_acc) { row[n] = o[19]; row[n + 1] = o[20]; row[n + 2] = o[21]; n += 3; }
  if (p.obs_act) { row[n] = o[22]; row[n + 1] = o[23]; row[n + 2] = o[24]; row[n + 3] = o[25]; n += 4; }
  if (EXT && p.obs_t2w) row[n++] = o[30];
  if (EXT && p.obs_t2t) row[n++] = o[31];
}

// coalesced copy of the CTA's [BLOCK][obs_dim] tile to global memory
template <typename R>
__device__ __forceinline__ void flush_tile(const Params<R>& p, const R* tile, long long block_first, const uint8_t* row_valid) {
  long long rows = p.end - block_first;
  if (rows > BLOCK) rows = BLOCK;
  int total = (int)rows * p.obs_dim;
  QS_CHECK(p, rows >= 0 && block_first >= p.begin && block_first + rows <= p.end, BC_OBS_GLOBAL);
  R* dst = p.obs + block_first * p.obs_dim;
  if (row_valid == nullptr && (total & 3) == 0 && ((p.obs_dim * BLOCK) & 3) == 0) {
    const V4<R>* t4 = reinterpret_cast<const V4<R>*>(tile);
    V4<R>* d4 = reinterpret_cast<V4<R>*>(dst);
    for (int k = threadIdx.x; k < (total >> 2); k += BLOCK) d4[k] = t4[k];
  } else {
    for (int k = threadIdx.x; k < total; k += BLOCK) {
      if (row_valid == nullptr || row_valid[k / p.obs_dim]) dst[k] = tile[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// QuadrotorEnv._reset (quadrotor.py:1059-1144) for one env
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__device__ __forceinline__ void reset_env(St<R>& s, const Params<R>& p, const NoiseSrc<R, T>& ns) {
  const R PI = R(3.14159265358979323846);
  if (T::MISC) {
    if (p.resample_goal || p.excite) {                                  // :1078-1081: the goal is re-made at every reset
      s.gx = p.goal[0]; s.gy = p.goal[1]; s.gz = p.goal[2];
      if (p.resample_goal) {
        R g[4];
        ns.goal4(true, g);
        s.gx = R(0); s.gy = R(0); s.gz = R(1.25) + R(0.75) * g[3];     // np.random.uniform(0.5, 2.0)
      }
    }
  }
  R u[4];
  if (T::INJECT) { u[0] = ns.reset_at(0); u[1] = ns.reset_at(1); u[2] = ns.reset_at(2); u[3] = ns.reset_at(3); }
  else ns.rng.unif4(SLOT_RESET_POS, u);
  R x = p.spawn_box * u[0] + s.gx, y = p.spawn_box * u[1] + s.gy, z = p.spawn_box * u[2] + s.gz;   // :1087
  if (T::MISC) {
    if (p.dim_mode == 1) { x = s.gx; y = s.gy; }                        // :1089-1092
    else if (p.dim_mode == 2) { y = s.gy; }
  }
  if (z < p.spawn_z_min) z = p.spawn_z_min;                             // :1094 (0.25)
  s.px = x; s.py = y; s.pz = z;
  const bool restricted = T::MISC && p.init_random == INIT_PITCH_ROLL_RESTRICTED;
  if (T::MISC && p.init_random && p.dim_mode != 3) {                    // :1101-1110 (rand() = (1 + u) / 2 of a unit draw)
    R a[4], b[4];
    if (T::INJECT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = ns.reset_at(4 + k); b[k] = ns.reset_at(8 + k); }
    } else {
      ns.rng.unif4(SLOT_RESET_VEL, a);
      ns.rng.unif4(SLOT_RESET_OM, b);
    }
    R vm = p.max_init_vel, om = p.max_init_omega;
    s.wx = s.wy = s.wz = R(0);
    s.vx = s.vy = R(0);
    R sp = R(0), cp = R(1);
    if (p.dim_mode == 1) {
      s.vz = vm * (R(0.5) + R(0.5) * a[3]);                             // :1103
    } else {
      s.wy = om * (R(0.5) + R(0.5) * b[3]);                             // :1105
      s.vx = vm * (R(0.5) + R(0.5) * a[0]); s.vz = vm * (R(0.5) + R(0.5) * a[2]);   // :1106-1107
      Mth<R>::sincos(PI * (R(0.5) + R(0.5) * u[3]), &sp, &cp);          // :1108-1109
    }
    s.r[0] = cp; s.r[1] = R(0); s.r[2] = -sp;                           // :1110
    s.r[3] = R(0); s.r[4] = R(1); s.r[5] = R(0);
    s.r[6] = sp; s.r[7] = R(0); s.r[8] = cp;
  } else if (p.init_random) {                                           // :1113-1115 -> :227-239 / :243-259
    R a[4], b[4], up[4], fw[4];
    if (T::INJECT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = ns.reset_at(4 + k); b[k] = ns.reset_at(8 + k); }
#pragma unroll
      for (int k = 0; k < 3; ++k) { up[k] = ns.reset_at(12 + k); fw[k] = ns.reset_at(15 + k); }
    } else {
      ns.rng.unif4(SLOT_RESET_VEL, a);
      ns.rng.unif4(SLOT_RESET_OM, b);
      if (restricted) ns.rng.unif4(SLOT_RESET_UP, up); else ns.rng.normal4(SLOT_RESET_UP, up);
    }
    R vm = p.max_init_vel, om = p.max_init_omega;
    R vx = vm * a[0], vy = vm * a[1], vz = vm * a[2];                   // :230
    R vmag = vm * (R(0.5) + R(0.5) * a[3]);                             // :231
    R f = vmag / (Mth<R>::sqrt(vx * vx + vy * vy + vz * vz) + R(1e-6)); // :232
    s.vx = f * vx; s.vy = f * vy; s.vz = f * vz;
    R wx = om * b[0], wy = om * b[1], wz = om * b[2];                   // :234
    R wmag = om * (R(0.5) + R(0.5) * b[3]);                             // :235
    f = wmag / (Mth<R>::sqrt(wx * wx + wy * wy + wz * wz) + R(1e-6));   // :236
    s.wx = f * wx; s.wy = f * wy; s.wz = f * wz;
    if (restricted) {           // :254-257: pitch, roll, yaw ~ U(-max, max); euler2mat 'sxyz' = Rz(yaw) Ry(pitch) Rx(roll)
      R si, ci, sj, cj, sk, ck;
      Mth<R>::sincos(p.roll_max * up[1], &si, &ci);
      Mth<R>::sincos(p.pitch_max * up[0], &sj, &cj);
      Mth<R>::sincos(p.yaw_max * up[2], &sk, &ck);
      s.r[0] = cj * ck; s.r[1] = sj * si * ck - ci * sk; s.r[2] = sj * ci * ck + si * sk;
      s.r[3] = cj * sk; s.r[4] = sj * si * sk + ci * ck; s.r[5] = sj * ci * sk - si * ck;
      s.r[6] = -sj; s.r[7] = cj * si; s.r[8] = cj * ci;
    } else {
    R ux = up[0], uy = up[1], uz = up[2];                               // quad_utils.py:47-57
    normalize3(ux, uy, uz);
    R fx, fy, fz;
    if (T::INJECT) { fx = fw[0]; fy = fw[1]; fz = fw[2]; normalize3(fx, fy, fz); }
    else {
#pragma unroll 1
      for (int tr = 0; tr < 8; ++tr) {                                  // rejection loop :51-52 (p(reject) = 2.5 %)
        ns.rng.normal4(SLOT_RESET_FWD + tr, fw);
        fx = fw[0]; fy = fw[1]; fz = fw[2];
        normalize3(fx, fy, fz);
        if (!(fx * ux + fy * uy + fz * uz > R(0.95))) break;
      }
    }
    R lx = uy * fz - uz * fy, ly = uz * fx - ux * fz, lz = ux * fy - uy * fx;   // left = normalize(cross(up, fwd))
    normalize3(lx, ly, lz);
    ux = fy * lz - fz * ly; uy = fz * lx - fx * lz; uz = fx * ly - fy * lx;     // up = cross(fwd, left)
    s.r[0] = fx; s.r[1] = lx; s.r[2] = ux;                              // column_stack([fwd, left, up])
    s.r[3] = fy; s.r[4] = ly; s.r[5] = uy;
    s.r[6] = fz; s.r[7] = lz; s.r[8] = uz;
    }
  } else {                                                              // :1118-1126
    s.vx = s.vy = s.vz = R(0);
    s.wx = s.wy = s.wz = R(0);
    R psi;
    if (T::INJECT) psi = PI * u[3];                                     // the accepted randyaw() draw
    else psi = Mth<R>::atan2(-y, -x) + (PI / R(3)) * u[3];              // same law as the rejection loop :1124-1126
    if (T::MISC) { if (p.dim_mode != 3) psi = R(0); }                   // :1120-1121
    R sp, cp;
    Mth<R>::sincos(psi, &sp, &cp);
    s.r[0] = cp; s.r[1] = -sp; s.r[2] = R(0);
    s.r[3] = sp; s.r[4] = cp; s.r[5] = R(0);
    s.r[6] = R(0); s.r[7] = R(0); s.r[8] = R(1);
  }
  // set_state rounds omega to float32 (:223)
  s.wx = (R)(float)s.wx; s.wy = (R)(float)s.wy; s.wz = (R)(float)s.wz;
  s.w_is_f32 = true;
  s.fresh = true;
  s.ax = s.ay = s.az = R(0);                                            // :220
#pragma unroll
  for (int k = 0; k < 4; ++k) { s.rd[k] = R(0); s.cd[k] = R(0); s.pa[k] = R(0); }   // :438-440, :1141
  s.tick = R(0);                                                        // :1140
  s.ret = R(0);
}

// ---------------------------------------------------------------------------------------------------------
// swarm contact (builder-authored spec, no reference counterpart; see oracle collide_swarm and DESIGN.md).
// M = n_agents consecutive lanes form one swarm env; each lane visits its M-1 partners by shuffle and sums the
// impulses computed from the PRE-collision velocities, so the result is independent of pair order.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ R shfl(R v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// half_thr: this drone's half of the contact distance, collision_k * arm.  One model per handle: every pair touches at
// collision_k * 2 * arm (p.coll_dist).  Per-env models (PEM): drones of one swarm may have different arms, the pair (i, j)
// touches at collision_k * (arm_i + arm_j) - the partner's half arrives by shuffle.
// FAST (fp32 throughput kernels): the distance by MUFU.RSQ (d2 * rsqrt(d2), ~2 ulp) and its reciprocal from the same rsqrt -
// no IEEE square root or division in the pair loop.  Late in an episode under random actions the drones of a swarm lie in contact
// on the floor and every pair takes the contact branch: the IEEE forms made those steps 28 % slower than the early ones
// (tools/phase_of_episode.py: 120 -> 153 us per 2,097,152-drone step).
template <typename R, bool PEM, bool FAST = false>
__device__ __forceinline__ bool swarm_collide(St<R>& s, const Params<R>& p, R half_thr) {
  const int M = p.n_agents;
  const int lane = threadIdx.x & 31;
  const int base = lane & ~(M - 1);
  R dvx = R(0), dvy = R(0), dvz = R(0);
  bool hit = false;
  auto pair = [&](int j) {
    int src = base | ((lane + j) & (M - 1));
    R qx = shfl(s.px, src), qy = shfl(s.py, src), qz = shfl(s.pz, src);
    R dx = s.px - qx, dy = s.py - qy, dz = s.pz - qz;
    const R d2 = dx * dx + dy * dy + dz * dz;
    R dist, is = R(0);
    if (FAST && sizeof(R) == 4) { is = q_rsqrt<R, true>(Mth<R>::max(d2, R(1e-30))); dist = d2 * is; }
    else dist = Mth<R>::sqrt(d2);
    bool c = dist < (PEM ? half_thr + shfl(half_thr, src) : p.coll_dist);
    hit = hit || c;
    // contacts are rare: the partner's velocity is fetched (warp-convergent shuffles) only when some lane needs it
    if (__any_sync(0xffffffffu, c)) {
      R ux = shfl(s.vx, src), uy = shfl(s.vy, src), uz = shfl(s.vz, src);
      R inv;
      if (FAST && sizeof(R) == 4) inv = dist > R(1e-5) ? is : R(1e5);     // 1 / max(dist, 1e-5): is = 1 / dist wherever d2 >= 1e-30
      else inv = R(1) / Mth<R>::max(dist, R(1e-5));
      R nx = dx * inv, ny = dy * inv, nz = dz * inv;
      R delta = (ux - s.vx) * nx + (uy - s.vy) * ny + (uz - s.vz) * nz;
      if (c && delta > R(0)) { dvx += delta * nx; dvy += delta * ny; dvz += delta * nz; }
    }
  };
  // (unrolling the seven partners of the BASELINE swarm size was measured slower: 117 -> 123 us early, 123 -> 133 us late)
#pragma unroll 1
  for (int j = 1; j < M; ++j) pair(j);
  s.vx += dvx; s.vy += dvy; s.vz += dvz;
  return hit;
}

// ---------------------------------------------------------------------------------------------------------
// QuadrotorEnv.step (quadrotor.py:942-1028) for one env held in registers, k_steps times, + folded auto-reset.
// Shared by the direct-load kernel and the TMA-pipelined kernel below.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct StepOut {
  R rew_sum;
  bool done_any;
  double st_ret, st_ret2, st_n;   // finished-episode statistics of this thread
  double svd_t;                   // value of the global re-orthonormalisation timer after the last substep
};

// Draws of one env step that depend only on (global env id, step index): generated BEFORE the state is needed so
// that the Philox / Box-Muller chains overlap the loads in flight (kernels with compile-time sim_steps, k_steps == 1)
template <typename R, class T>
struct Pre {
  static constexpr int N = T::SIM > 0 ? T::SIM : 1;
  R zou[N][4];
  R zobs[9];
};

template <typename R, class T>
__device__ __forceinline__ void predraw(const Params<R>& p, long long i, bool active, unsigned long long step0, Pre<R, T>& pre) {
  NoiseSrc<R, T> ns(p, active ? i : 0);
  ns.begin_step(step0);
  const bool ou_on = T::OU >= 0 ? (T::OU != 0) : (p.ou_on != 0);
  const bool sense = T::SENSE >= 0 ? (T::SENSE != 0) : (p.sense_on != 0);
  if (ou_on) {
#pragma unroll
    for (int sub = 0; sub < Pre<R, T>::N; ++sub) ns.ou(sub, pre.zou[sub]);
  }
  if (sense) { ns.obs3(false, 0, false, pre.zobs); ns.obs3(false, 2, false, pre.zobs + 3); ns.obs3(false, 4, false, pre.zobs + 6); }
}

template <typename R, class T>
__device__ __forceinline__ void run_steps(St<R>& s, const Params<R>& p, long long i, bool active, const R* a_first,
                                          double svd_t0, unsigned long long step0, R* o, StepOut<R>& out,
                                          const Pre<R, T>* pre = nullptr) {
  const int ctrl = T::CTRL >= 0 ? T::CTRL : p.controller;
  NoiseSrc<R, T> ns(p, active ? i : 0);
  const Mdl<R, T::PEM> m(p, active ? i : p.begin);
  double svd_t = svd_t0;
  R rew_sum = R(0);
  bool done_any = false;
  double st_ret = 0.0, st_ret2 = 0.0, st_n = 0.0;

  const int k_steps = T::PLAIN ? 1 : p.k_steps;
  const bool ou_on = T::OU >= 0 ? (T::OU != 0) : (p.ou_on != 0);
  const bool sense = T::SENSE >= 0 ? (T::SENSE != 0) : (p.sense_on != 0);
#pragma unroll 1
  for (int ks = 0; ks < k_steps; ++ks) {
    ns.begin_step(step0 + (unsigned long long)ks);
    // Known sim_steps: every draw of the step depends only on (env id, step index), so generate them up front - the
    // Philox / Box-Muller chains then overlap the state loads still in flight and each other instead of sitting on
    // the dynamics' critical path.
    constexpr int NPRE = T::SIM > 0 ? T::SIM : 1;
    constexpr bool PRE = T::SIM > 0 && T::PLAIN;                        // k_steps == 1: the caller drew everything already
    // (per-env models + F_MISC run out of registers even at 168: eight more live draws would be spilled, and spills are what
    // these kernels cannot afford - they draw inside the unrolled substeps instead)
    constexpr bool HOIST = T::SIM > 0 && !PRE && !(T::PEM && T::MISC);
    Pre<R, T> own;
    if (HOIST && ou_on) {
#pragma unroll
      for (int sub = 0; sub < NPRE; ++sub) ns.ou(sub, own.zou[sub]);
    }
    const Pre<R, T>& dr = PRE ? *pre : own;
    // ---- action (:943-944) ----
    R a[4] = {R(0), R(0), R(0), R(0)};
    if (a_first != nullptr && ks == 0) {
      a[0] = a_first[0]; a[1] = a_first[1]; a[2] = a_first[2]; a[3] = a_first[3];
    } else if (active && p.actions != nullptr) {
      QS_CHECK(p, i >= 0 && i < p.n && ks < p.k_steps && p.act_dim >= 1 && p.act_dim <= 4, BC_ACTION);
      const R* ap = p.actions + ((long long)ks * p.n + i) * p.act_dim;
      if (p.act_dim == 4) { V4<R> v = *reinterpret_cast<const V4<R>*>(ap); a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < p.act_dim) a[k] = ap[k];
      }
    }
    R a_prev[4] = {s.pa[0], s.pa[1], s.pa[2], s.pa[3]};                 // actions[1]
    if (T::MISC) {
      if (p.excite) {                                                   // :957-962: a new goal every 5th tick
        const int tk = (int)s.tick;
        if (tk % 5 == 0) {
          R g[4];
          ns.goal4(false, g);
          s.gx = R(0.5) * g[0]; s.gy = R(0.5) * g[1]; s.gz = R(2) + R(0.5) * g[2];
        }
      }
    }
    // ---- controller + dynamics (:966-970, :261-262) ----
    R cmd[4];
    controller_cmds<R, T>(s, p, m, ctrl, a, cmd);
    if (T::SIM > 0) {
#pragma unroll
      for (int sub = 0; sub < NPRE; ++sub) {
        const bool fire = svd_tick(svd_t, p.dt_d, p.svd_period);        // :381-386
        if (T::SIM > 0 && !PRE && !HOIST && ou_on) ns.ou(sub, own.zou[sub]);
        substep<R, T>(s, cmd, p, m, dr.zou[sub], fire);
      }
    } else {
#pragma unroll 1
      for (int sub = 0; sub < p.sim_steps; ++sub) {
        const bool fire = svd_tick(svd_t, p.dt_d, p.svd_period);        // :381-386
        R z[4] = {R(0), R(0), R(0), R(0)};
        if (ou_on) ns.ou(sub, z);
        substep<R, T>(s, cmd, p, m, z, fire);
      }
    }
    bool collided = false;
    if (T::SWARM) { if (p.n_agents > 1) collided = swarm_collide<R, T::PEM, T::FASTM>(s, p, T::PEM ? p.coll_k * m.arm() : R(0)); }
    // ---- crash, reward, tick, done (:974-987) ----
    bool crashed = s.pz <= m.arm();                                     // :977
    R costs[20];
    R rew = compute_reward<R, T>(s, p, a, a_prev, crashed, collided, (T::INFO && p.info) ? costs : nullptr);
    if (!Mth<R>::finite(rew) && active) atomicOr(p.status, 1);          // :633-636
    s.tick += R(1);                                                     // :986
    bool done = s.tick > (R)p.ep_len;                                   // :987
    s.ret += rew;
    rew_sum += rew;
    done_any = done_any || done;
#pragma unroll
    for (int k = 0; k < 4; ++k) s.pa[k] = a[k];                         // becomes actions[1] of the next step
    const bool last = (ks == k_steps - 1);
    if (T::INFO) {
      if (p.info != nullptr && last && active) {
        R inf[INFO_DIM];                   // assembled in registers, stored as twelve 4-wide vectors (a row is 48 reals)
        R am[3];
        accelerometer(s, p, am);
        inf[0] = s.ax; inf[1] = s.ay; inf[2] = s.az; inf[3] = am[0]; inf[4] = am[1]; inf[5] = am[2];
        inf[6] = s.tqx; inf[7] = s.tqy; inf[8] = s.tqz; inf[9] = s.odx; inf[10] = s.ody; inf[11] = s.odz;
        inf[12] = cmd[0]; inf[13] = cmd[1]; inf[14] = cmd[2]; inf[15] = cmd[3];
        inf[16] = crashed ? R(1) : R(0); inf[17] = collided ? R(1) : R(0);
#pragma unroll
        for (int k = 0; k < 10; ++k) { inf[18 + k] = costs[k]; inf[32 + k] = costs[10 + k]; }
        inf[28] = s.ret; inf[29] = s.tick; inf[30] = rew; inf[31] = done ? R(1) : R(0);
#pragma unroll
        for (int k = 42; k < INFO_DIM; ++k) inf[k] = R(0);
        QS_CHECK(p, i >= 0 && i < p.n, BC_INFO);
        V4<R>* dst = reinterpret_cast<V4<R>*>(p.info + i * INFO_DIM);
#pragma unroll
        for (int k = 0; k < INFO_DIM / 4; ++k) { V4<R> v; v.x = inf[4 * k]; v.y = inf[4 * k + 1]; v.z = inf[4 * k + 2]; v.w = inf[4 * k + 3]; dst[k] = v; }
      }
    }
    // ---- observation (:988) and episode end: statistics + folded reset (:1059-1144) ----
    const bool walk = T::NOISEX && p.gyro_bias_mode;                      // the gyro-bias walk advances on every call
    const bool resetting = done && p.auto_reset;
    const bool want_term = resetting && p.term_obs != nullptr;          // the obs the reference returns on its done step (:988)
    if ((last && (!resetting || want_term)) || walk) build_obs<R, T>(s, p, m, ns, false, a_prev, o, (PRE && sense) ? dr.zobs : nullptr);
    if (done) {
      if (active && s.tick == (R)(p.ep_len + 1)) { st_ret += (double)s.ret; st_ret2 += (double)s.ret * (double)s.ret; st_n += 1.0; }
      if (resetting) {
        QS_CHECK(p, !(want_term && last && active) || (i >= 0 && i < p.n), BC_INFO);
        if (want_term && last && active) obs_to_tile<R, T::PLAIN, T::MISC>(p, o, p.term_obs + i * (T::PLAIN ? 18 : p.obs_dim));
        reset_env<R, T>(s, p, ns);
        if (last || walk) build_obs<R, T>(s, p, m, ns, true, s.pa, o);     // :1143 (actions[1] = 0 after reset)
      }
    }
  }
  out.rew_sum = rew_sum; out.done_any = done_any;
  out.st_ret = st_ret; out.st_ret2 = st_ret2; out.st_n = st_n; out.svd_t = svd_t;
}

// finished-episode statistics of a CTA: warp shuffle -> shared -> one double atomic triple (called by all threads)
template <typename R>
__device__ __forceinline__ void block_stats(const Params<R>& p, double (*red)[BLOCK / 32], double st_ret, double st_ret2, double st_n) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    st_ret += __shfl_down_sync(0xffffffffu, st_ret, off);
    st_ret2 += __shfl_down_sync(0xffffffffu, st_ret2, off);
    st_n += __shfl_down_sync(0xffffffffu, st_n, off);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = st_ret; red[1][threadIdx.x >> 5] = st_ret2; red[2][threadIdx.x >> 5] = st_n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) { a0 += red[0][w]; a1 += red[1][w]; a2 += red[2][w]; }
    if (a2 > 0) { atomicAdd(&p.stats[0], a0); atomicAdd(&p.stats[1], a1); atomicAdd(&p.stats[2], a2); }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// kernel 1 (generic): one CTA per 128 envs, direct vector loads.  Used by the fp64 build and by every
// configuration outside the fast subset.
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__global__ void __launch_bounds__(BLOCK) env_step_kernel(const __grid_constant__ Params<R> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  R* tile = reinterpret_cast<R*>(smem_raw);                             // [BLOCK][obs_dim]
  __shared__ double red[3][BLOCK / 32];

  const long long block_first = p.begin + (long long)blockIdx.x * BLOCK;
  const long long i = block_first + threadIdx.x;
  const bool active = i < p.end;

  St<R> s;
  load_state<R, T>(s, p, i, active);
  R o[OBS_SCRATCH];
  StepOut<R> out;
  run_steps<R, T>(s, p, i, active, nullptr, p.clock->svd_timer, p.clock->step, o, out);

  // ---- write back: state (float4 rows), obs (shared-memory transpose), reward, done ----
  if (active) {
    store_state<R, T>(s, p, i);
    p.rew[i] = out.rew_sum;
    p.done[i] = out.done_any ? 1 : 0;
    obs_to_tile(p, o, tile + threadIdx.x * p.obs_dim);
  }
  const bool any = __syncthreads_or(out.done_any ? 1 : 0);              // also orders the tile writes
  finish_clock(p, out.svd_t, (unsigned long long)p.k_steps);            // every thread of the CTA has read the clock
  flush_tile(p, tile, block_first, nullptr);
  if (any) block_stats(p, red, out.st_ret, out.st_ret2, out.st_n);
}

// ---------------------------------------------------------------------------------------------------------
// kernel 2 (fp32 fast subset): persistent CTAs, TMA-pipelined.
//   * grid = min(tiles, resident CTA slots); each CTA walks tiles of 128 envs with stride gridDim.x;
//   * one elected thread fetches the NEXT tile (every state row + the action row: contiguous 2 KB runs of the
//     structure-of-arrays layout) with cp.async.bulk into the other shared-memory stage and arms an mbarrier with
//     the expected byte count, while all threads compute the current tile out of shared memory -> the HBM read
//     latency never sits on a warp's scoreboard;
//   * the state is written back straight from registers (coalesced 16-byte stores), the observation tile goes
//     through shared memory and leaves as ONE cp.async.bulk store per tile.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// hint only: pulls [src, src + bytes) into L2.  Legal ahead of griddepcontrol.wait - L2 is the point of coherence, so a
// line the previous grid is still writing can never be observed stale through it.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// One 128-byte line per thread and trip over the rows of one tile of a structure-of-arrays state buffer (+ its action
// row): pulls the tile into L2.  Pure hint (see bulk_prefetch_l2).  cta_envs is 32, 64 or 128: a tile row is cta_envs / 8
// lines, so thread t always owns line (t mod lines-per-row) of rows t / lines-per-row, + cta_envs / lines-per-row = 8, ...
__device__ __forceinline__ void prefetch_tile_lines(const char* state_row0, long long row_bytes, int n_rows, const char* actions,
                                                    int rows, int tid, int cta_envs) {
  const int sh = (cta_envs == 128) ? 4 : (cta_envs == 64 ? 3 : 2);      // log2(lines per row)
  const int e8 = (tid & ((1 << sh) - 1)) << 3;                           // first env of this thread's line
  if (e8 >= rows) return;
  const char* src = state_row0 + (long long)(tid >> sh) * row_bytes + (e8 << 4);
  for (int r = tid >> sh; r < n_rows; r += 8) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(src) : "memory");
    src += 8 * row_bytes;
  }
  if (actions != nullptr && (tid >> sh) == (n_rows & 7)) asm volatile("prefetch.global.L2 [%0];" ::"l"(actions + (e8 << 4)) : "memory");
}
// clock reads that may race with the previous grid (speculative, ahead of the wait) or follow such a read (after the
// wait): always served by L2, never by a line L1 picked up earlier
__device__ __forceinline__ unsigned long long ld_clock_step(const Clock* c) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(&c->step) : "memory");
  return v;
}
__device__ __forceinline__ double ld_clock_svd(const Clock* c) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(&c->svd_timer) : "memory");
  return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

constexpr int TMA_MIN_CTAS = 4;   // resident CTAs per SM the fast kernel is tuned for (registers <= 128)
constexpr int TMA_MIN_CTAS_RARE = 3;   // the catch-all configuration (every feature group, everything at run time) needs ~160 registers
// (per-env models on top of a feature group: on-demand model loads + the group's code do not fit 128 registers without ~300 B of spills)
template <class T> constexpr int tma_min_ctas() {
  // The F_MISC and F_INFO kernels need 140-150 registers: at 128 they spill 40-90 B, and with four CTAs' rings in shared
  // memory L1 has next to nothing left to absorb spills.  Three CTAs per SM without spills win (measured: t2w observation +
  // noise 0.62 -> 0.70, 1D mode 0.63 -> 0.74, info buffer 0.75 -> 0.78 of the byte model).
  // (F_NOISEX: 0.60 -> 0.62 with three; F_DRAG fits 128 registers and loses with three: 0.77 -> 0.70)
  if (T::MISC || T::INFO || T::NOISEX) return TMA_MIN_CTAS_RARE;
  return (T::FEAT == F_ALL || (T::PEM && T::RARE)) ? TMA_MIN_CTAS_RARE : TMA_MIN_CTAS;
}
constexpr int WARPS = BLOCK / 32;

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// warp-level finished-episode statistics: shuffle reduce + one double atomic triple per warp (rare: episode ends)
template <typename R>
__device__ __forceinline__ void warp_stats(const Params<R>& p, double st_ret, double st_ret2, double st_n) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    st_ret += __shfl_down_sync(0xffffffffu, st_ret, off);
    st_ret2 += __shfl_down_sync(0xffffffffu, st_ret2, off);
    st_n += __shfl_down_sync(0xffffffffu, st_n, off);
  }
  if ((threadIdx.x & 31) == 0 && st_n > 0) { atomicAdd(&p.stats[0], st_ret); atomicAdd(&p.stats[1], st_ret2); atomicAdd(&p.stats[2], st_n); }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent CTAs pull tiles of 128 envs from a device-wide work counter and run them through a two-stage
// full/empty mbarrier pipeline:
//   * producer = thread 0: fetches the tile after next - every state row and the action row are contiguous 2 KB runs
//     of the structure-of-arrays layout - with cp.async.bulk into the stage the four warps have just drained into
//     registers (completion counted on the stage's `full` barrier, the drain on its `empty` barrier), so two tiles
//     (2 x 18 KB) are in flight per CTA during every compute phase and HBM latency never sits on a scoreboard;
//   * tiles are handed out dynamically (atomicAdd on a work counter): when the kernel is bandwidth-bound the SMs do
//     not progress at the same rate, and a static stride would leave the fast ones idle at the end;
//   * a launch that fits in one wave (small N) skips the async machinery and uses plain vector loads;
//   * the state goes back to HBM straight from registers (coalesced 16-byte stores); the CTA transposes its 128
//     observation rows through shared memory and ships them as one cp.async.bulk store.
// ONEWAVE: the launch has no more tiles than resident CTA slots (small N).  That case is latency-bound, so it gets its
// own instantiation with the ring, the mbarriers and every CTA barrier compiled out: direct vector loads, per-warp
// observation flush.  (tests/test_gpu_fullsize.py holds both instantiations to bit-identical results.)
template <class T, bool ONEWAVE>
__global__ void __launch_bounds__(BLOCK, tma_min_ctas<T>())
env_step_tma_kernel(const __grid_constant__ Params<float> p) {
  typedef float R;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[2], empty_bar[2];
  __shared__ long long tile_slot[2];                                     // tile held by each stage (-1: none)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // a one-wave launch may use smaller CTAs (p.ow_threads = 64): 2-warp granularity spreads the warps evenly over the SMs
  const int cta_envs = ONEWAVE ? (int)blockDim.x : BLOCK;
  const int cta_warps = cta_envs >> 5;
  const int rows_in = p.tma_rows;                                        // state rows in the buffer (7 + OU + prev action)
  const bool has_act = p.actions != nullptr && p.act_dim == 4;          // (1- and 2-wide actions are read by run_steps itself)
  const int stage_vecs = (rows_in + 1) * BLOCK;                          // float4 slots per stage, last row = actions
  const int obs_dim = T::PLAIN ? 18 : p.obs_dim;
  const int obs_tile_floats = BLOCK * obs_dim;                           // per stage
  float4* in = reinterpret_cast<float4*>(smem_raw);
  // (a one-wave launch allocates only the observation tile)
  float* obs_tiles = reinterpret_cast<float*>(smem_raw + (ONEWAVE ? (size_t)0 : (size_t)2 * stage_vecs * sizeof(float4)));
  const long long n_tiles = (p.end - p.begin + cta_envs - 1) / cta_envs;
  const long long G = gridDim.x;
  constexpr bool multi = !ONEWAVE;                                       // more tiles than resident CTAs
  const bool stream_stores = ONEWAVE && !p.host_io;                // (see the state stores below)

  __shared__ unsigned int warps_read;
  if (!multi) {
    if (tid == 0) warps_read = 0u;
    __syncthreads();
  }
#ifdef QS_PHASE_TIMING
  unsigned long long ts_a;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_a) :: "memory");
#endif
  pdl_launch_dependents();                       // the next launch may start its prologue; it waits below for our completion
  QS_CHECK(p, !p.bounds_selftest, BC_SELFTEST);
  QS_CHECK(p, n_tiles > 0 && (ONEWAVE ? (long long)gridDim.x * cta_envs >= p.end - p.begin : G <= n_tiles), BC_TILE);
#ifdef QS_DEBUG_BOUNDS
  const size_t dyn_smem = dyn_smem_bytes();
  QS_CHECK(p, ONEWAVE ? dyn_smem >= (size_t)cta_envs * obs_dim * sizeof(float)
                      : dyn_smem >= (size_t)2 * stage_vecs * sizeof(float4) + (size_t)p.obs_stages * obs_tile_floats * sizeof(float) +
                                        ((p.actions != nullptr && p.act_dim < 4) ? (size_t)2 * BLOCK * 2 * sizeof(float) : (size_t)0), BC_RING);
  QS_CHECK(p, rows_in >= 7 && rows_in <= 12 && p.row_ou < rows_in && p.row_act < rows_in && p.row_gyro < rows_in && p.row_goal < rows_in, BC_ROW);
#endif

  // ---- ahead of the wait: nothing here depends on what the previous grid is still writing ------------------------------
  // (1) L2 prefetch of this CTA's first tile: the CTA is resident well before the previous grid completes (it took the
  //     slot of a CTA that retired), and the previous grid is in its compute / store phase with DRAM idle; this grid's
  //     loads then find their lines in L2 or already on their way.  A hint only: coherent whatever the previous grid
  //     still writes.
  if (p.prefetch > 0) {
    const long long first = p.begin + (long long)blockIdx.x * cta_envs;
    long long rows = p.end - first;
    if (rows > cta_envs) rows = cta_envs;
    if (p.prefetch == 1) {                                               // one bulk prefetch per row, issued by thread 0
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)rows * 16u;
        const char* src = reinterpret_cast<const char*>(p.state + first * 4);
        for (int r = 0; r < rows_in; ++r) { bulk_prefetch_l2(src, bytes); src += p.row_bytes; }
        if (has_act) bulk_prefetch_l2(p.actions + first * 4, bytes);
      }
    } else {                                                             // one line per thread and trip (measured best)
      prefetch_tile_lines(reinterpret_cast<const char*>(p.state + first * 4), p.row_bytes, rows_in,
                          has_act ? reinterpret_cast<const char*>(p.actions + first * 4) : nullptr, (int)rows, tid, cta_envs);
    }
  }
  // (2) the draws of the step depend on (env id, step index) alone.  The step index lives in device memory (CUDA-graph
  //     replays must see it advance); the previous launch of this handle advanced it EARLY in its run (see below), so
  //     a read here almost always returns the final value: draw from it now, check it after the wait, redraw if it
  //     moved (same code, second trip of the loop - results cannot depend on which trip produced them).
  constexpr bool PREDRAW = T::SIM > 0 && T::PLAIN;
  const long long i_first = p.begin + (long long)blockIdx.x * cta_envs + tid;
  Pre<R, T> pre;
  unsigned long long step0 = 0ull;
  double svd_t0 = 0.0;
  {
    const bool spec = ONEWAVE && PREDRAW && p.spec_draw != 0;
    unsigned long long drawn_for = ~0ull;
#pragma unroll 1
    for (int trip = spec ? 0 : 1; trip < 2; ++trip) {
      unsigned long long s_now;
      if (trip == 0) {
        s_now = ld_clock_step(p.clock);
      } else {
        pdl_wait();                              // everything above touched global memory only as a hint / a guess
        s_now = ld_clock_step(p.clock);
        svd_t0 = ld_clock_svd(p.clock);
        step0 = s_now;
        if (!(ONEWAVE && PREDRAW) || s_now == drawn_for) break;
      }
      predraw<R, T>(p, i_first, i_first < p.end, s_now, pre);
      drawn_for = s_now;
    }
  }

#ifdef QS_PHASE_TIMING
  // experiment build (tools/phase_timing.py): %globaltimer stamps of one lane per warp into the info buffer
  unsigned long long ts0, ts1 = 0, ts2 = 0, ts3 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts0) :: "memory");
#endif
  double svd_end = svd_t0;

  auto issue_load = [&](long long tile, int stage) {
    const long long first = p.begin + tile * BLOCK;
    long long rows = p.end - first;
    if (rows > BLOCK) rows = BLOCK;
    const uint32_t bytes = (uint32_t)rows * 16u;
    QS_CHECK(p, tile >= 0 && tile < n_tiles && (stage == 0 || stage == 1), BC_TILE);
    QS_CHECK(p, rows > 0 && rows <= BLOCK && first >= p.begin && first + rows <= p.end && p.end <= p.n, BC_SRC);
    mbar_expect_tx(&full_bar[stage], bytes * (uint32_t)(rows_in + (has_act ? 1 : 0)));
    float4* dst = in + (size_t)stage * stage_vecs;
    const char* src = reinterpret_cast<const char*>(p.state + first * 4);
    for (int r = 0; r < rows_in; ++r) { bulk_g2s(dst + r * BLOCK, src, bytes, &full_bar[stage]); src += p.row_bytes; }
    if (has_act) bulk_g2s(dst + rows_in * BLOCK, p.actions + first * 4, bytes, &full_bar[stage]);
  };
  // next unassigned tile (the first G tiles are owned statically: tile b belongs to CTA b)
  auto fetch_tile = [&]() -> long long {
    const long long t = (long long)atomicAdd(p.tile_counter, 1u) + G;
    return t < n_tiles ? t : -1;
  };

  if (multi) {
    if (tid == 0) {
      mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
      mbar_init(&empty_bar[0], WARPS); mbar_init(&empty_bar[1], WARPS);
      fence_mbar_init();
      // tile j of this CTA lives in stage j & 1.  Tile 0 is copied into stage 0 by the threads themselves (below);
      // the producer prefetches tile 1 now and refills a stage whenever the four warps have drained it.
      const long long t1 = fetch_tile();
      tile_slot[1] = t1;
      if (t1 >= 0) issue_load(t1, 1);
    }
    __syncthreads();
  }

  // 1- and 2-wide actions (VerticalControl, VertPlaneControl) are not part of the bulk-copy ring: each thread fetches its
  // own with a 4-byte cp.async into a two-slot shared-memory stage, one tile AHEAD (the index of the CTA's next tile is known a
  // whole tile early), so the load latency hides behind the compute phase of the current tile and no register is held for it
  // (a register prefetch was spilled by the compiler straight after the load - i.e. it waited for the load on the spot).
  const bool act_direct = multi && !has_act && p.actions != nullptr && p.act_dim < 4;
  float* act_stage = obs_tiles + (size_t)(multi ? p.obs_stages : 0) * obs_tile_floats;      // [2][BLOCK][2], requested by the host
  auto fetch_narrow_actions = [&](long long t, int slot) {
    const long long ia = p.begin + t * BLOCK + tid;
    QS_CHECK(p, t >= -1 && t < n_tiles && (slot == 0 || slot == 1), BC_ACTION);
    if (t >= 0 && ia < p.end) {
      float* dst = act_stage + ((size_t)slot * BLOCK + tid) * 2;
      cp_async4(dst, p.actions + ia * p.act_dim);
      if (p.act_dim == 2) cp_async4(dst + 1, p.actions + ia * 2 + 1);
    }
  };
  uint32_t full_phase = 0;        // bit s: parity to wait for on full_bar[s]
  uint32_t empty_phase = 0;       // (producer only) bit s: parity to wait for on empty_bar[s]
  int j = 0;
  long long tile = blockIdx.x;
  while (tile >= 0) {
    const int stage = j & 1;
    const long long block_first = p.begin + tile * cta_envs;
    const long long i = block_first + tid;
    const bool active = i < p.end;

    // ---- tile data.  First tile of the CTA: every thread copies its own nine 16-byte vectors into stage 0 with
    // cp.async (no register scoreboard is held, so the draws below are generated while the copies are in flight);
    // later tiles were put into the ring by the producer's bulk copies.
    const bool ou_row = T::OU >= 0 ? (T::OU != 0) : (p.ou_on != 0);
    const bool act_row = !T::PLAIN && p.row_act >= 0;
    float4* buf = in + (size_t)stage * stage_vecs + tid;
    St<R> s;
    R a0[4] = {0.f, 0.f, 0.f, 0.f};
    if (ONEWAVE) {
      load_state<R, T>(s, p, i, active);
      if (active && has_act) { float4 q = *reinterpret_cast<const float4*>(p.actions + i * 4); a0[0] = q.x; a0[1] = q.y; a0[2] = q.z; a0[3] = q.w; }
    } else if (j == 0) {
      const long long il = active ? i : p.end - 1;                       // lanes beyond the last env: any valid env
      QS_CHECK(p, il >= p.begin && il < p.end && p.end <= p.n, BC_SRC);
      const char* ptr = reinterpret_cast<const char*>(p.state + il * 4);
      for (int r = 0; r < rows_in; ++r) cp_async16(buf + r * BLOCK, ptr + r * p.row_bytes);
      if (has_act) cp_async16(buf + rows_in * BLOCK, p.actions + il * 4);
      if (act_direct) fetch_narrow_actions(tile, 0);                       // (same group as the rows of the first tile)
      cp_async_commit();
    }
    if (PREDRAW && !ONEWAVE) predraw<R, T>(p, i, active, step0, pre);
    unsigned int ticket = 0u;
    if (ONEWAVE) {
      // Early clock advance: every warp has read the clock by now (its draws depended on the value); the last CTA of
      // the grid to say so advances it - long before the grid ends, so the next launch's speculative read sees the final
      // value, and the end of this grid is its last store rather than an atomic round trip followed by a dependent
      // store.  The ticket comes back while the state loads are in flight; it is looked at after the step is computed.
      // (the increment is made to depend on the two values read, so the reads have completed before it is issued)
      QS_CHECK(p, !active || (i >= p.begin && i < p.end), BC_ENV);
      const unsigned int one = 1u + (unsigned int)(step0 >> 63) + (unsigned int)((unsigned long long)__double_as_longlong(svd_t0) >> 63);
      __syncwarp();
      if (lane == 0 && atomicAdd(&warps_read, one) == (unsigned)cta_warps - 1u) ticket = atomicAdd(p.launch_ticket, 1u) + 1u;
    }
    if (multi) {
      if (j == 0) {
        cp_async_wait_all();                                             // a thread reads back only what it copied itself
      } else {
        mbar_wait(&full_bar[stage], (full_phase >> stage) & 1u);
        full_phase ^= (1u << stage);
      }
      if (act_direct) {
        if (j > 0) cp_async_wait_all();                                  // this tile's actions: issued a whole tile ago
        if (active) { const float* src = act_stage + ((size_t)stage * BLOCK + tid) * 2; a0[0] = src[0]; if (p.act_dim == 2) a0[1] = src[1]; }
        fetch_narrow_actions(tile_slot[(j + 1) & 1], (j + 1) & 1);         // (slot of the next tile: written at least one barrier ago)
        cp_async_commit();
      }
      // lanes beyond the last env of a ragged tile read stale shared memory: harmless, nothing of theirs is stored
      unpack_core(s, buf[0 * BLOCK], buf[1 * BLOCK], buf[2 * BLOCK], buf[3 * BLOCK], buf[4 * BLOCK], buf[5 * BLOCK], buf[6 * BLOCK]);
      init_optional<R, T::SWARM>(s, p);
      if (ou_row) { float4 q = buf[p.row_ou * BLOCK]; s.ou[0] = q.x; s.ou[1] = q.y; s.ou[2] = q.z; s.ou[3] = q.w; }
      if (act_row) { float4 q = buf[p.row_act * BLOCK]; s.pa[0] = q.x; s.pa[1] = q.y; s.pa[2] = q.z; s.pa[3] = q.w; }
      if (has_act) { float4 q = buf[rows_in * BLOCK]; a0[0] = q.x; a0[1] = q.y; a0[2] = q.z; a0[3] = q.w; }
      if (T::NOISEX) { if (p.row_gyro >= 0) { float4 q = buf[p.row_gyro * BLOCK]; s.gb[0] = q.x; s.gb[1] = q.y; s.gb[2] = q.z; } }
      if (T::MISC) { if (p.row_goal >= 0) { float4 q = buf[p.row_goal * BLOCK]; s.gx = q.x; s.gy = q.y; s.gz = q.z; } }
      finish_load(s);

      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);                     // this warp is done with the stage
      if (tid == 0) {
        // refill the stage with the tile after next as soon as all four warps have drained it
        mbar_wait(&empty_bar[stage], (empty_phase >> stage) & 1u);
        empty_phase ^= (1u << stage);
        const long long t = fetch_tile();
        tile_slot[stage] = t;                                            // read by everybody two tiles from now
        if (t >= 0) issue_load(t, stage);
      }
    }

#ifdef QS_PHASE_TIMING
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts1) : "f"(s.px + s.r[8] + a0[0] + s.cd[3]) : "memory");   // data arrived
#endif
    R o[OBS_SCRATCH];
    StepOut<R> out;
    run_steps<R, T>(s, p, i, active, (has_act || act_direct) ? a0 : nullptr, svd_t0, step0, o, out, &pre);
    svd_end = out.svd_t;
#ifdef QS_PHASE_TIMING
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts2) : "f"(o[0] + o[17] + out.rew_sum + s.vz) : "memory");   // computed
#endif

    if (!multi) {
      // (the clock was read by every warp of every CTA before the last ticket was taken: nobody can see the update)
      QS_CHECK(p, ticket <= (unsigned)gridDim.x, BC_TICKET);
      if (ticket == (unsigned)gridDim.x) {
        *p.launch_ticket = 0u;
        if (p.clock_advance) { p.clock->svd_timer = out.svd_t; p.clock->step = step0 + (unsigned long long)p.k_steps; }
      }
    }
    if (active) {
      // One-wave launches into DEVICE buffers store with the streaming (evict-first) hint: in the ring regime the lines written
      // here are evicted long before anybody re-reads them, and marking them so lets L2 drop them first (5.81 -> 5.78 us per
      // 65,536-env step; the L2-resident single-batch case 5.52 -> 5.49).  Not for the pipelined instantiation (+2 % at the
      // HBM-resident sizes) and not for pinned host buffers (posted writes across PCIe were 3 % slower with it).
      if (ONEWAVE && stream_stores) store_state<R, T, ONEWAVE>(s, p, i);
      else store_state<R, T, false>(s, p, i);
      p.rew[i] = out.rew_sum;
      p.done[i] = out.done_any ? 1 : 0;
    }
    if (!multi) {
      // ---- one-wave launch (small N): latency-bound, so no CTA barrier anywhere - every warp transposes and ships
      // its own 32 observation rows.  (Same run_steps instance as the pipelined path: results do not depend on which
      // path a batch size selects.)
      float* wtile = obs_tiles + (size_t)warp * 32 * obs_dim;
#ifdef QS_DEBUG_BOUNDS
      QS_CHECK(p, (size_t)(warp + 1) * 32 * obs_dim * sizeof(float) <= dyn_smem, BC_OBS_SMEM);
#endif
      obs_to_tile<R, T::PLAIN, T::MISC>(p, o, wtile + lane * obs_dim);
      __syncwarp();
      const long long warp_first = i - lane;
      long long wrows = p.end - warp_first;
      if (wrows > 32) wrows = 32;
      QS_CHECK(p, wrows <= 0 || (warp_first >= p.begin && warp_first + wrows <= p.end), BC_OBS_GLOBAL);
      if (wrows > 0) {
        const int total = (int)wrows * obs_dim;
        float* dst = p.obs + warp_first * obs_dim;             // 128 * obs_dim bytes per full warp: 16-byte aligned
        if (T::PLAIN && wrows == 32) {                         // 32 x 18 words = 144 vectors: four full trips + half a trip
          const float4* t4 = reinterpret_cast<const float4*>(wtile) + lane;
          float4* d4 = reinterpret_cast<float4*>(dst) + lane;
          const float4 v0 = t4[0], v1 = t4[32], v2 = t4[64], v3 = t4[96];
          if (stream_stores) {
            __stcs(d4, v0); __stcs(d4 + 32, v1); __stcs(d4 + 64, v2); __stcs(d4 + 96, v3);
            if (lane < 16) __stcs(d4 + 128, t4[128]);
          } else {
            d4[0] = v0; d4[32] = v1; d4[64] = v2; d4[96] = v3;
            if (lane < 16) d4[128] = t4[128];
          }
        } else if ((total & 3) == 0) {
          const float4* t4 = reinterpret_cast<const float4*>(wtile);
          float4* d4 = reinterpret_cast<float4*>(dst);
          for (int k = lane; k < (total >> 2); k += 32) d4[k] = t4[k];
        } else {
          for (int k = lane; k < total; k += 32) dst[k] = wtile[k];
        }
      }
#ifdef QS_PHASE_TIMING
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts3) :: "memory");                                            // stores issued
      if (lane == 0 && p.timing != nullptr) {
        unsigned int smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        // two record sets (step parity), so the last two launches of a graph can be laid side by side
        unsigned long long* rec = p.timing + ((long long)(((p.env_id_offset / p.n) + (long long)step0) & 1) * gridDim.x * cta_warps + (long long)blockIdx.x * cta_warps + warp) * 6;
        rec[0] = ts_a; rec[1] = ts0; rec[2] = ts1; rec[3] = ts2; rec[4] = ts3; rec[5] = smid;
      }
#endif
      if (__any_sync(0xffffffffu, out.done_any)) warp_stats(p, out.st_ret, out.st_ret2, out.st_n);
      return;
    }
    // observation rows: registers -> shared tile -> one bulk store per CTA
    // (one observation tile instead of two when the host found that this buys a fourth resident CTA: its previous bulk
    // store - issued a whole load + compute phase ago - must then have been read out before the tile is overwritten)
    const bool one_obs_tile = p.obs_stages == 1;
    float* tile_obs = obs_tiles + (size_t)(one_obs_tile ? 0 : stage) * obs_tile_floats;
    if (one_obs_tile) {
      if (tid == 0) bulk_wait_read0();
      __syncthreads();
    }
#ifdef QS_DEBUG_BOUNDS
    QS_CHECK(p, (size_t)((const unsigned char*)(tile_obs + obs_tile_floats) - smem_raw) <= dyn_smem && (one_obs_tile || p.obs_stages == 2), BC_OBS_SMEM);
#endif
    obs_to_tile<R, T::PLAIN, T::MISC>(p, o, tile_obs + tid * obs_dim);
    if (multi) {
      fence_proxy_async();                       // generic-proxy tile writes -> async proxy
      if (tid == 0) bulk_wait_read0();           // last tile's store has drained the OTHER stage (it had a whole tile)
    }
    __syncthreads();                             // tile complete; also keeps the warps' 512-byte row segments adjacent in time
    long long rows = p.end - block_first;
    if (rows > BLOCK) rows = BLOCK;
    const uint32_t obs_bytes = (uint32_t)rows * (uint32_t)obs_dim * 4u;
    QS_CHECK(p, rows > 0 && rows <= BLOCK && block_first >= p.begin && block_first + rows <= p.end, BC_OBS_GLOBAL);
    if (multi && (obs_bytes & 15u) == 0u) {
      if (tid == 0) { bulk_s2g(p.obs + block_first * obs_dim, tile_obs, obs_bytes); bulk_commit(); }
    } else {
      flush_tile(p, tile_obs, block_first, nullptr);                     // one-wave launch, or ragged last tile
    }
    if (__any_sync(0xffffffffu, out.done_any)) warp_stats(p, out.st_ret, out.st_ret2, out.st_n);
    ++j;
    tile = multi ? tile_slot[j & 1] : -1;        // written at least one barrier ago
    QS_CHECK(p, tile == -1 || (tile >= G && tile < n_tiles), BC_TILE);
  }
  if (tid == 0) bulk_wait_read0();               // shared memory must outlive the bulk stores
  __syncthreads();                               // every thread of the CTA has read the clock
  finish_launch(p, svd_end, (unsigned long long)p.k_steps);
}

// ---------------------------------------------------------------------------------------------------------
// masked reset (QuadrotorEnv.reset, quadrotor.py:1149-1150)
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__global__ void __launch_bounds__(BLOCK) env_reset_kernel(const __grid_constant__ Params<R> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  R* tile = reinterpret_cast<R*>(smem_raw);
  __shared__ uint8_t valid[BLOCK];
  const long long block_first = p.begin + (long long)blockIdx.x * BLOCK;
  const long long i = block_first + threadIdx.x;
  const bool active = i < p.end;
  const bool mine = active && (p.reset_mask == nullptr || p.reset_mask[i] != 0);
  valid[threadIdx.x] = mine ? 1 : 0;
  if (mine) {
    St<R> s;
    load_state<R, T>(s, p, i, true);
    NoiseSrc<R, T> ns(p, i);
    ns.begin_step(p.clock->step);
    reset_env<R, T>(s, p, ns);
    R o[OBS_SCRATCH];
    const Mdl<R, T::PEM> m(p, i);
    build_obs<R, T>(s, p, m, ns, true, s.pa, o);
    store_state<R, T>(s, p, i);
    obs_to_tile(p, o, tile + threadIdx.x * p.obs_dim);
  }
  __syncthreads();
  flush_tile(p, tile, block_first, p.reset_mask ? valid : nullptr);
  // a reset consumes one Philox step index so that its draws are never reused by a step
  finish_clock(p, p.clock->svd_timer, 1ull);
}

// ---------------------------------------------------------------------------------------------------------
// state_<obs_repr>(env) / SensorNoise.add_noise on the bound state (get_state.py, sensor_noise.py:100-159): a fresh
// noisy observation without stepping.  One add_noise call: draws come from the reset-observation slots.
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__global__ void __launch_bounds__(BLOCK) env_observe_kernel(const __grid_constant__ Params<R> p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  R* tile = reinterpret_cast<R*>(smem_raw);
  const long long block_first = p.begin + (long long)blockIdx.x * BLOCK;
  const long long i = block_first + threadIdx.x;
  const bool active = i < p.end;
  if (active) {
    St<R> s;
    load_state<R, T>(s, p, i, true);
    NoiseSrc<R, T> ns(p, i);
    ns.begin_step(p.clock->step);
    if (p.acc_in != nullptr) {        // world acceleration that reproduces the given reading: R a - (0, 0, gravity)
      const R ax = p.acc_in[i * 3], ay = p.acc_in[i * 3 + 1], az = p.acc_in[i * 3 + 2];
      s.ax = s.r[0] * ax + s.r[1] * ay + s.r[2] * az;
      s.ay = s.r[3] * ax + s.r[4] * ay + s.r[5] * az;
      s.az = s.r[6] * ax + s.r[7] * ay + s.r[8] * az - p.gravity;
    } else {
      s.fresh = true;                 // (0, 0, GRAV), the reading set_state leaves behind (quadrotor.py:221)
    }
    R o[OBS_SCRATCH];
    const Mdl<R, T::PEM> m(p, i);
    build_obs<R, T>(s, p, m, ns, true, s.pa, o);
    if (T::NOISEX) { if (p.row_gyro >= 0) st4(p.state, p.n, p.row_gyro, i, s.gb[0], s.gb[1], s.gb[2], R(0)); }   // the bias walks on
    obs_to_tile(p, o, tile + threadIdx.x * p.obs_dim);
  }
  __syncthreads();
  flush_tile(p, tile, block_first, nullptr);
  finish_clock(p, p.clock->svd_timer, 1ull);
}

// ---------------------------------------------------------------------------------------------------------
// controller.step_func / QuadrotorDynamics.step on the bound state, no env bookkeeping
// ---------------------------------------------------------------------------------------------------------
template <typename R, class T>
__global__ void __launch_bounds__(BLOCK) dynamics_step_kernel(const __grid_constant__ Params<R> p) {
  const long long i = p.begin + (long long)blockIdx.x * BLOCK + threadIdx.x;
  const bool active = i < p.end;
  St<R> s;
  load_state<R, T>(s, p, i, active);
  NoiseSrc<R, T> ns(p, active ? i : 0);
  ns.begin_step(p.clock->step);
  R a[4] = {R(0), R(0), R(0), R(0)};
  int adim = p.act_dim;
  if (active && p.actions != nullptr) {
#pragma unroll
    for (int k = 0; k < 4; ++k) if (k < adim) a[k] = p.actions[i * adim + k];
  }
  R cmd[4];
  const Mdl<R, T::PEM> m(p, active ? i : p.begin);
  if (p.dyn_controller < 0) { cmd[0] = a[0]; cmd[1] = a[1]; cmd[2] = a[2]; cmd[3] = a[3]; }
  else controller_cmds<R, T>(s, p, m, p.dyn_controller, a, cmd);
  double svd_t = p.clock->svd_timer;
#pragma unroll 1
  for (int sub = 0; sub < p.dyn_substeps; ++sub) {
    const bool fire = svd_tick(svd_t, p.dt_d, p.svd_period);
    R z[4] = {R(0), R(0), R(0), R(0)};
    if (p.ou_on) ns.ou(sub, z);
    substep<R, T>(s, cmd, p, m, z, fire);
  }
  if (active) {
    store_state<R, T>(s, p, i);
    if (p.info != nullptr) {
      R* inf = p.info + i * INFO_DIM;
      R am[3];
      accelerometer(s, p, am);
      inf[0] = s.ax; inf[1] = s.ay; inf[2] = s.az; inf[3] = am[0]; inf[4] = am[1]; inf[5] = am[2];
      inf[6] = s.tqx; inf[7] = s.tqy; inf[8] = s.tqz; inf[9] = s.odx; inf[10] = s.ody; inf[11] = s.odz;
      inf[12] = cmd[0]; inf[13] = cmd[1]; inf[14] = cmd[2]; inf[15] = cmd[3];
    }
  }
  __syncthreads();
  finish_clock(p, svd_t, 1ull);
}

// ---------------------------------------------------------------------------------------------------------
// QuadrotorDynamics.F(s, dt) / G(s) (quadrotor.py:461-518): the affine representation s' = dt (F(s) + G(s) u) for n
// caller-supplied states s[n][21] = xyz, Vxyz, R (row-major), omega, goal.  One thread per state; the 21-word rows go
// through shared memory so that every global access is coalesced.  G is 21 x C with C = 1 / 2 / 4 controls (control_mx,
// :121-129); only its velocity and body-rate rows are non-zero.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct AffineArgs {
  const R* s;
  R* F;
  R* G;
  long long n;
  R dt;
  int C;
};

template <typename R, bool PEM>
__global__ void __launch_bounds__(BLOCK) affine_fg_kernel(const __grid_constant__ Params<R> p, const AffineArgs<R> a) {
  __shared__ R tile[BLOCK * 21];          // states in, then F out
  __shared__ R gnz[BLOCK * 24];           // the six non-zero rows of G (x up to 4 controls) per state
  const long long first = (long long)blockIdx.x * BLOCK;
  long long rows = a.n - first;
  if (rows > BLOCK) rows = BLOCK;
  const int tid = threadIdx.x;
  const int total = (int)rows * 21;
  for (int k = tid; k < total; k += BLOCK) tile[k] = a.s[first * 21 + k];
  __syncthreads();
  const Mdl<R, PEM> m(p, 0);
  const int C = a.C;
  R f[21];
  if (tid < rows) {
    const R* x = tile + tid * 21;
    const R vx = x[3], vy = x[4], vz = x[5];
    R r[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = x[6 + k];
    const R ox = x[15], oy = x[16], oz = x[17];
    f[0] = vx; f[1] = vy; f[2] = vz;                                                       // :471
    const R vd = R(1) - m.vel_damp();
    f[3] = (vd * vx - vx) / a.dt + R(0); f[4] = (vd * vy - vy) / a.dt + R(0); f[5] = (vd * vz - vz) / a.dt + R(-9.81);   // :475
    const R wx = r[0] * ox + r[1] * oy + r[2] * oz, wy = r[3] * ox + r[4] * oy + r[5] * oz;  // :479 body -> world
    const R wz = r[6] * ox + r[7] * oy + r[8] * oz;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                                          // :481-484 dR = [w]x R
      f[6 + c] = -wz * r[3 + c] + wy * r[6 + c];
      f[9 + c] = wz * r[c] - wx * r[6 + c];
      f[12 + c] = -wy * r[c] + wx * r[3 + c];
    }
    const R iwx = m.inertia(0) * ox, iwy = m.inertia(1) * oy, iwz = m.inertia(2) * oz;
    const R gx = (-oy) * iwz - (-oz) * iwy, gy = (-oz) * iwx - (-ox) * iwz, gz = (-ox) * iwy - (-oy) * iwx;   // :488
    const R dqx = R(1) - clampr(m.damp_omega_q() * (ox * ox), R(0), R(1));                  // :489
    const R dqy = R(1) - clampr(m.damp_omega_q() * (oy * oy), R(0), R(1));
    const R dqz = R(1) - clampr(m.damp_omega_q() * (oz * oz), R(0), R(1));
    f[15] = dqx * (m.inv_inertia(0) * gx); f[16] = dqy * (m.inv_inertia(1) * gy); f[17] = dqz * (m.inv_inertia(2) * gz);   // :490
    f[18] = f[19] = f[20] = R(0);
    // G (:501-518): base matrix 21 x 4, then @ control_mx
    R g[6][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const R tm = m.thrust_max(k);
      g[0][k] = (r[2] / m.mass()) * tm; g[1][k] = (r[5] / m.mass()) * tm; g[2][k] = (r[8] / m.mass()) * tm;   // :509
      g[3][k] = dqx * (m.inv_inertia(0) * (tm * m.pcx(k)));                                 // :188-191, :516
      g[4][k] = dqy * (m.inv_inertia(1) * (tm * m.pcy(k)));
      g[5][k] = dqz * (m.inv_inertia(2) * m.tz(k));
    }
    R* go = gnz + tid * 24;
#pragma unroll
    for (int row = 0; row < 6; ++row) {
      if (C == 4) { go[row * 4] = g[row][0]; go[row * 4 + 1] = g[row][1]; go[row * 4 + 2] = g[row][2]; go[row * 4 + 3] = g[row][3]; }
      else if (C == 2) { go[row * 2] = g[row][0] + g[row][1]; go[row * 2 + 1] = g[row][2] + g[row][3]; }
      else go[row] = ((g[row][0] + g[row][1]) + g[row][2]) + g[row][3];
    }
  }
  __syncthreads();                         // every state of the tile has been read
  if (tid < rows) {
#pragma unroll
    for (int k = 0; k < 21; ++k) tile[tid * 21 + k] = f[k];
  }
  __syncthreads();
  if (a.F != nullptr)
    for (int k = tid; k < total; k += BLOCK) a.F[first * 21 + k] = tile[k];
  if (a.G != nullptr) {
    const int per = 21 * C;
    for (int k = tid; k < (int)rows * per; k += BLOCK) {
      const int st = k / per, rem = k - st * per, row = rem / C, c = rem - row * C;
      R v = R(0);
      if (row >= 3 && row < 6) v = gnz[st * 24 + (row - 3) * C + c];
      else if (row >= 15 && row < 18) v = gnz[st * 24 + (row - 12) * C + c];
      a.G[first * per + k] = v;
    }
  }
}

// advances the clock by one env step on behalf of a set of read-only (clock_advance = 0) chunk launches
template <typename R>
__global__ void clock_advance_kernel(const __grid_constant__ Params<R> p) {
  double t = p.clock->svd_timer;
  for (int k = 0; k < p.k_steps * p.sim_steps; ++k) svd_tick(t, p.dt_d, p.svd_period);
  p.clock->svd_timer = t;
  p.clock->step += (unsigned long long)p.k_steps;
}

// ---------------------------------------------------------------------------------------------------------
// launch interface implemented once per real type (qs_kernels_f32.cu / qs_kernels_f64.cu)
// ---------------------------------------------------------------------------------------------------------
template <typename R> cudaError_t launch_step(const Params<R>& p, cudaStream_t st, const char** variant_name);
template <typename R> cudaError_t launch_reset(const Params<R>& p, cudaStream_t st);
template <typename R> cudaError_t launch_dynamics(const Params<R>& p, cudaStream_t st);
template <typename R> cudaError_t launch_observe(const Params<R>& p, cudaStream_t st);
template <typename R> cudaError_t launch_clock_advance(const Params<R>& p, cudaStream_t st);
template <typename R> cudaError_t launch_affine(const Params<R>& p, const AffineArgs<R>& a, cudaStream_t st);
// fp32 fast subset (qs_fast.inl, one translation unit per controller x noise source)
#define QS_DECL_FAST(name) cudaError_t name(const Params<float>& p, cudaStream_t st, const char** variant_name)
QS_DECL_FAST(launch_fast_raw_philox);
QS_DECL_FAST(launch_fast_raw_inject);
QS_DECL_FAST(launch_fast_omega_philox);
QS_DECL_FAST(launch_fast_omega_inject);
QS_DECL_FAST(launch_fast_mellinger_philox);
QS_DECL_FAST(launch_fast_mellinger_inject);
QS_DECL_FAST(launch_fast_shifted_philox);
QS_DECL_FAST(launch_fast_shifted_inject);
QS_DECL_FAST(launch_fast_velyaw_philox);
QS_DECL_FAST(launch_fast_velyaw_inject);
QS_DECL_FAST(launch_tma_generic_all_philox);    // every other fp32 configuration (qs_tma_generic.inl): catch-all ...
QS_DECL_FAST(launch_tma_generic_all_inject);
QS_DECL_FAST(launch_tma_generic_drag_philox);   // ... and one kernel per feature group (Feat) at sim_steps = 2
QS_DECL_FAST(launch_tma_generic_drag_inject);
QS_DECL_FAST(launch_tma_generic_noise_philox);
QS_DECL_FAST(launch_tma_generic_noise_inject);
QS_DECL_FAST(launch_tma_generic_misc_philox);
QS_DECL_FAST(launch_tma_generic_misc_inject);
QS_DECL_FAST(launch_tma_generic_info_philox);
QS_DECL_FAST(launch_tma_generic_info_inject);
QS_DECL_FAST(launch_tma_generic_dragnoise_philox);   // ... and the one two-group pairing with a kernel of its own: F_DRAG | F_NOISEX
QS_DECL_FAST(launch_tma_generic_dragnoise_inject);
#undef QS_DECL_FAST

}  // namespace qs
