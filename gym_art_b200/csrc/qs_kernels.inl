// qs_kernels.inl - generic kernel instantiation + variant dispatch for one real type (QS_REAL), included by
// qs_kernels_f32.cu / qs_kernels_f64.cu so that the two precisions compile in parallel.
#include "qs_device.cuh"

namespace qs {

template <typename R, class T>
static cudaError_t launch_step_t(const Params<R>& p, cudaStream_t st) {
  const unsigned grid = (unsigned)((p.end - p.begin + BLOCK - 1) / BLOCK);
  const size_t smem = (size_t)BLOCK * p.obs_dim * sizeof(R);
  env_step_kernel<R, T><<<grid, BLOCK, smem, st>>>(p);
  return cudaGetLastError();
}

typedef Traits<-1, -1, -1, true, F_ALL> TGenInject;
typedef Traits<-1, -1, -1, false, F_ALL> TGenPhilox;
typedef Traits<-1, -1, -1, true, F_ALL, -1, false, false, true> TGenInjectPEM;     // per-env model constants
typedef Traits<-1, -1, -1, false, F_ALL, -1, false, false, true> TGenPhiloxPEM;

template <>
cudaError_t launch_step<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st, const char** name) {
  typedef QS_REAL R;
#ifdef QS_FAST_VARIANTS
  // F_MISC with the RawControl / Mellinger controllers at the default sim_steps: compile-time controller twins (qs_fast.inl)
  // (a twin of F_DRAG was measured too: 0.77 -> 0.79 for 16 more kernels per translation unit - not kept)
  if ((p.feat == F_MISC || (p.feat == F_NOISEX && p.models == nullptr)) && p.act_dim == 4 && p.sim_steps == 2 && p.dim_mode == 3 &&
      (p.controller == CTRL_RAW || p.controller == CTRL_MELLINGER)) {
    if (p.controller == CTRL_RAW) return p.inject ? launch_fast_raw_inject(p, st, name) : launch_fast_raw_philox(p, st, name);
    return p.inject ? launch_fast_mellinger_inject(p, st, name) : launch_fast_mellinger_philox(p, st, name);
  }
  // per-env models whose drag / linearity / damping rows are in use, in the common step shape: the specialised kernels with
  // F_DRAG compiled in (qs_fast.inl)
  const bool pem_drag_plain = p.feat == F_DRAG && p.models != nullptr && p.sim_steps == 2 && p.k_steps == 1 && p.obs_dim == 18 &&
                              !p.obs_rel && p.row_act < 0;
  if ((p.feat == F_NONE || pem_drag_plain) && p.act_dim == 4 &&
      (p.controller == CTRL_RAW || p.controller == CTRL_OMEGA_THRUST || p.controller == CTRL_MELLINGER ||
       p.controller == CTRL_SHIFTED || p.controller == CTRL_VELOCITY_YAW)) {
    if (p.controller == CTRL_RAW) return p.inject ? launch_fast_raw_inject(p, st, name) : launch_fast_raw_philox(p, st, name);
    if (p.controller == CTRL_OMEGA_THRUST) return p.inject ? launch_fast_omega_inject(p, st, name) : launch_fast_omega_philox(p, st, name);
    if (p.controller == CTRL_SHIFTED) return p.inject ? launch_fast_shifted_inject(p, st, name) : launch_fast_shifted_philox(p, st, name);
    if (p.controller == CTRL_VELOCITY_YAW) return p.inject ? launch_fast_velyaw_inject(p, st, name) : launch_fast_velyaw_philox(p, st, name);
    return p.inject ? launch_fast_mellinger_inject(p, st, name) : launch_fast_mellinger_philox(p, st, name);
  }
#endif
#ifdef QS_FAST_VARIANTS
  // fp32: everything outside the specialised subset still runs through the persistent TMA harness - a kernel that carries
  // just the feature group the configuration needs when it is a single group at the default sim_steps, the catch-all otherwise
  if (p.tma_generic) {
    if (p.models != nullptr && p.sim_steps == 2) {       // (swarms / the log-distance reward with per-env models carry F_MISC)
      if (p.feat == F_DRAG) return p.inject ? launch_tma_generic_drag_inject(p, st, name) : launch_tma_generic_drag_philox(p, st, name);
      if (p.feat == F_MISC) return p.inject ? launch_tma_generic_misc_inject(p, st, name) : launch_tma_generic_misc_philox(p, st, name);
    }
    if (p.models == nullptr && p.sim_steps == 2) {
      if (p.feat == F_DRAG) return p.inject ? launch_tma_generic_drag_inject(p, st, name) : launch_tma_generic_drag_philox(p, st, name);
      if (p.feat == F_NOISEX) return p.inject ? launch_tma_generic_noise_inject(p, st, name) : launch_tma_generic_noise_philox(p, st, name);
      if (p.feat == F_INFO) return p.inject ? launch_tma_generic_info_inject(p, st, name) : launch_tma_generic_info_philox(p, st, name);
      // drag model + full sensor noise (the sim-to-real pairing): 0.49 of its byte model through the catch-all
      if (p.feat == (F_DRAG | F_NOISEX)) return p.inject ? launch_tma_generic_dragnoise_inject(p, st, name) : launch_tma_generic_dragnoise_philox(p, st, name);
      if (p.feat == F_MISC || p.feat == F_NONE) return p.inject ? launch_tma_generic_misc_inject(p, st, name) : launch_tma_generic_misc_philox(p, st, name);
    }
    return p.inject ? launch_tma_generic_all_inject(p, st, name) : launch_tma_generic_all_philox(p, st, name);
  }
#endif
  if (p.models != nullptr) {
    if (p.inject) { if (name) *name = "generic-permodel-injected"; return launch_step_t<R, TGenInjectPEM>(p, st); }
    if (name) *name = "generic-permodel-philox";
    return launch_step_t<R, TGenPhiloxPEM>(p, st);
  }
  if (p.inject) { if (name) *name = "generic-injected"; return launch_step_t<R, TGenInject>(p, st); }
  if (name) *name = "generic-philox";
  return launch_step_t<R, TGenPhilox>(p, st);
}

template <>
cudaError_t launch_reset<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st) {
  typedef QS_REAL R;
  const unsigned grid = (unsigned)((p.end - p.begin + BLOCK - 1) / BLOCK);
  const size_t smem = (size_t)BLOCK * p.obs_dim * sizeof(R);
  if (p.inject) env_reset_kernel<R, TGenInject><<<grid, BLOCK, smem, st>>>(p);
  else env_reset_kernel<R, TGenPhilox><<<grid, BLOCK, smem, st>>>(p);
  return cudaGetLastError();
}

template <>
cudaError_t launch_dynamics<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st) {
  typedef QS_REAL R;
  const unsigned grid = (unsigned)((p.end - p.begin + BLOCK - 1) / BLOCK);
  if (p.models != nullptr) {
    if (p.inject) dynamics_step_kernel<R, TGenInjectPEM><<<grid, BLOCK, 0, st>>>(p);
    else dynamics_step_kernel<R, TGenPhiloxPEM><<<grid, BLOCK, 0, st>>>(p);
  } else if (p.inject) dynamics_step_kernel<R, TGenInject><<<grid, BLOCK, 0, st>>>(p);
  else dynamics_step_kernel<R, TGenPhilox><<<grid, BLOCK, 0, st>>>(p);
  return cudaGetLastError();
}

template <>
cudaError_t launch_observe<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st) {
  typedef QS_REAL R;
  const unsigned grid = (unsigned)((p.end - p.begin + BLOCK - 1) / BLOCK);
  const size_t smem = (size_t)BLOCK * p.obs_dim * sizeof(R);
  if (p.inject) env_observe_kernel<R, TGenInject><<<grid, BLOCK, smem, st>>>(p);
  else env_observe_kernel<R, TGenPhilox><<<grid, BLOCK, smem, st>>>(p);
  return cudaGetLastError();
}

template <>
cudaError_t launch_affine<QS_REAL>(const Params<QS_REAL>& p, const AffineArgs<QS_REAL>& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.n + BLOCK - 1) / BLOCK);
  if (p.models != nullptr) affine_fg_kernel<QS_REAL, true><<<grid, BLOCK, 0, st>>>(p, a);
  else affine_fg_kernel<QS_REAL, false><<<grid, BLOCK, 0, st>>>(p, a);
  return cudaGetLastError();
}

template <>
cudaError_t launch_clock_advance<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st) {
  clock_advance_kernel<QS_REAL><<<1, 1, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace qs
