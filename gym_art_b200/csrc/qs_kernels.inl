// qs_kernels.inl - kernel instantiation + variant dispatch for one real type (QS_REAL), included by
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

typedef Traits<-1, -1, -1, true, true> TGenInject;
typedef Traits<-1, -1, -1, false, true> TGenPhilox;

#ifdef QS_FAST_VARIANTS
// default-feature subset, compile-time controller / OU / sensor-noise switches (fp32 throughput path)
template <typename R, int C>
static cudaError_t launch_fast_c(const Params<R>& p, cudaStream_t st) {
  if (p.ou_on) {
    if (p.sense_on) return launch_step_t<R, Traits<C, 1, 1, false, false> >(p, st);
    return launch_step_t<R, Traits<C, 1, 0, false, false> >(p, st);
  }
  if (p.sense_on) return launch_step_t<R, Traits<C, 0, 1, false, false> >(p, st);
  return launch_step_t<R, Traits<C, 0, 0, false, false> >(p, st);
}
#endif

template <>
cudaError_t launch_step<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st, const char** name) {
  typedef QS_REAL R;
  if (p.inject) { if (name) *name = "generic-injected"; return launch_step_t<R, TGenInject>(p, st); }
#ifdef QS_FAST_VARIANTS
  if (!p.need_rare) {
    switch (p.controller) {
      case CTRL_RAW: if (name) *name = "fast-raw"; return launch_fast_c<R, CTRL_RAW>(p, st);
      case CTRL_OMEGA_THRUST: if (name) *name = "fast-omega-thrust"; return launch_fast_c<R, CTRL_OMEGA_THRUST>(p, st);
      case CTRL_MELLINGER: if (name) *name = "fast-mellinger"; return launch_fast_c<R, CTRL_MELLINGER>(p, st);
      default: break;
    }
  }
#endif
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
  if (p.inject) dynamics_step_kernel<R, TGenInject><<<grid, BLOCK, 0, st>>>(p);
  else dynamics_step_kernel<R, TGenPhilox><<<grid, BLOCK, 0, st>>>(p);
  return cudaGetLastError();
}

template <>
cudaError_t launch_clock_advance<QS_REAL>(const Params<QS_REAL>& p, cudaStream_t st) {
  clock_advance_kernel<QS_REAL><<<1, 1, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace qs
