// qs_fast.inl - the fp32 fast subset: TMA-pipelined persistent kernel, compile-time controller / OU / sensor-noise
// switches.  Included twice (Philox and injected noise) so the two sets compile in parallel; the injected set
// exists so that the parity tests exercise exactly the arithmetic the throughput path runs.
#include "qs_device.cuh"
#include "qs_tma_launch.inl"

namespace qs {

template <int C, int SIM, bool PLAIN, bool SWARM>
static cudaError_t launch_fast_cs(const Params<float>& p, cudaStream_t st) {
  if (p.ou_on) {
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
  }
  if (p.sense_on) return launch_tma_t<Traits<C, 0, 1, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
  return launch_tma_t<Traits<C, 0, 0, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
}

// the common shape (sim_steps = 2, one env step per launch, plain 18-word observation, no previous-action row) gets
// fully specialised kernels; everything else in the fast subset takes the run-time-parameterised twins
#ifdef QS_FAST_MISC
// the two controllers the reference env selects itself (raw_control=True / False) also get the F_MISC feature group with
// compile-time controller / OU / sensor-noise switches: per-env and moving goals, quaternion / t2w / t2t observations, the
// restricted spawn law, swarms outside the common shape, per-env models in any of those (host: feat == F_MISC, sim_steps == 2)
template <int C>
static cudaError_t launch_fast_misc(const Params<float>& p, cudaStream_t st) {
  if (p.models != nullptr) {                           // per-env models (the OU row is always there): t2w observations, swarms ...
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, F_MISC, 2, false, false, true> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, F_MISC, 2, false, false, true> >(p, st);
  }
  if (p.ou_on) {
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, F_MISC, 2> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, F_MISC, 2> >(p, st);
  }
  if (p.sense_on) return launch_tma_t<Traits<C, 0, 1, QS_FAST_INJECT != 0, F_MISC, 2> >(p, st);
  return launch_tma_t<Traits<C, 0, 0, QS_FAST_INJECT != 0, F_MISC, 2> >(p, st);
}
#endif

template <int C>
static cudaError_t launch_fast_c(const Params<float>& p, cudaStream_t st) {
#ifdef QS_FAST_MISC
  if (p.feat == F_MISC) return p.sim_steps == 2 ? launch_fast_misc<C>(p, st) : cudaErrorInvalidValue;
  if (p.feat == F_NOISEX) {                            // uniform / quaternion / gyro-bias noise (sensor noise is on by definition)
    if (p.sim_steps != 2 || p.models != nullptr || !p.sense_on) return cudaErrorInvalidValue;
    if (p.ou_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, F_NOISEX, 2> >(p, st);
    return launch_tma_t<Traits<C, 0, 1, QS_FAST_INJECT != 0, F_NOISEX, 2> >(p, st);
  }
#endif
  const bool plain = p.sim_steps == 2 && p.k_steps == 1 && p.obs_dim == 18 && !p.obs_rel && p.row_act < 0;
  if (p.models != nullptr) {                           // per-env model constants: common shape only (host guarantees it)
    if (!plain || p.n_agents > 1 || p.log_pos) return cudaErrorInvalidValue;
    if (p.feat == F_DRAG) {   // models with linearity != 1 / rotor drag / damping (any RelativeSampler perturbation of the motors)
      if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, F_DRAG, 2, true, false, true> >(p, st);
      return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, F_DRAG, 2, true, false, true> >(p, st);
    }
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, F_NONE, 2, true, false, true> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, F_NONE, 2, true, false, true> >(p, st);
  }
  const bool swarm = p.n_agents > 1 || p.log_pos;      // the host only routes these here when k_steps == 1 and plain
  if (swarm) {
    if (!plain) return cudaErrorInvalidValue;
    return launch_fast_cs<C, 2, true, true>(p, st);
  }
  if (plain) return launch_fast_cs<C, 2, true, false>(p, st);
  // other observation layouts / previous-action row / K-step launches at the default sim_steps still get the unrolled
  // substeps and hoisted OU draws
  if (p.sim_steps == 2) return launch_fast_cs<C, 2, false, false>(p, st);
  return launch_fast_cs<C, -1, false, false>(p, st);
}

cudaError_t QS_FAST_NAME(const Params<float>& p, cudaStream_t st, const char** name) {
  if (name) *name = QS_FAST_LABEL QS_FAST_SUFFIX;
  return launch_fast_c<QS_FAST_CTRL>(p, st);
}

}  // namespace qs
