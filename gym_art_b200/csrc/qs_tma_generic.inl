// qs_tma_generic.inl - the run-time-parameterised fp32 configuration (rotor drag / linearity / damping, full sensor
// noise, 1D / 2D modes, per-env goals, info buffer, K-step swarms, per-env models outside the common shape) through the
// same persistent TMA harness as the specialised variants.  One translation unit per noise source.
#include "qs_device.cuh"
#include "qs_tma_launch.inl"

namespace qs {

cudaError_t QS_GEN_NAME(const Params<float>& p, cudaStream_t st, const char** name) {
  if (p.models != nullptr) {
    if (name) *name = "tma-generic-permodel" QS_GEN_SUFFIX;
    return launch_tma_t<Traits<-1, -1, -1, QS_GEN_INJECT != 0, true, -1, false, false, true> >(p, st);
  }
  if (name) *name = "tma-generic" QS_GEN_SUFFIX;
  return launch_tma_t<Traits<-1, -1, -1, QS_GEN_INJECT != 0, true> >(p, st);
}

}  // namespace qs
