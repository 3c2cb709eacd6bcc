// qs_tma_generic.inl - the run-time-parameterised fp32 configurations (controller, OU and sensor noise decided at run time)
// through the same persistent TMA harness as the specialised variants.  QS_GEN_FEAT selects the family:
//   F_ALL     catch-all: every feature group, run-time sim_steps, + its per-env-model twin
//   F_DRAG    rotor drag / linearity / damping only         }
//   F_NOISEX  uniform / quaternion / gyro-bias noise only   }  sim_steps = 2 (substeps unrolled, OU draws hoisted)
//   F_MISC    1D / 2D, goals, quaternion / t2w layouts      }
//   F_INFO    diagnostics buffer only                       }
//   F_DRAG | F_NOISEX   the one two-group pairing with its own kernel (drag model + full sensor noise), sim_steps = 2
// One translation unit per family and noise source, so that they compile in parallel.
#include "qs_device.cuh"
#include "qs_tma_launch.inl"

namespace qs {

cudaError_t QS_GEN_NAME(const Params<float>& p, cudaStream_t st, const char** name) {
#if QS_GEN_FEAT == 15
  if (p.models != nullptr) {
    if (name) *name = "tma-generic-permodel" QS_GEN_SUFFIX;
    return launch_tma_t<Traits<-1, -1, -1, QS_GEN_INJECT != 0, F_ALL, -1, false, false, true> >(p, st);
  }
  if (name) *name = "tma-generic" QS_GEN_SUFFIX;
  return launch_tma_t<Traits<-1, -1, -1, QS_GEN_INJECT != 0, F_ALL> >(p, st);
#else
#if QS_GEN_FEAT == 1 || QS_GEN_FEAT == 4
  // per-env models (dynamics randomisation) with one feature group: drag / linearity / damping rows of the model buffer in
  // use (any RelativeSampler perturbation of the motor constants), or a non-plain step shape (t2w observations, sensor noise
  // layouts, goals ...)
  if (p.models != nullptr) {
    if (name) *name = "tma-generic-permodel-" QS_GEN_LABEL QS_GEN_SUFFIX;
    return launch_tma_t<Traits<-1, -1, -1, QS_GEN_INJECT != 0, QS_GEN_FEAT, 2, false, false, true> >(p, st);
  }
#endif
  if (name) *name = "tma-generic-" QS_GEN_LABEL QS_GEN_SUFFIX;
  return launch_tma_t<Traits<-1, -1, -1, QS_GEN_INJECT != 0, QS_GEN_FEAT, 2> >(p, st);
#endif
}

}  // namespace qs
