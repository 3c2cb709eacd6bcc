// fp32 run-time-parameterised configuration through the TMA harness: family "noise", philox noise
#define QS_GEN_INJECT 0
#define QS_GEN_FEAT 2
#define QS_GEN_LABEL "noise"
#define QS_GEN_NAME launch_tma_generic_noise_philox
#define QS_GEN_SUFFIX ""
#include "qs_tma_generic.inl"
