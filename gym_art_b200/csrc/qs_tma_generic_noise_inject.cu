// fp32 run-time-parameterised configuration through the TMA harness: family "noise", inject noise
#define QS_GEN_INJECT 1
#define QS_GEN_FEAT 2
#define QS_GEN_LABEL "noise"
#define QS_GEN_NAME launch_tma_generic_noise_inject
#define QS_GEN_SUFFIX "-injected"
#include "qs_tma_generic.inl"
