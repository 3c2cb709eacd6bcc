// fp32 run-time-parameterised configuration through the TMA harness: family "info", inject noise
#define QS_GEN_INJECT 1
#define QS_GEN_FEAT 8
#define QS_GEN_LABEL "info"
#define QS_GEN_NAME launch_tma_generic_info_inject
#define QS_GEN_SUFFIX "-injected"
#include "qs_tma_generic.inl"
