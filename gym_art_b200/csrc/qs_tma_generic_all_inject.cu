// fp32 run-time-parameterised configuration through the TMA harness: family "all", inject noise
#define QS_GEN_INJECT 1
#define QS_GEN_FEAT 15
#define QS_GEN_LABEL ""
#define QS_GEN_NAME launch_tma_generic_all_inject
#define QS_GEN_SUFFIX "-injected"
#include "qs_tma_generic.inl"
