// fp32 run-time-parameterised configuration through the TMA harness: family "all", philox noise
#define QS_GEN_INJECT 0
#define QS_GEN_FEAT 15
#define QS_GEN_LABEL ""
#define QS_GEN_NAME launch_tma_generic_all_philox
#define QS_GEN_SUFFIX ""
#include "qs_tma_generic.inl"
