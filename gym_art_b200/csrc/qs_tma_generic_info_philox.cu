// fp32 run-time-parameterised configuration through the TMA harness: family "info", philox noise
#define QS_GEN_INJECT 0
#define QS_GEN_FEAT 8
#define QS_GEN_LABEL "info"
#define QS_GEN_NAME launch_tma_generic_info_philox
#define QS_GEN_SUFFIX ""
#include "qs_tma_generic.inl"
