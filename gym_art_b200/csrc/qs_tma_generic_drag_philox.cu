// fp32 run-time-parameterised configuration through the TMA harness: family "drag", philox noise
#define QS_GEN_INJECT 0
#define QS_GEN_FEAT 1
#define QS_GEN_LABEL "drag"
#define QS_GEN_NAME launch_tma_generic_drag_philox
#define QS_GEN_SUFFIX ""
#include "qs_tma_generic.inl"
