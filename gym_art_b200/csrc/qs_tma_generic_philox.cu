// fp32 generic configuration through the TMA harness, philox noise
#define QS_GEN_INJECT 0
#define QS_GEN_NAME launch_tma_generic_philox
#define QS_GEN_SUFFIX ""
#include "qs_tma_generic.inl"
