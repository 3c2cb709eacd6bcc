// fp32 generic configuration through the TMA harness, inject noise
#define QS_GEN_INJECT 1
#define QS_GEN_NAME launch_tma_generic_inject
#define QS_GEN_SUFFIX "-injected"
#include "qs_tma_generic.inl"
