// fp32 fast subset, host-injected noise: the same arithmetic as the throughput path, for the parity tests
#define QS_FAST_INJECT 1
#define QS_FAST_NAME launch_fast_inject
#define QS_FAST_SUFFIX "-injected"
#include "qs_fast.inl"
