// fp32 fast subset, on-device Philox noise: the throughput path
#define QS_FAST_INJECT 0
#define QS_FAST_NAME launch_fast_philox
#define QS_FAST_SUFFIX ""
#include "qs_fast.inl"
