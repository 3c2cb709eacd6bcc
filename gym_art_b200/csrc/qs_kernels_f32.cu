// fp32 build of the env-step kernels: the throughput path (generic + specialised "fast" variants)
#define QS_REAL float
#define QS_FAST_VARIANTS 1
#include "qs_kernels.inl"
