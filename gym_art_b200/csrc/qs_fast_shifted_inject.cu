// fp32 fast subset, controller shifted, inject noise (one translation unit per controller x noise source: they compile in parallel)
#define QS_FAST_INJECT 1
#define QS_FAST_CTRL qs::CTRL_SHIFTED
#define QS_FAST_LABEL "tma-shifted"
#define QS_FAST_NAME launch_fast_shifted_inject
#define QS_FAST_SUFFIX "-injected"
#include "qs_fast.inl"
