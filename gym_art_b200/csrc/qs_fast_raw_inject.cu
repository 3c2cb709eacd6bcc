// fp32 fast subset, controller raw, inject noise (one translation unit per controller x noise source: they compile in parallel)
#define QS_FAST_INJECT 1
#define QS_FAST_CTRL qs::CTRL_RAW
#define QS_FAST_LABEL "tma-raw"
#define QS_FAST_NAME launch_fast_raw_inject
#define QS_FAST_SUFFIX "-injected"
#define QS_FAST_MISC 1
#include "qs_fast.inl"
