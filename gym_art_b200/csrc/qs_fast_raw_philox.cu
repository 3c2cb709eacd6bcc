// fp32 fast subset, controller raw, philox noise (one translation unit per controller x noise source: they compile in parallel)
#define QS_FAST_INJECT 0
#define QS_FAST_CTRL qs::CTRL_RAW
#define QS_FAST_LABEL "tma-raw"
#define QS_FAST_NAME launch_fast_raw_philox
#define QS_FAST_SUFFIX ""
#define QS_FAST_MISC 1
#include "qs_fast.inl"
