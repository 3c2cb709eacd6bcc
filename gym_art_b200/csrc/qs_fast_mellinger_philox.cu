// fp32 fast subset, controller mellinger, philox noise (one translation unit per controller x noise source: they compile in parallel)
#define QS_FAST_INJECT 0
#define QS_FAST_CTRL qs::CTRL_MELLINGER
#define QS_FAST_LABEL "tma-mellinger"
#define QS_FAST_NAME launch_fast_mellinger_philox
#define QS_FAST_SUFFIX ""
#define QS_FAST_MISC 1
#include "qs_fast.inl"
