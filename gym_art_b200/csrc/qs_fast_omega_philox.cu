// fp32 fast subset, controller omega, philox noise (one translation unit per controller x noise source: they compile in parallel)
#define QS_FAST_INJECT 0
#define QS_FAST_CTRL qs::CTRL_OMEGA_THRUST
#define QS_FAST_LABEL "tma-omega-thrust"
#define QS_FAST_NAME launch_fast_omega_philox
#define QS_FAST_SUFFIX ""
#include "qs_fast.inl"
