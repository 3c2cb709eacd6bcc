// fp32 fast subset, controller velyaw, inject noise (one translation unit per controller x noise source: they compile in parallel)
#define QS_FAST_INJECT 1
#define QS_FAST_CTRL qs::CTRL_VELOCITY_YAW
#define QS_FAST_LABEL "tma-velocity-yaw"
#define QS_FAST_NAME launch_fast_velyaw_inject
#define QS_FAST_SUFFIX "-injected"
#include "qs_fast.inl"
