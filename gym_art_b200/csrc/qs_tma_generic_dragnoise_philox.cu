// fp32 run-time-parameterised configuration through the TMA harness: family "drag-noise" (rotor drag / linearity / damping AND
// the extra sensor-noise terms: the sim-to-real pairing of the CrazyFlie model), philox noise
#define QS_GEN_INJECT 0
#define QS_GEN_FEAT 3
#define QS_GEN_LABEL "drag-noise"
#define QS_GEN_NAME launch_tma_generic_dragnoise_philox
#define QS_GEN_SUFFIX ""
#include "qs_tma_generic.inl"
