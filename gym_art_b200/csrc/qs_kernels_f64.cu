// fp64 build of the env-step kernels: verification path (matches the float64 reference to 1e-12 per step)
#define QS_REAL double
#include "qs_kernels.inl"
