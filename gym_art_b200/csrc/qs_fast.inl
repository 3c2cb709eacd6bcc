// qs_fast.inl - the fp32 fast subset: TMA-pipelined persistent kernel, compile-time controller / OU / sensor-noise
// switches.  Included twice (Philox and injected noise) so the two sets compile in parallel; the injected set
// exists so that the parity tests exercise exactly the arithmetic the throughput path runs.
#include "qs_device.cuh"

namespace qs {

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <class T>
static cudaError_t launch_tma_t(const Params<float>& p, cudaStream_t st) {
  const long long tiles = (p.end - p.begin + BLOCK - 1) / BLOCK;
  const size_t smem = (size_t)2 * (p.tma_rows + 1) * BLOCK * sizeof(float4) + (size_t)2 * BLOCK * p.obs_dim * sizeof(float);
  static bool attr_set = false;   // one flag per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(env_step_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024 + 128));
  if (per_sm > TMA_MIN_CTAS) per_sm = TMA_MIN_CTAS;
  if (per_sm < 1) per_sm = 1;
  long long slots = (long long)sm_count() * per_sm;
  const unsigned grid = (unsigned)(tiles < slots ? tiles : slots);
  env_step_tma_kernel<T><<<grid, BLOCK, smem, st>>>(p);
  return cudaGetLastError();
}

template <int C>
static cudaError_t launch_fast_c(const Params<float>& p, cudaStream_t st) {
  if (p.ou_on) {
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, false> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, false> >(p, st);
  }
  if (p.sense_on) return launch_tma_t<Traits<C, 0, 1, QS_FAST_INJECT != 0, false> >(p, st);
  return launch_tma_t<Traits<C, 0, 0, QS_FAST_INJECT != 0, false> >(p, st);
}

cudaError_t QS_FAST_NAME(const Params<float>& p, cudaStream_t st, const char** name) {
  switch (p.controller) {
    case CTRL_RAW: if (name) *name = "tma-raw" QS_FAST_SUFFIX; return launch_fast_c<CTRL_RAW>(p, st);
    case CTRL_OMEGA_THRUST: if (name) *name = "tma-omega-thrust" QS_FAST_SUFFIX; return launch_fast_c<CTRL_OMEGA_THRUST>(p, st);
    default: if (name) *name = "tma-mellinger" QS_FAST_SUFFIX; return launch_fast_c<CTRL_MELLINGER>(p, st);
  }
}

}  // namespace qs
