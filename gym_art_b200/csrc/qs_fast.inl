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

template <class T, bool ONEWAVE>
static cudaError_t launch_tma_w(const Params<float>& p, cudaStream_t st, unsigned grid, size_t smem) {
  static bool attr_set = false;   // one flag per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(env_step_tma_kernel<T, ONEWAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(BLOCK);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // measured on B200 (tools/small_sweep.py): for grids under ~320 CTAs the dependent-launch hand-over costs more
  // than the launch ramp it hides (5.1 vs 3.6 us per step at 128 CTAs); above, it wins (6.8 vs 7.5 us at 512 CTAs)
  cfg.numAttrs = (p.pdl && grid >= 320u) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, env_step_tma_kernel<T, ONEWAVE>, p);
}

template <class T>
static cudaError_t launch_tma_t(const Params<float>& p, cudaStream_t st) {
  const long long tiles = (p.end - p.begin + BLOCK - 1) / BLOCK;
  const size_t smem = (size_t)2 * (p.tma_rows + 1) * BLOCK * sizeof(float4) + (size_t)2 * BLOCK * p.obs_dim * sizeof(float);
  int per_sm = (int)((227 * 1024) / (smem + 1024 + 128));
  if (per_sm > TMA_MIN_CTAS) per_sm = TMA_MIN_CTAS;
  if (per_sm < 1) per_sm = 1;
  const long long slots = (long long)sm_count() * per_sm;
  // one wave: one tile per CTA.  The kernel needs only the observation tile, but it is given the full allocation so
  // that at most per_sm CTAs share an SM: the launch is latency-bound and its time is set by the fullest SM.
  if (tiles <= slots) return launch_tma_w<T, true>(p, st, (unsigned)tiles, smem);
  return launch_tma_w<T, false>(p, st, (unsigned)slots, smem);                        // persistent CTAs, tile ring
}

template <int C, int SIM, bool PLAIN, bool SWARM>
static cudaError_t launch_fast_cs(const Params<float>& p, cudaStream_t st) {
  if (p.ou_on) {
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
  }
  if (p.sense_on) return launch_tma_t<Traits<C, 0, 1, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
  return launch_tma_t<Traits<C, 0, 0, QS_FAST_INJECT != 0, false, SIM, PLAIN, SWARM> >(p, st);
}

// the common shape (sim_steps = 2, one env step per launch, plain 18-word observation, no previous-action row) gets
// fully specialised kernels; everything else in the fast subset takes the run-time-parameterised twins
template <int C>
static cudaError_t launch_fast_c(const Params<float>& p, cudaStream_t st) {
  const bool plain = p.sim_steps == 2 && p.k_steps == 1 && p.obs_dim == 18 && !p.obs_rel && p.row_act < 0;
  if (p.models != nullptr) {                           // per-env model constants: common shape only (host guarantees it)
    if (!plain || p.n_agents > 1 || p.log_pos) return cudaErrorInvalidValue;
    if (p.sense_on) return launch_tma_t<Traits<C, 1, 1, QS_FAST_INJECT != 0, false, 2, true, false, true> >(p, st);
    return launch_tma_t<Traits<C, 1, 0, QS_FAST_INJECT != 0, false, 2, true, false, true> >(p, st);
  }
  const bool swarm = p.n_agents > 1 || p.log_pos;      // the host only routes these here when k_steps == 1 and plain
  if (swarm) {
    if (!plain) return cudaErrorInvalidValue;
    return launch_fast_cs<C, 2, true, true>(p, st);
  }
  if (plain) return launch_fast_cs<C, 2, true, false>(p, st);
  return launch_fast_cs<C, -1, false, false>(p, st);
}

cudaError_t QS_FAST_NAME(const Params<float>& p, cudaStream_t st, const char** name) {
  if (name) *name = QS_FAST_LABEL QS_FAST_SUFFIX;
  return launch_fast_c<QS_FAST_CTRL>(p, st);
}

}  // namespace qs
