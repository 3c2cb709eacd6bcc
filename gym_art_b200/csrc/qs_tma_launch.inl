// qs_tma_launch.inl - host-side launch of env_step_tma_kernel: grid sizing, one-wave vs pipelined instantiation, PDL.
// Included by the translation units that instantiate the kernel (qs_fast.inl, qs_tma_generic.inl).
#pragma once
#include <stdlib.h>

#include "qs_device.cuh"

namespace qs {

static int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

constexpr size_t SMEM_PER_SM = 228 * 1024, SMEM_MAX_CTA = 227 * 1024, SMEM_RESERVED = 1024 + 256;   // driver reserve + static

template <class T, bool ONEWAVE>
static cudaError_t launch_tma_w(const Params<float>& p, cudaStream_t st, unsigned grid, unsigned threads, size_t smem, bool pdl) {
  static bool attr_set[64] = {};   // per instantiation and per device (function attributes are per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(env_step_tma_kernel<T, ONEWAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(SMEM_MAX_CTA - 256));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, env_step_tma_kernel<T, ONEWAVE>, p);
}

template <class T>
static cudaError_t launch_tma_t(const Params<float>& p_in, cudaStream_t st) {
  Params<float> p = p_in;
  const long long n = p.end - p.begin;
  const long long tiles = (n + BLOCK - 1) / BLOCK;
  // (+ a two-slot stage for 1- and 2-wide actions, which the threads prefetch themselves: see act_direct in the kernel)
  const size_t act_stage = (p.actions != nullptr && p.act_dim < 4) ? (size_t)2 * BLOCK * 2 * sizeof(float) : 0;
  const size_t ring = (size_t)2 * (p.tma_rows + 1) * BLOCK * sizeof(float4) + act_stage, obs_tile = (size_t)BLOCK * p.obs_dim * sizeof(float);
  const int cap = tma_min_ctas<T>();
  // two observation tiles (the bulk store of one drains while the next is filled) unless dropping to one lets another CTA
  // fit on the SM: configurations with extra state rows (gyro bias, previous action, goal) are otherwise held to 3 CTAs = 12
  // warps per SM by shared memory, and these kernels are issue-bound
  size_t smem = ring + 2 * obs_tile;
  int per_sm = (int)(SMEM_PER_SM / (smem + SMEM_RESERVED));
  p.obs_stages = 2;
  if (per_sm < cap && (int)(SMEM_PER_SM / (ring + obs_tile + SMEM_RESERVED)) > per_sm) {
    smem = ring + obs_tile;
    per_sm = (int)(SMEM_PER_SM / (smem + SMEM_RESERVED));
    p.obs_stages = 1;
  }
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  const int sms = sm_count();
  const long long slots = (long long)sms * per_sm;
  // Up to ~2.3 x slots tiles the one-tile-per-CTA kernel still wins (measured, tools/mid_sweep.py + tools/mid_ab.sh with the
  // equal-depth waves below: 7.2 us at 81,920 envs, 10.6 at 131,072, 12.0 vs 12.7 at 147,456, 13.1 vs 13.5 at 163,840; three
  // waves at 196,608 lose to the pipeline's 14.9): the CTAs beyond the first wave start as residents retire, while the
  // persistent pipeline would run 1-2 tiles per CTA with nothing to overlap.
  static double factor = -1.0;
  if (factor < 0.0) { const char* e = getenv("QUADSIM_ONEWAVE_FACTOR"); factor = e ? atof(e) : 2.3; }
  // Programmatic dependent launch at every size.  (Round 1 measured the hand-over costing more than it hid under ~40,960 envs;
  // with the L2 prefetch, the speculative draws and the early clock advance ahead of griddepcontrol.wait it now wins
  // everywhere - tools/small_sweep.py: 1,024 envs 3.79 -> 2.65 us per step, 16,384 envs 4.18 -> 3.03, 32,768 envs 5.28 -> 3.84.)
  static long long pdl_min = -1;                       // A/B knob QUADSIM_PDL_MIN_ENVS
  if (pdl_min < 0) { const char* e = getenv("QUADSIM_PDL_MIN_ENVS"); pdl_min = e ? atoll(e) : 0; }
  const bool pdl = p.pdl && n >= pdl_min;
  if ((double)tiles <= factor * (double)slots) {
    // One wave: one tile per CTA, no ring.  The launch is latency-bound and its time is set by the fullest SM, so the
    // CTAs are made small (p.ow_threads = 64: two warps) and the shared-memory request is sized so that exactly
    // ceil(CTAs / SMs) of them fit on an SM (never more than the register file holds): 65,536 envs = 1,024 CTAs = 7 per
    // SM = 14 warps on every SM, instead of 16 warps on 68 SMs and 12 on the other 80 with 128-thread CTAs.
    const unsigned threads = (p.ow_threads == 32 || p.ow_threads == 64) ? (unsigned)p.ow_threads : (unsigned)BLOCK;
    const long long ctas = (n + threads - 1) / threads;
    long long want = (ctas + sms - 1) / sms;
    // resident CTAs per SM the register file allows for THIS instantiation at this CTA size (asked of the runtime once per
    // device; the one-wave instantiations need 80 - 100 registers, well under the 128 the launch bound promises: ten 64-thread
    // CTAs instead of eight -> up to 94,720 envs in one evenly spread wave: 81,920 envs 9.5 -> 7.1 us)
    static int occ_cache[64][3] = {};
    const int ti = threads == 32 ? 0 : (threads == 64 ? 1 : 2);
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = (dev >= 0 && dev < 64) ? occ_cache[dev][ti] : 0;
    if (occ == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, env_step_tma_kernel<T, true>, (int)threads,
                                                        (size_t)threads * p.obs_dim * sizeof(float)) != cudaSuccess || occ <= 0) {
        cudaGetLastError();
        occ = cap * (BLOCK / (int)threads);
      }
      if (dev >= 0 && dev < 64) occ_cache[dev][ti] = occ;
    }
    static int reg_cap_knob = -1;                      // A/B knob QUADSIM_OW_REGCAP: 0 = the launch-bound formula (round 2: 8)
    if (reg_cap_knob < 0) { const char* e = getenv("QUADSIM_OW_REGCAP"); reg_cap_knob = e ? atoi(e) : 1; }
    const long long reg_cap = reg_cap_knob ? (long long)occ : (long long)cap * (BLOCK / threads);
    // More CTAs than fit at once: waves of EQUAL depth rather than one full wave and a remainder (131,072 envs = 2,048 CTAs: two
    // waves of 7 per SM take 11.1 us, 10 + 4 per SM 12.1 us - a wave costs its latency chain however thin it is).
    if (reg_cap_knob && want > reg_cap) {
      const long long waves = (want + reg_cap - 1) / reg_cap;
      want = (want + waves - 1) / waves;
    }
    const int extra = p.ow_extra;                      // A/B knob: CTA slots per SM left free for the dependent launch
    want += extra;
    if (want > reg_cap + extra) want = reg_cap + extra;
    // Host-buffer launches (qs_step_host, zero-copy): the CTAs read their actions and write their results across PCIe.  With
    // every CTA resident at once all of them read first and then all of them write; fewer resident CTAs = several waves, so the
    // action reads of later CTAs travel (the link is full duplex) while earlier ones are posting their results.
    if (p.host_ctas > 0 && p.host_ctas < want) want = p.host_ctas;
    if (want < 1) want = 1;
    size_t dyn = (SMEM_PER_SM / (size_t)want) - SMEM_RESERVED;
    dyn &= ~(size_t)127;
    if (dyn > SMEM_MAX_CTA - 512) dyn = SMEM_MAX_CTA - 512;
    const size_t need = (size_t)threads * p.obs_dim * sizeof(float);
    if (dyn < need) dyn = need;
    return launch_tma_w<T, true>(p, st, (unsigned)ctas, threads, dyn, pdl);
  }
  return launch_tma_w<T, false>(p, st, (unsigned)slots, BLOCK, smem, pdl);           // persistent CTAs, tile ring
}

}  // namespace qs
