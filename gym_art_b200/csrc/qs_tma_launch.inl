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

template <class T, bool ONEWAVE>
static cudaError_t launch_tma_w(const Params<float>& p, cudaStream_t st, unsigned grid, size_t smem) {
  static bool attr_set[64] = {};   // per instantiation and per device (function attributes are per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(env_step_tma_kernel<T, ONEWAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
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
  const int cap = T::RARE ? TMA_MIN_CTAS_RARE : TMA_MIN_CTAS;
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  const long long slots = (long long)sm_count() * per_sm;
  // one wave: one tile per CTA.  The kernel needs only the observation tile, but it is given the full allocation so
  // that at most per_sm CTAs share an SM: the launch is latency-bound and its time is set by the fullest SM.
  // Up to ~1.8 x slots tiles the one-tile-per-CTA kernel still wins (measured, tools/mid_sweep.py: 9.4 vs 10.6 us at
  // 81,920 envs, 12.0 vs 12.5 at 131,072, 14.4 vs 13.9 at 163,840): the excess CTAs start as residents retire - a natural
  // partial second wave - while the persistent pipeline would run 1-2 tiles per CTA with nothing to overlap.
  static double factor = -1.0;
  if (factor < 0.0) { const char* e = getenv("QUADSIM_ONEWAVE_FACTOR"); factor = e ? atof(e) : 1.8; }
  if ((double)tiles <= factor * (double)slots) return launch_tma_w<T, true>(p, st, (unsigned)tiles, smem);
  return launch_tma_w<T, false>(p, st, (unsigned)slots, smem);                        // persistent CTAs, tile ring
}

}  // namespace qs
