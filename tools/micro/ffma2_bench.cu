// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 (FFMA2) issue throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu -o ffma2_bench && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__global__ void scalar_k(float* out, float x, float y, int iters) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, x, y); a1 = fmaf(a1, x, y); a2 = fmaf(a2, x, y); a3 = fmaf(a3, x, y);
    a4 = fmaf(a4, x, y); a5 = fmaf(a5, x, y); a6 = fmaf(a6, x, y); a7 = fmaf(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void packed_k(float* out, float x, float y, int iters) {
  float t = threadIdx.x;
  unsigned long long a0 = pack(t, t + 1), a1 = pack(t + 2, t + 3), a2 = pack(t + 4, t + 5), a3 = pack(t + 6, t + 7);
  const unsigned long long xx = pack(x, x), yy = pack(y, y);
  for (int i = 0; i < iters; ++i) {
    a0 = ffma2(a0, xx, yy); a1 = ffma2(a1, xx, yy); a2 = ffma2(a2, xx, yy); a3 = ffma2(a3, xx, yy);
  }
  float lo, hi, s = 0.f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a0)); s += lo + hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a1)); s += lo + hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2)); s += lo + hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a3)); s += lo + hi;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 1 << 16;
  float* out;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) scalar_k<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
      else packed_k<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)blocks * threads * iters * 8;
      if (rep == 2) printf("%s: %.3f ms  %.1f TFLOP/s fp32 (2 flop per fma)\n", mode ? "FFMA2 packed" : "FFMA scalar", ms, 2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
