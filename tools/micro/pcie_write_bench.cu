// Posted-write rate of one B200 into pinned host memory, by store flavour (what bounds qs_step_host's zero-copy path):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pcie_write_bench tools/micro/pcie_write_bench.cu && /tmp/pcie_write_bench
// (a) st.global.v4 per lane (512 B per warp instruction)   (b) cp.async.bulk shared -> host, CHUNK bytes per copy
// (c) copy engine cudaMemcpyAsync D2H of the same bytes      (d) kernel writes device memory, copy engine follows in 4 chunks
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void st_v4(float4* dst, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = make_float4((float)i, 1.f, 2.f, 3.f);
}

__global__ void st_bulk(char* dst, size_t bytes, int chunk) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (int k = threadIdx.x; k < chunk / 4; k += blockDim.x) reinterpret_cast<float*>(sm)[k] = (float)k;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t chunks = bytes / chunk;
    for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk),
                   "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"((uint32_t)chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const size_t bytes = 5046272 / 4096 * 4096;   // ~ the D2H bytes of one 65,536-env step
  char *h, *d;
  CK(cudaHostAlloc((void**)&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  CK(cudaMalloc((void**)&d, bytes));
  char* hd;
  CK(cudaHostGetDevicePointer((void**)&hd, h, 0));
  cudaStream_t st, s2[4];
  CK(cudaStreamCreate(&st));
  for (int k = 0; k < 4; ++k) CK(cudaStreamCreate(&s2[k]));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(st_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int reps = 200;
  auto report = [&](const char* name) {
    float ms; CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-52s %8.1f us  %6.1f GB/s\n", name, ms * 1e3 / reps, bytes / (ms / reps) * 1e-6);
  };
  for (int pass = 0; pass < 2; ++pass) {
    for (int grid : {148, 296, 592, 1184}) {
      for (int i = 0; i < 5; ++i) st_v4<<<grid, 128, 0, st>>>((float4*)hd, bytes / 16);
      CK(cudaEventRecord(e0, st));
      for (int i = 0; i < reps; ++i) st_v4<<<grid, 128, 0, st>>>((float4*)hd, bytes / 16);
      CK(cudaEventRecord(e1, st));
      char nm[96]; snprintf(nm, sizeof nm, "st.global.v4 -> host, grid %d x 128", grid);
      report(nm);
    }
    for (int chunk : {2048, 4096, 16384, 65536}) {
      for (int grid : {37, 148}) {
        for (int i = 0; i < 5; ++i) st_bulk<<<grid, 128, chunk, st>>>(hd, bytes, chunk);
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < reps; ++i) st_bulk<<<grid, 128, chunk, st>>>(hd, bytes, chunk);
        CK(cudaEventRecord(e1, st));
        char nm[96]; snprintf(nm, sizeof nm, "cp.async.bulk smem -> host, %d B chunks, grid %d", chunk, grid);
        report(nm);
      }
    }
    for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(e1, st));
    report("copy engine D2H, one copy");
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i)
      for (int k = 0; k < 4; ++k) CK(cudaMemcpyAsync(h + k * (bytes / 4), d + k * (bytes / 4), bytes / 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(e1, st));
    report("copy engine D2H, four copies back to back");
    // host-timed synchronous round: one D2H + stream synchronize per step (what a synchronous API call pays)
    CK(cudaStreamSynchronize(st));
    {
      struct timespec a, b; clock_gettime(CLOCK_MONOTONIC, &a);
      for (int i = 0; i < reps; ++i) { CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
      clock_gettime(CLOCK_MONOTONIC, &b);
      double us = ((b.tv_sec - a.tv_sec) * 1e9 + (b.tv_nsec - a.tv_nsec)) / 1e3 / reps;
      printf("%-52s %8.1f us  %6.1f GB/s (host clock)\n", "copy engine D2H + synchronize per step", us, bytes / us * 1e-3);
      clock_gettime(CLOCK_MONOTONIC, &a);
      for (int i = 0; i < reps; ++i) { st_v4<<<296, 128, 0, st>>>((float4*)hd, bytes / 16); CK(cudaStreamSynchronize(st)); }
      clock_gettime(CLOCK_MONOTONIC, &b);
      us = ((b.tv_sec - a.tv_sec) * 1e9 + (b.tv_nsec - a.tv_nsec)) / 1e3 / reps;
      printf("%-52s %8.1f us  %6.1f GB/s (host clock)\n", "st.global.v4 -> host + synchronize per step", us, bytes / us * 1e-3);
    }
  }
  return 0;
}
