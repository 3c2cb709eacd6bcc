// Does an L2 prefetch hint to an unmapped / freed global address fault on sm_100a?  (It must not: the step kernel issues
// such hints for buffers whose lifetime it does not own.)   nvcc -arch=sm_100a prefetch_fault.cu -o prefetch_fault
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const char* p, unsigned bytes, int bulk) {
  if (bulk) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
  else asm volatile("prefetch.global.L2 [%0];" ::"l"(p + threadIdx.x * 128) : "memory");
}
int main() {
  char* d = nullptr;
  cudaMalloc(&d, 64 << 20);
  const char* targets[3] = {d, nullptr, nullptr};
  char* freed = nullptr;
  cudaMalloc(&freed, 1 << 30);
  cudaFree(freed);                                   // a 1 GiB block given back to the driver: unmapped now
  targets[1] = freed;
  targets[2] = (const char*)0x7f00dead0000ull;       // never mapped
  const char* names[3] = {"mapped", "freed", "never mapped"};
  for (int t = 0; t < 3; ++t)
    for (int bulk = 0; bulk < 2; ++bulk) {
      k<<<4, 64>>>(targets[t], 1024, bulk);
      cudaError_t e = cudaDeviceSynchronize();
      printf("%-13s %-26s -> %s\n", names[t], bulk ? "cp.async.bulk.prefetch.L2" : "prefetch.global.L2", cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
