/* abi_smoke.c - drives libquadsim.so from plain C: no Python, no torch, only the CUDA runtime for buffers.
 *
 *   gcc tests/c/abi_smoke.c -Iinclude -I/usr/local/cuda/include -L/usr/local/cuda/lib64 -lcudart -ldl -o abi_smoke
 *   ./abi_smoke gym_art_b200/libquadsim.so config.bin steps
 *
 * config.bin holds the bytes of a QsConfig (written by the test from the Python host pipeline); the program checks
 * sizeof(QsConfig) against it - the C header and the ctypes mirror must agree on the struct layout - resets, runs
 * `steps` env steps with zero actions (Mellinger controller: the drones must fly to the goal) and prints summary
 * statistics that the test compares with the same run made through the Python binding. */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "quadsim.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define QS(x) do { int rc_ = (x); if (rc_ != QS_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, p_last_error(h)); return 3; } } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: abi_smoke lib config.bin steps\n"); return 1; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  int (*p_abi)(void) = (int (*)(void))dlsym(lib, "qs_abi_version");
  int (*p_layout)(const QsConfig*, QsStateLayout*) = (int (*)(const QsConfig*, QsStateLayout*))dlsym(lib, "qs_state_layout");
  int (*p_create)(const QsConfig*, QsHandle**) = (int (*)(const QsConfig*, QsHandle**))dlsym(lib, "qs_create");
  int (*p_destroy)(QsHandle*) = (int (*)(QsHandle*))dlsym(lib, "qs_destroy");
  const char* (*p_last_error)(const QsHandle*) = (const char* (*)(const QsHandle*))dlsym(lib, "qs_last_error");
  int (*p_bind)(QsHandle*, void*) = (int (*)(QsHandle*, void*))dlsym(lib, "qs_bind_state");
  int (*p_reset)(QsHandle*, const uint8_t*, void*, qs_stream_t) = (int (*)(QsHandle*, const uint8_t*, void*, qs_stream_t))dlsym(lib, "qs_reset");
  int (*p_step)(QsHandle*, const void*, void*, void*, uint8_t*, qs_stream_t) =
      (int (*)(QsHandle*, const void*, void*, void*, uint8_t*, qs_stream_t))dlsym(lib, "qs_step");
  int (*p_stats)(QsHandle*, double*, int, qs_stream_t) = (int (*)(QsHandle*, double*, int, qs_stream_t))dlsym(lib, "qs_get_stats");
  int (*p_check)(QsHandle*, qs_stream_t) = (int (*)(QsHandle*, qs_stream_t))dlsym(lib, "qs_check");
  if (!p_abi || !p_layout || !p_create || !p_destroy || !p_last_error || !p_bind || !p_reset || !p_step || !p_stats || !p_check) {
    fprintf(stderr, "missing symbol\n"); return 1;
  }
  if (p_abi() != QS_ABI_VERSION) { fprintf(stderr, "ABI %d != header %d\n", p_abi(), QS_ABI_VERSION); return 1; }

  QsConfig cfg;
  FILE* f = fopen(argv[2], "rb");
  if (!f) { perror("config"); return 1; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz != (long)sizeof(QsConfig)) { fprintf(stderr, "sizeof(QsConfig) = %zu but config.bin has %ld bytes\n", sizeof(QsConfig), sz); return 4; }
  if (fread(&cfg, 1, sizeof(cfg), f) != sizeof(cfg)) return 4;
  fclose(f);
  if (cfg.struct_size != (int32_t)sizeof(QsConfig) || cfg.dtype != QS_F32) { fprintf(stderr, "bad config\n"); return 4; }
  const int steps = atoi(argv[3]);
  QsHandle* h = NULL;
  QsStateLayout lay;
  if (p_layout(&cfg, &lay) != QS_OK) { fprintf(stderr, "layout: %s\n", p_last_error(NULL)); return 3; }
  const long long n = cfg.n_envs;

  float *state, *obs, *rew; uint8_t* done;
  CK(cudaSetDevice(cfg.device));
  CK(cudaMalloc((void**)&state, (size_t)lay.rows * n * 4 * sizeof(float)));
  CK(cudaMemset(state, 0, (size_t)lay.rows * n * 4 * sizeof(float)));
  CK(cudaMalloc((void**)&obs, (size_t)n * lay.obs_dim * sizeof(float)));
  CK(cudaMalloc((void**)&rew, (size_t)n * sizeof(float)));
  CK(cudaMalloc((void**)&done, (size_t)n));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  if (p_create(&cfg, &h) != QS_OK) { fprintf(stderr, "create: %s\n", p_last_error(NULL)); return 3; }
  QS(p_bind(h, state));
  QS(p_reset(h, NULL, obs, st));
  for (int t = 0; t < steps; ++t) QS(p_step(h, NULL, obs, rew, done, st));    /* NULL actions: Mellinger ignores them */
  QS(p_check(h, st));
  CK(cudaStreamSynchronize(st));

  float* hobs = (float*)malloc((size_t)n * lay.obs_dim * sizeof(float));
  float* hrew = (float*)malloc((size_t)n * sizeof(float));
  CK(cudaMemcpy(hobs, obs, (size_t)n * lay.obs_dim * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hrew, rew, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  double dist = 0.0, rsum = 0.0, osum = 0.0;
  for (long long i = 0; i < n; ++i) {
    const float* o = hobs + i * lay.obs_dim;
    dist += sqrt((double)o[0] * o[0] + (double)o[1] * o[1] + (double)o[2] * o[2]);
    rsum += hrew[i];
    for (int k = 0; k < lay.obs_dim; ++k) osum += o[k];
  }
  double stats[3];
  QS(p_stats(h, stats, 0, st));
  printf("n %lld rows %d obs_dim %d mean_dist %.9g mean_rew %.9g obs_sum %.9g episodes %.0f\n", n, lay.rows, lay.obs_dim, dist / n,
         rsum / n, osum, stats[2]);
  QS(p_destroy(h));
  cudaFree(state); cudaFree(obs); cudaFree(rew); cudaFree(done);
  free(hobs); free(hrew);
  return 0;
}
