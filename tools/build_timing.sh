#!/bin/bash
# Builds the QS_PHASE_TIMING experiment library (raw-controller Philox variants only are needed by tools/phase_timing.py)
set -e
cd "$(dirname "$0")/../gym_art_b200"
mkdir -p _exp
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DQS_PHASE_TIMING"
for f in qs_kernels_f32 qs_kernels_f64 quadsim_abi qs_tma_generic_philox qs_tma_generic_inject qs_fast_raw_philox qs_fast_raw_inject qs_fast_omega_philox qs_fast_omega_inject qs_fast_mellinger_philox qs_fast_mellinger_inject qs_fast_shifted_philox qs_fast_shifted_inject qs_fast_velyaw_philox qs_fast_velyaw_inject; do
  nvcc $F -c csrc/$f.cu -o _exp/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _exp/libquadsim_timing.so _exp/*.o
rm -f _exp/*.o
ls -la _exp/libquadsim_timing.so
