#!/bin/bash
# A/B of two library builds over the configurations outside the plain shape + the batch-size sweep (runs on the GPU box):
# tools/ab_generic.sh <tag> <lib_a> <lib_b>   (paths relative to the repo)
tag=$1; shift
for lib in "$@"; do
  echo "== $lib: generic_perf"
  QUADSIM_LIB=$PWD/$lib timeout 600 python tools/generic_perf.py 2>&1 | tee gpurun_out/generic_perf_${tag}_$(basename $lib .so).log
done
for lib in "$@"; do
  echo "== $lib: sweep"
  QUADSIM_LIB=$PWD/$lib timeout 600 python tools/sweep.py cfg2 cfg3 2>&1 | tee gpurun_out/sweep_${tag}_$(basename $lib .so).log
done
