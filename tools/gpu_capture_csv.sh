#!/bin/bash
# Runs on the GPU box: one ncu --set full capture per configuration, exported THERE to the raw and source-page CSVs (gzipped;
# the .ncu-rep files are 18 MB each and gpurun brings back at most 64 MiB).   usage: tools/gpu_capture_csv.sh <tag> <config ...>
tag=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  P="python tools/profile_step.py --config $c --envs 4194304 --steps 6"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step -s 4 -c 1 -f -o /tmp/prof_${c}_$tag $P > gpurun_out/ncu_${c}_$tag.log 2>&1
  ncu -i /tmp/prof_${c}_$tag.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/raw_${c}_$tag.csv.gz
  ncu -i /tmp/prof_${c}_$tag.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/src_${c}_$tag.csv.gz
  tail -2 gpurun_out/ncu_${c}_$tag.log
done
ls -la gpurun_out | tail -12
