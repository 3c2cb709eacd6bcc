#!/bin/bash
# Runs on the GPU box: step times of the configurations outside the fast subset + one ncu --set full capture per feature kernel.
# usage: tools/gpu_generic.sh <tag> [config ...]
tag=${1:-r2}; shift
cfgs=${@:-drag noise vertical}
mkdir -p gpurun_out
python tools/generic_perf.py > gpurun_out/generic_perf_$tag.log 2>&1
cat gpurun_out/generic_perf_$tag.log
for c in $cfgs; do
  P="python tools/profile_step.py --config $c --envs 4194304 --steps 6"
  timeout 300 $P > gpurun_out/plain_${c}_$tag.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step -s 4 -c 1 -f -o gpurun_out/prof_${c}_$tag $P > gpurun_out/ncu_${c}_$tag.log 2>&1
  cat gpurun_out/plain_${c}_$tag.log
done
