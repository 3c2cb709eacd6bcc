#!/bin/bash
# Runs on the GPU box: ncu launch list of the bench command + one --set full capture of the step kernel per workload, summarised
# THERE into the text files profiles/ keeps (the .ncu-rep files are 18 MB each and gpurun brings back at most 64 MiB; the
# headline capture is kept as a report as well).    usage: tools/gpu_profile.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --steps 1002 --warmup 10 --no-cpu"
$B > gpurun_out/bench_plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/bench_ncu_$tag.log 2>&1
python tools/ncu_summary.py list gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.txt 2>&1
for spec in "cfg2 65536" "cfg2 4194304" "cfg3 4194304" "cfg5 4194304" "cfg4 4194304"; do
  set -- $spec
  P="python tools/profile_step.py --config $1 --envs $2 --steps 6"
  out=/tmp/prof_${1}_$2_$tag
  [ "$2" = "65536" ] && out=gpurun_out/prof_${1}_$2_$tag
  $P > gpurun_out/plain_${1}_$2_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:env_step -s 4 -c 1 -f -o $out $P > gpurun_out/ncu_${1}_$2_$tag.log 2>&1
  python tools/ncu_summary.py rep $out.ncu-rep $2 > gpurun_out/ncu_${1}_$2_$tag.txt 2>&1
  cat gpurun_out/plain_${1}_$2_$tag.log
done
