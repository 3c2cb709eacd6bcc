#!/bin/bash
# Runs on the GPU box: parity tests, per-config timings, and a light ncu pass (instructions / duration / DRAM bytes).
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for c in cfg2 cfg3 cfg5; do python tools/profile_step.py --config $c --steps 20; done
python tools/profile_step.py --config cfg2 --envs 65536 --steps 100
python tools/profile_step.py --config cfg2 --envs 1048576 --steps 50
for c in cfg2 cfg3 cfg5; do
  python tools/profile_step.py --config $c --steps 4 > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:env_step -s 4 -c 1 --csv --log-file gpurun_out/light_$c.csv python tools/profile_step.py --config $c --steps 4 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/light_$c.csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value")
d={r[i_n]:r[i_v] for r in rows[1:]}
inst=float(d["smsp__inst_executed.sum"].replace(",",""))
print("$c ncu: instr/thread %.0f  dur %s us  dram rd %s wr %s  issue %s%%"%(inst/(4194304/32), d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"], d["smsp__issue_active.avg.pct_of_peak_sustained_active"]))
PY
done
