#!/bin/bash
# warp-instructions per env-step thread of the step kernel (ncu, one launch) + the headline A/B timing.  usage: tools/inst_count.sh [envs] [config]
n=${1:-65536}; c=${2:-cfg2}
P="python tools/profile_step.py --config $c --envs $n --steps 6"
$P > /dev/null 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:env_step -s 4 -c 1 --csv --log-file gpurun_out/inst_$c.csv $P > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/inst_$c.csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value")
d={r[i_n]:r[i_v] for r in rows[1:]}
inst=float(d["smsp__inst_executed.sum"].replace(",",""))
print("$c $n envs: warp-instructions per thread %.0f, ncu duration %s us" % (inst/($n/32), d["gpu__time_duration.sum"]))
PY
