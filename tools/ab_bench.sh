#!/bin/bash
# A/B of library builds through bench.py itself (runs on the GPU box): tools/ab_bench.sh <lib> [<lib> ...]
for i in 1 2; do
for lib in "$@"; do
  QUADSIM_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
o=d['other_workloads']
print('$lib', 'us/step %.3f' % (d['ms_per_step']*1e3), 'frac %.3f' % d['roofline']['frac'], 'l2res %.3f' % (d['l2_resident_same_batch']['ms_per_step']*1e3), 'big %.1f' % (d['hbm_resident']['ms_per_step']*1e3), 'cfg3 %.1f cfg4 %.1f cfg5 %.1f' % (o['cfg3']['ms_per_step']*1e3, o['cfg4']['ms_per_step']*1e3, o['cfg5']['ms_per_step']*1e3), 'e2e %.3e' % d['e2e']['value'], [round(x*1e3,2) for x in d['ms_repeats']])"
done; done
