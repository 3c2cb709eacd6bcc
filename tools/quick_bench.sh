#!/bin/bash
# quick A/B on the GPU box: bench headline numbers + per-config large-N timings
python bench.py --steps ${1:-20040} --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('us/step %.3f  value %.3e  frac %.3f | big frac %.3f us %.1f | e2e %.3e' % (d['ms_per_step']*1e3, d['value'], d['roofline']['frac'], d['hbm_resident']['roofline']['frac'], d['hbm_resident']['ms_per_step']*1e3, d['e2e']['value']))"
for c in cfg2 cfg3 cfg5 cfg4; do python tools/profile_step.py --config $c --steps 30; done
