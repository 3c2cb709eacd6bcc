#!/bin/bash
# A/B of the one-wave residency rule (QUADSIM_OW_REGCAP) and the one-wave / pipelined switch-over (QUADSIM_ONEWAVE_FACTOR) at
# mid-size batches (runs on the GPU box)
for knob in "QUADSIM_OW_REGCAP=0" "QUADSIM_OW_REGCAP=1" "QUADSIM_OW_REGCAP=1 QUADSIM_ONEWAVE_FACTOR=2.3"; do
  echo "== $knob"
  env $knob timeout 300 python tools/mid_sweep.py ${1:-cfg2} 2>&1 | grep "^cfg"
done
