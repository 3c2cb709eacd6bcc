#!/bin/bash
# A/B on one GPU box: tile_sync 0 vs 1 for the three single-drone configs, large and small N
python tools/profile_step.py --config cfg2 --steps 30 --copy
for ts in 0 1 0 1; do
  echo "== QUADSIM_TILE_SYNC=$ts"
  for c in cfg2 cfg3 cfg5; do QUADSIM_TILE_SYNC=$ts python tools/profile_step.py --config $c --steps 30; done
  QUADSIM_TILE_SYNC=$ts python tools/profile_step.py --config cfg2 --envs 65536 --steps 200
done
