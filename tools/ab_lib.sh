#!/bin/bash
# A/B of library builds on the bench headline (runs on the GPU box): tools/ab_lib.sh <lib> [<lib> ...], paths relative to the repo
for i in 1 2; do
for lib in "$@"; do
  echo "== $lib (pass $i)"
  QUADSIM_LIB=$PWD/$lib python tools/ow_ab.py 65536 2004 | grep -v "^n_envs"
done; done
