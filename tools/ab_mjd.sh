#!/bin/bash
# A/B timing of library variants on Merton levels (default sampler); usage: tools/ab_mjd.sh "<levels>" [level_times args]
lv=$1; shift
for lib in multilevel_mc_sdes_b200/csrc/libmlmcb.so build/variants/lib_*.so; do
  echo "== $lib"
  for l in $lv; do MLMCB_LIB=$PWD/$lib python tools/time_lib.py --model 1 --payoff 3 --M 2 --l $l --batch 16777216 --reps 3 "$@" 2>&1 | tail -1; done
done
