#!/bin/bash
# A/B timing of library variants (build/variants/lib_*.so, see multilevel_mc_sdes_b200/build.py compile_lib(only=...))
# on the headline workload with the default (fp64) sampler; usage: tools/ab_f64.sh [time_lib args]
for lib in multilevel_mc_sdes_b200/csrc/libmlmcb.so build/variants/lib_*.so; do
  MLMCB_LIB=$PWD/$lib python tools/time_lib.py --reps 3 "$@" 2>&1 | tail -1
done
