#!/bin/bash
# profiles/rNN_gbm_levels.txt and rNN_merton_levels.txt: usage tools/level_tables.sh gbm|merton [flags]
what=$1; flags=${2:-0}
if [ "$what" = gbm ]; then
  for c in Euro_GBM Asian_GBM Lookback_GBM; do
    python tools/level_times.py --cls $c --M 2 --Lmax 12 --flags $flags
    python tools/level_times.py --cls $c --M 4 --Lmax 6 --flags $flags
  done
  python tools/level_times.py --cls Amer_GBM --M 2 --Lmax 9 --flags $flags
else
  for c in Euro_Merton Asian_Merton Lookback_Merton Digital_Merton; do
    python tools/level_times.py --cls $c --M 2 --Lmax 12 --flags $flags
    python tools/level_times.py --cls $c --M 4 --Lmax 6 --flags $flags
  done
fi
