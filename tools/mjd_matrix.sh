#!/bin/bash
# per-level Merton timings, block-uniform vs event-driven kernel and run_blocks settings
run() { echo "== $*"; env "$@" python tools/level_times.py --cls Digital_Merton --M 2 --Lmax 11 --n 4194304 | awk '{print $3, $4, $5, $10, $11, $12}'; }
run MLMCB_MJD_FLAT_BELOW=0
run MLMCB_MJD_FLAT_BELOW=100000
run MLMCB_MJD_FLAT_BELOW=100000 MLMCB_MJD_RUN_BLOCKS=4
run MLMCB_MJD_FLAT_BELOW=100000 MLMCB_MJD_RUN_BLOCKS=16
run MLMCB_MJD_FLAT_BELOW=100000 MLMCB_MJD_RUN_BLOCKS=64
run MLMCB_MJD_FLAT_BELOW=100000 MLMCB_MJD_RUN_BLOCKS=100000
