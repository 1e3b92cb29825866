#!/bin/bash
# usage: tools/ncu_level.sh <tag> <level> <kernel-regex> [env...]
tag=$1; l=$2; kr=$3; shift 3
env "$@" python tools/time_lib.py --model 1 --M 2 --l $l --batch 8388608 --reps 1 || exit 1
env "$@" ncu --set full --clock-control none --import-source on -k regex:$kr -c 1 -f -o gpurun_out/$tag python tools/time_lib.py --model 1 --M 2 --l $l --batch 8388608 --reps 1 > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_summary.py gpurun_out/$tag.ncu-rep > gpurun_out/${tag}_summary.txt
python tools/ncu_hot.py gpurun_out/$tag.ncu-rep > gpurun_out/${tag}_hot.txt
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv
rm -f gpurun_out/$tag.ncu-rep
