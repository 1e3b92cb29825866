#!/bin/bash
# usage: tools/ncu_case.sh <tag> <kernel-regex> <time_lib args...>
tag=$1; kr=$2; shift 2
python tools/time_lib.py "$@" --reps 1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$kr -c 1 -f -o gpurun_out/$tag python tools/time_lib.py "$@" --reps 1 > gpurun_out/ncu_$tag.log 2>&1
python tools/ncu_summary.py gpurun_out/$tag.ncu-rep > gpurun_out/${tag}_summary.txt
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv
rm -f gpurun_out/$tag.ncu-rep
