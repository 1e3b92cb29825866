#!/bin/bash
# One line per kernel family at the headline level (l=6, M=4 unless noted): profiles/rNN_timing_matrix.txt
# flags: 0 fp64 sampler (default) | 1 fp32 sampler | 16 fp32 paths | 32 antithetic | 64 Milstein
t() { python tools/time_lib.py "$@" | cut -c42-; }
echo "# default (fp64) sampler"
for p in 0 1 2 3; do t --model 0 --payoff $p; done
t --model 0 --payoff 0 --l 12 --M 2
t --model 0 --payoff 2 --l 12 --M 2
t --model 0 --payoff 0 --flags 64
t --model 0 --payoff 0 --flags 32
t --model 0 --payoff 1 --flags 32
t --model 0 --payoff 1 --flags 96
t --model 0 --payoff 0 --flags 16
for p in 0 1 2 3; do t --model 1 --payoff $p --batch 16777216; done
t --model 1 --payoff 0 --flags 64 --batch 16777216
t --model 0 --payoff 4 --l 12 --M 2 --batch 4194304
echo "# opt-in fp32 sampler"
for p in 0 1 2 3; do t --model 0 --payoff $p --flags 1; done
t --model 0 --payoff 0 --flags 65
t --model 0 --payoff 0 --flags 33
t --model 0 --payoff 0 --flags 17
for p in 0 1 2 3; do t --model 1 --payoff $p --flags 1 --batch 16777216; done
t --model 0 --payoff 4 --l 12 --M 2 --flags 1 --batch 4194304
