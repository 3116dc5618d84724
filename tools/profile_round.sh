#!/bin/bash
# usage (GPU box, via gpurun): tools/profile_round.sh <tag>
# 1. launch list of a short bench run (every launch with its device time), 2. one full capture of the top kernel,
# 3. full captures of the other kernels of the step (one launch each).
tag=${1:-r01}
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --md-steps 8 --equil 32 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_pair_force2<\(int\)0" -s 40 -c 1 -f -o gpurun_out/pair_$tag python bench.py $ARGS > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_chain_kick|k_reorder|k_pair_force2<\(int\)2|k_place|k_bin|k_scan|k_final_sum" -s 150 -c 10 -f -o gpurun_out/others_$tag python bench.py $ARGS > gpurun_out/ncu_others_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_full_$tag.log; tail -n 2 gpurun_out/ncu_others_$tag.log
