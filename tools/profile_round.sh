#!/bin/bash
# usage (GPU box, via gpurun): tools/profile_round.sh <tag>
# 1. launch list of a short bench run (every launch with its device time), 2. one full capture of the top kernel.
tag=${1:-r01}
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --md-steps 5 --equil 30 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_force2 -s 40 -c 1 -o gpurun_out/pair_$tag python bench.py $ARGS > gpurun_out/ncu_full_$tag.log 2>&1
tail -1 gpurun_out/ncu_full_$tag.log
