#!/bin/bash
# usage (GPU box): tools/ncu_pair3.sh <tag>  -- one full ncu capture of the three-thread pair engine on C1 (15 000 particles)
tag=$1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:k_pair_force2<\(int\)0, \(bool\)1, \(bool\)1, \(int\)3" -s 30 -c 1 -f -o gpurun_out/pair3_$tag \
  python bench.py --lipids 5000 --steps 2 --warmup 3 --md-steps 8 --equil 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pair3_$tag.log 2>&1
tail -2 gpurun_out/ncu_pair3_$tag.log | cut -c1-200
