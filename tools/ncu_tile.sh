#!/bin/bash
# usage (GPU box): tools/ncu_tile.sh <tag> [lib.so]  -- one full ncu capture of a force launch of k_pair_tile on C2 (bench workload)
tag=$1; lib=${2:-softmold_b200/libsoftmold_b200.so}
mkdir -p gpurun_out
SOFTMOLD_B200_LIB=$PWD/$lib ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:k_pair_tile<\(int\)0" -s 30 -c 1 -f -o gpurun_out/tile_$tag \
  python bench.py --steps 2 --warmup 3 --md-steps 8 --equil 32 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tile_$tag.log 2>&1
tail -2 gpurun_out/ncu_tile_$tag.log | cut -c1-300
