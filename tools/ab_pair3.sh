#!/bin/bash
# usage (GPU box): tools/ab_pair3.sh -- A/B of the three-threads-per-particle pair engine (SMD_PAIR3) on small and large systems
for L in 5000 10000 20000 80000; do
  for e in 0 1; do
    echo -n "lipids $L SMD_PAIR3=$e  "; SMD_PAIR3=$e tools/bench_phases.sh --lipids $L
  done
done
