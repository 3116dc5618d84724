#!/bin/bash
for lib in lib_s64_b4.so lib_s64_b3.so lib_s64_b2.so; do
  echo "== $lib"
  SOFTMOLD_B200_LIB=$PWD/softmold_b200/$lib SMD_PAIR_ENGINE=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -k "pair_force_potential or total_force or periodic_gas or overfull" 2>&1 | tail -2
  SOFTMOLD_B200_LIB=$PWD/softmold_b200/$lib python tools/pair_microbench.py "SMD_PAIR_ENGINE=0" 2>&1 | tail -1
done
python tools/pair_microbench.py "SMD_PAIR_ENGINE=0" 2>&1 | tail -1
