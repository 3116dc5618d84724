#!/bin/bash
# usage (GPU box): tools/ab_libs.sh lib1.so lib2.so ...   -- same bench, different builds of the library (A/B of kernel variants)
for lib in "$@"; do
  SOFTMOLD_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${AB_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'us/step %.1f' % d['us_per_md_step'], 'pair %.1f' % d['roofline']['avg_launch_us'], {k: round(v,1) for k,v in d['phases_us_per_md_step'].items()})"
done
