#!/bin/bash
# usage (GPU box): tools/ab_env.sh "VAR=a" "VAR=b" ...   -- same bench, same library, different environment settings
for kv in "$@"; do
  env $kv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${AB_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$kv', 'us/step %.1f' % d['us_per_md_step'], 'pair %.1f' % d['roofline']['avg_launch_us'], {k: round(v,1) for k,v in d['phases_us_per_md_step'].items()})"
done
