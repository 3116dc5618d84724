#!/bin/bash
# usage (GPU box): tools/bench_phases.sh [bench args]  -- one short device-resident bench line, phases printed
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g us/step %.1f pair %.1f' % (d['value'], d['us_per_md_step'], d['roofline']['avg_launch_us']), {k: round(v,1) for k,v in d['phases_us_per_md_step'].items()})"
