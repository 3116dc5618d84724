#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_check.sh <tag> [bench args]
# runs the GPU parity tests and one bench line; outputs land in gpurun_out/
tag=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -20 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print("value %.4g  e2e %.4g  us/md-step %.1f  cpu %s" % (d["value"], (d.get("e2e") or {}).get("value", 0), d["us_per_md_step"], (d.get("cpu_baseline") or {}).get("value")))
    print("phases", {k: round(v, 1) for k, v in d["phases_us_per_md_step"].items()})
    print("roofline", d["roofline"] and {k: d["roofline"][k] for k in ("avg_launch_us", "share_of_step", "frac")}, d["roofline"] and d["roofline"]["fp64"]["frac_of_mul_add_peak"])
    print("clocks", d["clocks"], "launches", d["gpu_launches"])
except Exception as e:
    print("no bench line:", e)
PY
