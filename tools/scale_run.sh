#!/bin/bash
# usage (GPU box with N GPUs, via gpurun --gpus N): tools/scale_run.sh <tag> <n1> <n2> ...   -- bench.py at each GPU count
tag=$1; shift
mkdir -p gpurun_out
port=29540
for n in "$@"; do
  port=$((port+1))
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --workload bilayer --steps 5 --no-cpu-baseline > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_${tag}_n$n.json"))
    w = (d.get("weak_scaling_reference") or {}).get("value")
    print("N=$n value %.4g  per-gpu %.4g  weak-ref(1 gpu tile) %s  e2e %.4g  us/md-step %.1f" % (d["value"], d["value"] / $n, ("%.4g" % w) if w else None, (d.get("e2e") or {}).get("value", 0), d["us_per_md_step"]))
    print("   phases", {k: round(v, 1) for k, v in d["phases_us_per_md_step"].items()}, "box moves", d["box_moves"], "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("N=$n: no bench line:", e)
    import subprocess; print(subprocess.run("tail -5 gpurun_out/scale_${tag}_n$n.err", shell=True, capture_output=True, text=True).stdout)
PY
done
