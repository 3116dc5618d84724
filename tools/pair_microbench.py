#!/usr/bin/env python
"""micro-benchmark of the pair force launch alone on the equilibrated C2 vesicle (GPU box):
   tools/pair_microbench.py "ENV=a ENV2=b" "ENV=c" ...   each argument = one environment for a fresh context"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import softmold_b200 as sm
from softmold_b200 import workloads

lip = int(os.environ.get("MB_LIPIDS", "80000"))
m = workloads.liposome(lip, 3.45, 777)
ctx = sm.Context.from_dict(m)
ctx.compute_forces(step=0)
ctx.step(0, int(os.environ.get("MB_EQUIL", "400")))
xyz, typ, vel = ctx.get_particles()
ctx.close()
m = dict(m); m["xyz"], m["vel"] = xyz, vel
for arg in sys.argv[1:]:
    env = dict(kv.split("=", 1) for kv in arg.split())
    os.environ.update(env)
    try:
        c = sm.Context.from_dict(m)
        for _ in range(3):
            c.compute_forces(mask=1 << sm.TERM_PAIR)
        c.synchronize()
        n = 30
        t0 = time.perf_counter()
        for _ in range(n):
            c.compute_forces(mask=1 << sm.TERM_PAIR)
        c.synchronize()
        dt = (time.perf_counter() - t0) / n
        print("%-50s %.1f us per launch (zero + pair)" % (arg, dt * 1e6), flush=True)
        c.close()
    finally:
        for k in env:
            del os.environ[k]
