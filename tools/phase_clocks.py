#!/usr/bin/env python
"""Cycles per warp between the phase marks of k_pair_force2 (GPU box).  Needs a library built with -DSMD_PHASE_CLOCKS:
    nvcc <flags of softmold_b200/csrc/Makefile> -DSMD_PHASE_CLOCKS -shared -o scratch/pc/libsoftmold_b200.so softmold_b200/csrc/smd_core.cu softmold_b200/csrc/mpd_io.cpp
    SOFTMOLD_B200_LIB=scratch/pc/libsoftmold_b200.so python tools/phase_clocks.py
The equilibrated C2 vesicle as it is, with every particle relabelled TAIL (long range), and with every particle relabelled HEAD
(short range): where the part of the launch that does not depend on the candidate count sits."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import softmold_b200 as sm
from softmold_b200 import workloads

NAMES = ["entry (tables, dealing, own record)", "range set-up", "phase 1", "periodic-image rows", "hand-over (3 barriers)", "phase 2 + epilogue"]
m = workloads.liposome(80000, 3.45, 777)
ctx = sm.Context.from_dict(m)
ctx.compute_forces(step=0)
ctx.step(0, 400)
xyz, typ, vel = ctx.get_particles()
ctx.close()
m = dict(m, xyz=xyz, vel=vel)
L = sm.lib()
L.smd_phase_clocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for name, t in (("mixed", typ), ("all TAIL", np.full_like(typ, 3)), ("all HEAD", np.full_like(typ, 2))):
    c = sm.Context.from_dict(dict(m, type=t))
    for _ in range(3):
        c.compute_forces(mask=1 << sm.TERM_PAIR)
    buf = (C.c_uint64 * 16)()
    L.smd_phase_clocks(c.h, buf, 1)
    reps = 10
    c.profile(["pair"])
    for _ in range(reps):
        c.compute_forces(mask=1 << sm.TERM_PAIR)
    ms, cnt = c.profile_read()["pair"]
    L.smd_phase_clocks(c.h, buf, 1)
    v = np.array(buf[:6], dtype=np.float64)
    nwarps = buf[15] * 4 / 1.0
    print(f"{name}: launch {ms * 1e3 / cnt:.1f} us; cycles per warp (mean over {int(nwarps)} warps):")
    for k, nm in enumerate(NAMES):
        print(f"    {nm:45s} {v[k] / nwarps:9.0f}  {100 * v[k] / v.sum():5.1f} %")
    print(f"    {'total':45s} {v.sum() / nwarps:9.0f}")
    c.close()
