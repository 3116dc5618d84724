#!/usr/bin/env python
"""Device-side timeline of one steady-state MD step (GPU box).  Needs a library built with -DSMD_TIMELINE:

    nvcc <flags of softmold_b200/csrc/Makefile> -DSMD_TIMELINE -shared -o scratch/tl/libsoftmold_b200.so softmold_b200/csrc/smd_core.cu softmold_b200/csrc/mpd_io.cpp
    SOFTMOLD_B200_LIB=scratch/tl/libsoftmold_b200.so python tools/timeline.py [lipids ...]

Every kernel of the step stamps %globaltimer when its first block starts, when its last block starts and when its last block
ends; printed relative to the first block of k_scan, averaged over REPS steps.  Under programmatic dependent launches the
kernels overlap (a block of the next kernel becomes resident as soon as one of this kernel's leaves), which event pairs
around the launches cannot show -- and which they destroy."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import softmold_b200 as sm
from softmold_b200 import workloads

NAMES = ["k_scan", "k_place", "k_reorder", "k_pair_force2", "k_chain_kick"]
REPS = 20


def run(lipids):
    m = workloads.liposome(lipids, 3.45, 777)
    ctx = sm.Context.from_dict(m)
    L = sm.lib()
    L.smd_timeline_read.argtypes = [C.c_void_p, C.c_void_p]
    ctx.compute_forces(step=0)
    ctx.step(0, 400)
    rows = []
    step = 400
    for _ in range(REPS):
        ctx.step(step, 6)          # the stamps are those of the 5th of the 6 steps
        step += 6
        buf = (C.c_uint64 * 48)()
        assert L.smd_timeline_read(ctx.h, buf) == 0
        t = np.array(buf[:15], dtype=np.float64).reshape(5, 3)
        rows.append((t - t[0, 0]) * 1e-3)
    ctx.close()
    t = np.mean(rows, axis=0)
    print(f"liposome {lipids} lipids = {3 * lipids} particles; us after the first block of k_scan (mean of {REPS} steps)")
    print(f"  {'kernel':16s} {'first block':>12s} {'last block in':>14s} {'last block out':>15s}")
    for k, nm in enumerate(NAMES):
        print(f"  {nm:16s} {t[k, 0]:12.1f} {t[k, 1]:14.1f} {t[k, 2]:15.1f}")
    print(f"  scan -> seam end: {t[4, 2]:.1f} us")


if __name__ == "__main__":
    for a in (sys.argv[1:] or ["5000", "80000"]):
        run(int(a))
