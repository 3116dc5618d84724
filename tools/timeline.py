#!/usr/bin/env python
"""Device-side timeline of one steady-state MD step (GPU box):  python tools/timeline.py [lipids ...]

smd_timeline: every kernel of the step stamps %globaltimer when its first block starts working (after its
griddepcontrol.wait), when its last block starts and when its last block ends; printed relative to the first block of k_scan,
averaged over REPS steps.  Under programmatic dependent launches the kernels overlap (a block of the next kernel becomes resident
as soon as one of this kernel's leaves), which event pairs around the launches cannot show -- and which they destroy."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import softmold_b200 as sm
from softmold_b200 import workloads

REPS = 20


def run(lipids):
    m = workloads.liposome(lipids, 3.45, 777)
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces(step=0)
    ctx.step(0, 400)
    ctx.timeline(True)
    rows, step = [], 400
    for _ in range(REPS):
        ctx.step(step, 6)          # the stamps are those of the 5th of the 6 steps
        step += 6
        rows.append(ctx.timeline_read())
    ctx.close()
    print(f"liposome {lipids} lipids = {3 * lipids} particles; us after the first block of k_scan (mean of {REPS} steps)")
    print(f"  {'kernel':16s} {'first block':>12s} {'last block in':>14s} {'last block out':>15s}")
    for nm in ctx.TIMELINE_KERNELS:
        t = np.mean([r[nm] for r in rows], axis=0)
        print(f"  {nm:16s} {t[0]:12.1f} {t[1]:14.1f} {t[2]:15.1f}")


if __name__ == "__main__":
    for a in (sys.argv[1:] or ["5000", "80000"]):
        run(int(a))
