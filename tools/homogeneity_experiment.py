#!/usr/bin/env python
"""How much of the pair kernel's time is the mix of long-range (TAIL) and short-range (HEAD) particles inside a block?
(GPU box)  The equilibrated C2 vesicle, pair force launch timed by CUDA events through smd_profile:
   (a) as it is: 160 000 TAIL particles (cutoff rc) + 80 000 HEAD particles (purely repulsive: cutoff rm), mixed in every block
   (b) the same positions with every particle relabelled TAIL: 240 000 long-range particles, homogeneous blocks
   (c) every particle relabelled HEAD: 240 000 short-range particles, homogeneous blocks
If blocks were homogeneous, the mixed system would cost about (160 000 b + 80 000 c) / 240 000."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import softmold_b200 as sm
from softmold_b200 import workloads

m = workloads.liposome(80000, 3.45, 777)
ctx = sm.Context.from_dict(m)
ctx.compute_forces(step=0)
ctx.step(0, 400)
xyz, typ, vel = ctx.get_particles()
ctx.close()
m = dict(m, xyz=xyz, vel=vel)
HEAD, TAIL = 2, 3
out = {}
for name, t in (("mixed", typ), ("all_tail", np.full_like(typ, TAIL)), ("all_head", np.full_like(typ, HEAD))):
    c = sm.Context.from_dict(dict(m, type=t))
    for _ in range(3):
        c.compute_forces(mask=1 << sm.TERM_PAIR)
    c.profile(["pair"])
    for _ in range(20):
        c.compute_forces(mask=1 << sm.TERM_PAIR)
    ms, cnt = c.profile_read()["pair"]
    out[name] = ms * 1e3 / cnt
    print(f"{name:9s} pair launch {out[name]:7.1f} us   in-range pairs {c.count_pairs(per_particle=False)[0]}")
    c.close()
est = (160000 * out["all_tail"] + 80000 * out["all_head"]) / 240000
print(f"homogeneous-block estimate for the mix: {est:.1f} us against {out['mixed']:.1f} measured ({100 * (out['mixed'] / est - 1):.0f} % above)")
