#!/usr/bin/env python
"""North-star item (1), "rebuilt only past a skin-distance trigger": the measurement behind profiles/r02c_skin_sweep.md.

  tools/skin_experiment.py cadence [liposome|bilayer]
      the product library steps the equilibrated workload one MD step at a time; a max-displacement trigger is replayed on
      the unwrapped positions for skins of 0.2 / 0.3 / 0.5 / 1.0 sigma (rebuild when the largest displacement since the last
      rebuild exceeds skin / 2) -> average number of steps a list survives.  Also writes the equilibrated state for `lists`.
  SMD_SKIN=s SMD_PAIR_SPLIT=1 tools/skin_experiment.py lists <root of the experiment tree>
      the two-kernel engine of experiments/smd_pair_split.cuh (commit 677b197 + an SMD_SKIN hook that widens the phase-1
      cutoffs, built under scratch/): k_pair_lists writes the per-particle candidate lists to global memory, k_pair_drain
      evaluates them with the exact FP64 test.  Run under `ncu --metrics gpu__time_duration.sum`: the drain alone is what a
      reuse step of a skin list would cost, lists + drain what a rebuild step would.
"""
import json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATE = os.path.join(ROOT, "gpurun_out", "skin_state_%s.npz")


def workload(name):
    sys.path.insert(0, ROOT)
    from softmold_b200 import workloads
    return workloads.liposome(80000, 3.45, 777) if name == "liposome" else workloads.bilayer(332928, 3.11, seed=5)


def cadence(name):
    sys.path.insert(0, ROOT)
    import softmold_b200 as sm
    m = workload(name)
    ctx = sm.Context.from_dict(m, track_unwrapped=True)
    ctx.compute_forces(step=0)
    ctx.step(0, 400)
    xyz, typ, vel = ctx.get_particles()
    os.makedirs(os.path.dirname(STATE), exist_ok=True)
    np.savez(STATE % name, xyz=xyz, vel=vel)
    skins = (0.2, 0.3, 0.5, 1.0)
    ref = {s: ctx.get_unwrapped().copy() for s in skins}
    rebuilds = {s: 0 for s in skins}
    per_step = []
    nsteps = 200
    prev = ref[skins[0]].copy()
    for k in range(nsteps):
        ctx.step(400 + k, 1)
        u = ctx.get_unwrapped()
        per_step.append(float(np.sqrt(((u - prev) ** 2).sum(axis=1)).max()))
        prev = u.copy()
        for s in skins:
            d = np.sqrt(((u - ref[s]) ** 2).sum(axis=1)).max()
            if d > 0.5 * s:
                rebuilds[s] += 1
                ref[s] = u.copy()
    ctx.close()
    out = {"workload": name, "particles": int(len(xyz)), "steps": nsteps,
           "max_displacement_per_step_sigma": {"mean": float(np.mean(per_step)), "max": float(np.max(per_step))},
           "steps_per_rebuild": {str(s): (nsteps / rebuilds[s] if rebuilds[s] else None) for s in skins}}
    print(json.dumps(out))


def lists(root):
    sys.path.insert(0, root)
    import softmold_b200 as sm
    assert os.path.realpath(os.path.dirname(sm.__file__)).startswith(os.path.realpath(root))
    m = dict(workload("liposome"))
    st = np.load(STATE % "liposome")
    m["xyz"], m["vel"] = st["xyz"], st["vel"]
    ctx = sm.Context.from_dict(m)
    for _ in range(4):
        ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    ctx.synchronize()
    tot = ctx.count_pairs()[0] if hasattr(ctx, "count_pairs") else -1
    print(json.dumps({"skin": float(os.environ.get("SMD_SKIN", "0")), "split": os.environ.get("SMD_PAIR_SPLIT", "0"), "pairs_in_range": int(tot)}))
    ctx.close()


if __name__ == "__main__":
    if sys.argv[1] == "cadence":
        cadence(sys.argv[2] if len(sys.argv) > 2 else "liposome")
    else:
        lists(os.path.abspath(sys.argv[2]))
