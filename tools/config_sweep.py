#!/usr/bin/env python
"""The BASELINE.json configurations C1 - C4 through the two executables, side by side: inputs made by the reference's own
generators (prebuilt under oracle/_ref/ -- they travel to the GPU box with the repo), equilibrated by MD_b200, then the
steady-state loop of `MD_b200 <name>` against the unmodified `MD <name>` (all host cores) on the same .mpd with store and
measure intervals pushed past the end ("exclude file load ... no I/O is timed").

usage (GPU box): python tools/config_sweep.py [--steps N] [--ref-steps M] [--configs C1,C3,...]   -> JSON lines

Per configuration: N, molecules, particle-steps/s of both loops (MD_b200: the SMD_TIMING line of its main loop; MD: the
wall time of a run minus the wall time of a 0-step run, i.e. without load + initial forces + final store), the ratio, and
the t = 0 potential both executables print for the same file (6 digits) as a parity spot check."""
import argparse, json, os, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc   # test infrastructure: .mpd text reader / writer only

REF = os.path.join(ROOT, "oracle", "_ref")
MD_B200 = os.path.join(ROOT, "softmold_b200", "MD_b200")

CONFIGS = {
    # name: (generator, args after the name)                                                         SURVEY 8d
    "C1": ("liposome", ["5000", "5000", "3.45"]),
    "C2": ("liposome", ["777", "80000", "3.45"]),
    "C3": ("continuumSphereAndLiposome", ["1234", "5000", "3.45", "4", "-6", "40", "5.88", "0", "1", "0", "2.0"]),
    "C4": ("lipoCyto", ["4321", "-6", "0", "20000", "3.45", "0", "10", "2"]),
}


def run(cmd, cwd, env=None, timeout=3600):
    t = time.time()
    r = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    return r, time.time() - t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000, help="MD steps of the timed MD_b200 run")
    ap.add_argument("--ref-steps", type=int, default=40, help="MD steps of the timed reference run")
    ap.add_argument("--equil", type=int, default=500)
    ap.add_argument("--configs", default="C1,C2,C3,C4")
    ap.add_argument("--tension", action="store_true", help="add deltaLXY 0.01 + tension 0.1 (box moves every 8 steps)")
    a = ap.parse_args()
    cores = os.cpu_count()
    for name in a.configs.split(","):
        gen, args = CONFIGS[name]
        with tempfile.TemporaryDirectory() as d:
            r, _ = run([os.path.join(REF, gen), "in"] + args, d)
            assert r.returncode == 0 and os.path.exists(os.path.join(d, "in.mpd")), (gen, r.stderr[-500:])
            m = orc.read_mpd(os.path.join(d, "in.mpd"))
            if a.tension:
                m["deltaLXY"], m["tension"] = 0.01, 0.1
            dt = m["deltaT"]
            # equilibrate with MD_b200 (thermal cell occupancy instead of the generator's lattice), checkpoint = the input
            m.update(initialTime=0.0, finalTime=a.equil * dt, storeInterval=1e9, measureInterval=1e9)
            orc.write_mpd(os.path.join(d, "eq.mpd"), m)
            r, _ = run([MD_B200, "eq"], d)
            assert r.returncode == 0, r.stderr[-1000:]
            eq = orc.read_mpd(os.path.join(d, "eq.mpd"))
            out = {"config": name, "generator": gen + " " + " ".join(args), "n_particles": eq["nParticles"],
                   "molecules": [[int(mol["type"]), int(len(mol["bonds"]))] for mol in eq["molecules"]],
                   "box_moves": bool(a.tension), "host_cores": cores}
            t0 = eq["initialTime"]

            def job(sub, exe, steps, env=None):
                dd = os.path.join(d, sub)
                os.makedirs(dd)
                mm = dict(eq, finalTime=t0 + steps * dt, storeInterval=1e9, measureInterval=1e9)
                orc.write_mpd(os.path.join(dd, "c.mpd"), mm)
                r, wall = run([exe, "c"], dd, env=env)
                assert r.returncode == 0, (exe, r.stderr[-1000:])
                return r, wall, dd

            r, wall, _ = job("ours", MD_B200, a.steps, dict(os.environ, SMD_TIMING="1"))
            loop = float([ln for ln in r.stderr.splitlines() if ln.startswith("SMD_TIMING")][0].split()[2])
            out["md_b200"] = {"md_steps": a.steps, "loop_s": round(loop, 4), "process_wall_s": round(wall, 3),
                              "particle_steps_per_s": eq["nParticles"] * a.steps / loop}
            if os.path.exists(os.path.join(REF, "MD")):
                env = dict(os.environ, OMP_NUM_THREADS=str(cores))
                _, w0, _ = job("ref0", os.path.join(REF, "MD"), 0, env)
                _, w1, _ = job("ref1", os.path.join(REF, "MD"), a.ref_steps, env)
                loop_ref = max(w1 - w0, 1e-9)
                out["reference_md"] = {"md_steps": a.ref_steps, "loop_s": round(loop_ref, 3), "threads": cores,
                                       "particle_steps_per_s": eq["nParticles"] * a.ref_steps / loop_ref}
                out["ratio"] = out["md_b200"]["particle_steps_per_s"] / out["reference_md"]["particle_steps_per_s"]
                # parity spot check through the files: both executables measure the same configuration at restart + 1 step
                pots = {}
                for tag, exe in (("ours", MD_B200), ("ref", os.path.join(REF, "MD"))):
                    dd = os.path.join(d, "p" + tag)
                    os.makedirs(dd)
                    mm = dict(eq, initialTime=0.0, finalTime=0.0, storeInterval=1e9, measureInterval=1e9)
                    orc.write_mpd(os.path.join(dd, "c.mpd"), mm)
                    r, _ = run([exe, "c"], dd, env=env)
                    assert r.returncode == 0, (exe, r.stderr[-500:])
                    pots[tag] = float(open(os.path.join(dd, "potential_c.dat")).read().split()[1])
                out["potential_t0"] = pots
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
