#!/usr/bin/env python
"""Whole-executable timing of MD_b200 on C2 (liposome, 80 000 lipids) with the reference's default output cadence
(storeInterval = 1000 steps: 26 MB .mpd + xyz frame; measureInterval = 100 steps), asynchronous writer against
SMD_SYNC_IO=1.  usage (GPU box): python tools/e2e_driver.py [md_steps]  -> JSON lines on stdout"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softmold_b200 import workloads
from oracle import orc   # test infrastructure: only its .mpd text writer is used here

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
m = workloads.liposome(80000)
m.update(finalTime=steps * 0.02, storeInterval=20.0, measureInterval=2.0)
exe = os.path.join(ROOT, "softmold_b200", "MD_b200")
for mode in ("async", "sync", "async"):
    with tempfile.TemporaryDirectory() as d:
        orc.write_mpd(os.path.join(d, "c2.mpd"), m)
        env = dict(os.environ, SMD_TIMING="1")
        if mode == "sync":
            env["SMD_SYNC_IO"] = "1"
        t = time.time()
        r = subprocess.run([exe, "c2"], cwd=d, env=env, capture_output=True, text=True)
        wall = time.time() - t
        assert r.returncode == 0, r.stderr[-1000:]
        nfr = open(os.path.join(d, "frames_c2.xyz")).read().count("test\n")
        loop, drained = [float(x) for x in [ln for ln in r.stderr.splitlines() if ln.startswith("SMD_TIMING")][0].split()[2::2]]
        print(json.dumps({"mode": mode, "md_steps": steps, "process_wall_s": round(wall, 3), "loop_s": round(loop, 3), "loop_and_drain_s": round(drained, 3),
                          "particle_steps_per_s_loop_and_drain": m["nParticles"] * steps / drained,
                          "frames": nfr, "measures": len(open(os.path.join(d, "potential_c2.dat")).read().splitlines())}))
