#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python oracle/make_golden.py

Every fixture holds an input configuration (as the reference's generators / its own `MD` wrote it) and the
reference's outputs on it:
  * per-term accelerations / potentials / dPotentials, cell ids and linked-list order from
    oracle/_ref/ref_harness (which calls the reference headers: CellOpt, Blob::do*Force, ...),
  * for the `traj_*` entries the state written by the reference `MD` executable itself
    (OMP_NUM_THREADS=1, so the Langevin noise is the single MT19937 stream MTRand(seed)) after K steps.
The reference ships no golden vectors of its own; these are "outputs of the reference
run here".
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import orc  # noqa: E402

REF = os.path.join(HERE, "_ref")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ENV1 = dict(os.environ, OMP_NUM_THREADS="1")


def run(cmd, cwd):
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=ENV1)


def run_md(cwd, m, name, nsteps):
    """run the reference MD for nsteps from m (dict); returns the dict it stored at its last step.
    The reference stores inside iteration endInt, after Verlet::first and before the new forces
    (MD.cpp:373-381), so the returned state is {x after first(), v at the half kick}."""
    m = dict(m)
    t0 = m["initialTime"]
    start = int(t0 / m["deltaT"] + 1e-7)
    end = start + nsteps
    m["finalTime"] = end * m["deltaT"]
    m["measureInterval"] = 1e6
    # store only at the final step: i % storeint == 0 for i in (start, end] only at i == end when storeint == end
    m["storeInterval"] = end * m["deltaT"]
    orc.write_mpd(os.path.join(cwd, name + ".mpd"), m)
    run([os.path.join(REF, "MD"), name], cwd)
    return orc.read_mpd(os.path.join(cwd, name + ".mpd"))


def harness(cwd, name, scale=None, substrate=False):
    cmd = [os.path.join(REF, "ref_harness"), "dump", name, name + ".bin"]
    if scale is not None:
        cmd += [repr(float(s)) for s in scale]
    if substrate:
        cmd += ["substrate"]     # (after an explicit scale) the molecule switch of MDsubstrate.cpp instead of MD.cpp's
    run(cmd, cwd)
    return orc.read_dump(os.path.join(cwd, name + ".bin"))


def pack(m, g, extra=None):
    d = {"nTypes": m["nTypes"], "box": np.array(m["size"]), "cutoff": m["cutoff"], "deltaT": m["deltaT"],
         "gamma": m["gamma"], "temperature": m["initialTemp"], "seed": m["seed"], "initialTime": m["initialTime"],
         "deltaLXY": m.get("deltaLXY", 0.0), "tension": m.get("tension", 0.0),
         "xyz": m["xyz"], "vel": m["vel"], "type": m["type"], "fC": m["twoBodyFconst"], "uC": m["twoBodyUconst"],
         "nmol": len(m["molecules"])}
    for k, mol in enumerate(m["molecules"]):
        d[f"mol{k}_type"] = mol["type"]
        d[f"mol{k}_bonds"] = mol["bonds"]
        d[f"mol{k}_const"] = mol["constants"]
    for key in ("scale", "cell_id", "cell_next", "nCells_nFull", "fullCells", "a_pair", "U_pair", "dU_pair",
                "U_mol", "dU_mol", "kinetic"):
        if key in g:
            d["ref_" + key] = g[key]
    for k in range(len(m["molecules"])):
        d[f"ref_a_mol{k}"] = g[f"a_mol{k}"]
    if extra:
        d.update(extra)
    return d


def traj(cwd, m, name, nsteps):
    f = run_md(cwd, m, name, nsteps)
    return {"traj_steps": nsteps, "traj_xyz": f["xyz"], "traj_vel": f["vel"], "traj_box": np.array(f["size"]),
            "traj_time": f["initialTime"]}, f


def save(name, d):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: N={len(d['xyz'])} nmol={d['nmol']} -> {os.path.getsize(os.path.join(OUT, name + '.npz')) // 1024} KiB")


def main():
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        # ---- liposome, 300 lipids (CHAIN only; box 400^3 -> 8e6 cells, 0.01% occupied)
        run([os.path.join(REF, "liposome"), "lipo", "42", "300", "3.45"], tmp)
        m = orc.read_mpd(os.path.join(tmp, "lipo.mpd"))
        t, eq = traj(tmp, m, "lipo_run", 24)
        save("lipo_t0", pack(m, harness(tmp, "lipo"), t))
        _, eq = traj(tmp, m, "lipo_run", 150)
        orc.write_mpd(os.path.join(tmp, "lipo_eq.mpd"), eq)
        t, _ = traj(tmp, eq, "lipo_run2", 16)  # restart path: initialTime != 0 (MD.cpp:274-308)
        save("lipo_eq", pack(eq, harness(tmp, "lipo_eq", (0.9993, 0.9993, 1.0 / 0.9993 ** 2)), t))

        # ---- explicit BOND + BEND lists equivalent to the CHAIN (list code paths, system.h:1880-2040)
        mb = dict(eq)
        ch = eq["molecules"][0]
        start, nch, ln = [int(x) for x in ch["bonds"][0]]
        bonds = np.array([[s + l, s + l + 1] for s in range(start, start + nch * ln, ln) for l in range(ln - 1)], np.int32)
        bends = np.array([[s + l, s + l + 1, s + l + 2] for s in range(start, start + nch * ln, ln) for l in range(ln - 2)], np.int32)
        mb["molecules"] = [{"type": orc.BOND, "constants": ch["constants"][:2], "bonds": bonds},
                           {"type": orc.BEND, "constants": ch["constants"][2:], "bonds": bends}]
        mb["nMolecules"] = 2
        orc.write_mpd(os.path.join(tmp, "bondbend.mpd"), mb)
        save("bondbend", pack(mb, harness(tmp, "bondbend")))

        # ---- flat bilayer, fully periodic small box (pairs across the boundary), MC box moves with tension
        run([os.path.join(REF, "bilayer"), "bl", "99", "600", "3.11", "0", "0", "0", "0"], tmp)
        m = orc.read_mpd(os.path.join(tmp, "bl.mpd"))
        m["tension"] = 0.5
        orc.write_mpd(os.path.join(tmp, "bl.mpd"), m)
        t, _ = traj(tmp, m, "bl_run", 24)  # MC trials at i = 8, 16, 24
        save("bilayer_t0", pack(m, harness(tmp, "bl"), t))
        _, eq = traj(tmp, m, "bl_run", 200)
        orc.write_mpd(os.path.join(tmp, "bl_eq.mpd"), eq)
        t, _ = traj(tmp, eq, "bl_run2", 16)
        save("bilayer_eq", pack(eq, harness(tmp, "bl_eq", (1.0004, 1.0004, 1.0 / 1.0004 ** 2)), t))

        # ---- liposome + cytoskeleton: CHAIN(3) + CHAIN(6) + BOND anchors
        run([os.path.join(REF, "lipoCyto"), "lc", "4321", "-6", "0", "800", "3.45", "0", "6", "1"], tmp)
        m = orc.read_mpd(os.path.join(tmp, "lc.mpd"))
        _, eq = traj(tmp, m, "lc_run", 100)
        orc.write_mpd(os.path.join(tmp, "lc_eq.mpd"), eq)
        save("lipocyto_eq", pack(eq, harness(tmp, "lc_eq")))

        # ---- vesicle + continuum-sphere bead (BEAD molecule, 8 types).  The generator starts the bead out of range
        # (SURVEY 8c), so pull it onto the outer leaflet.
        run([os.path.join(REF, "continuumSphereAndLiposome"), "cs", "1234", "1200", "3.45", "3", "-6", "40", "5.88",
             "0", "1", "0", "2.0"], tmp)
        m = orc.read_mpd(os.path.join(tmp, "cs.mpd"))
        b = int(m["molecules"][1]["bonds"][0, 0])
        m["xyz"][b, 2] -= 2.2
        orc.write_mpd(os.path.join(tmp, "bead1.mpd"), m)
        t, _ = traj(tmp, m, "bead1_run", 16)  # covers the double mass division quirk (MD.cpp:340-355, :480-494)
        save("bead1", pack(m, harness(tmp, "bead1"), t))

        # two beads in one BEAD molecule, close to each other and to the membrane (bead-bead term + exclusion quirk Q7)
        m2 = dict(m)
        m2["xyz"] = np.vstack([m["xyz"], m["xyz"][b] + np.array([4.6, 0.3, -0.4])])
        m2["vel"] = np.vstack([m["vel"], [0.1, -0.2, 0.05]])
        m2["type"] = np.append(m["type"], m["type"][b]).astype(np.int32)
        m2["nParticles"] = m["nParticles"] + 1
        m2["molecules"] = [m["molecules"][0], {"type": orc.BEAD, "constants": m["molecules"][1]["constants"],
                                               "bonds": np.array([[b], [b + 1]], np.int32)}]
        orc.write_mpd(os.path.join(tmp, "bead2.mpd"), m2)
        t, _ = traj(tmp, m2, "bead2_run", 8)
        save("bead2", pack(m2, harness(tmp, "bead2"), t))

        # ---- MT19937 known answers (MersenneTwister.h) for the barostat / reference-noise streams
        run([os.path.join(REF, "ref_harness"), "mt", "5000", "64", "mt.bin"], tmp)
        g = orc.read_dump(os.path.join(tmp, "mt.bin"))
        run([os.path.join(REF, "ref_harness"), "mt", "42", "64", "mt2.bin"], tmp)
        g2 = orc.read_dump(os.path.join(tmp, "mt2.bin"))
        np.savez_compressed(os.path.join(OUT, "mt19937.npz"), seed5000_rand53=g["rand53"],
                            seed5000_u32=g["randInt"].view(np.uint32), seed42_rand53=g2["rand53"],
                            seed42_u32=g2["randInt"].view(np.uint32))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ball():
    """BALL molecule (system.h:1936-1971: a half-harmonic wall around a centre particle; no shipped generator emits
    one, `MD` handles it, MD.cpp:251,467,662): the equilibrated 300-lipid vesicle of lipo_eq with two balls in one
    record list -- the records of the second centre exercise the reference's quirk of taking the centre of every
    record from record 0 in the force routine only."""
    tmp = tempfile.mkdtemp(prefix="golden_ball_")
    try:
        m, _ = orc.load_golden(os.path.join(OUT, "lipo_eq.npz"))
        m = dict(m)
        n = m["nParticles"]
        heads = [i for i in range(n) if m["type"][i] == 2]
        c1, c2 = 1, 451   # two tail particles
        rec = [[c1, j] for j in heads[0::2]] + [[c2, j] for j in heads[1::2]]
        m["molecules"] = list(m["molecules"]) + [{"type": orc.BALL, "constants": np.array([5.0, 20.0]),
                                                  "bonds": np.array(rec, np.int32)}]
        m["nMolecules"] = len(m["molecules"])
        orc.write_mpd(os.path.join(tmp, "ball.mpd"), m)
        t, _ = traj(tmp, m, "ball_run", 16)
        save("ball", pack(m, harness(tmp, "ball", (0.9993, 0.9993, 1.0 / 0.9993 ** 2)), t))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def spangler_row(rc, rm, rad, density, Umax, Umin):
    """one 22-constant bead row (potentials/laradjiSpangler.h:186-223, laradjiSpanglerFCrow)"""
    D = density * np.pi * rad
    A = Umax - Umin
    out = [rc + rad, -7.0 / 4.0 * rm, 2.0 * rm * rm, Umin * np.pi * rad * density / (rm * rm * rm), rad, rm,
           -D * A / (2.0 * rm * rm), 2.0 * D * A / (3.0 * rm), -D * Umin, 2.0 * D * Umin * rm, D * 1.3 * Umin * rm * rm]
    D = (2.0 * np.pi * density * rad) ** 2.0 / (rm * rm * rm)
    out += [2.0 * rad + rm, 2.0 * rad + rc, Umin * D * 2.0 / 30.0, -Umin * D * 7.0 * rm / 20.0, Umin * D * rm * rm / 2.0,
            -D * A * rm / 20.0, D * A * rm * rm / 12.0, -D * Umin * rm ** 3.0 / 6.0, D * Umin * rm ** 4.0 / 2.0,
            D * 1.3 * Umin * rm ** 5.0 / 2.0, D * 13.0 * Umin * rm ** 6.0 / 60.0]
    return out


def fields():
    """The remaining molecule kinds of MD.cpp's switch (MD.cpp:414-478): BOUNDARY, FLOATING_BASE, ZTORQUE, ZPOWERPOTENTIAL,
    NANOCORE (+ a BALL, + the kinds MD parses and ignores: SOLID, OFFSET_BOUNDARY), on the equilibrated periodic flat
    bilayer of bilayer_eq WITH Metropolis box moves, so that the 24-step trajectory of the reference `MD` also pins which
    terms take part in the box move (none of these: MD.cpp:642-669 drops the NANOCORE and BALL results)."""
    tmp = tempfile.mkdtemp(prefix="golden_fields_")
    try:
        m, _ = orc.load_golden(os.path.join(OUT, "bilayer_eq.npz"))
        m = dict(m)
        n, nT = m["nParticles"], m["nTypes"]
        xyz, typ = m["xyz"], m["type"]
        z = xyz[:, 2]
        zmid = float(np.median(z))
        ch = m["molecules"][0]
        # two nanocore particles (type 1) above and below the bilayer, within range of the head groups
        top, bot = float(z.max()), float(z.min())
        cx, cy = m["size"][0] / 2, m["size"][1] / 2
        extra = np.array([[cx, cy, top + 3.45], [cx * 0.4, cy * 1.3, bot - 4.4]])   # gap R + 1.4: the attractive tail
        m["xyz"] = np.vstack([xyz, extra])
        m["vel"] = np.vstack([m["vel"], [[0.05, -0.1, 0.02], [-0.03, 0.04, 0.01]]])
        m["type"] = np.append(typ, [1, 1]).astype(np.int32)
        m["nParticles"] = n + 2
        z = m["xyz"][:, 2]
        heads = np.where(m["type"][:n] == 2)[0]
        low = heads[np.argsort(z[heads])[:120]]          # lowest head groups
        wall = float(z[low].min()) - 1.05                 # |d| from 1.05 upwards: some inside sqrt(2), none near 0
        fb = np.zeros((nT, 6))
        for t in range(nT):
            fb[t] = [zmid - 0.4 + 0.1 * t, zmid + 1.6 + 0.05 * t, 0.0002 * (t + 1), 1e-6 * (t + 1), 0.3, 0.0015 * (t + 1)]
        mols = list(m["molecules"])
        mols.append({"type": orc.BOUNDARY, "constants": np.array([2.0, wall, 0.0, 0.35]), "bonds": low.reshape(-1, 1).astype(np.int32)})
        mols.append({"type": orc.FLOATING_BASE, "constants": fb.ravel(), "bonds": np.arange(1, n, 3, dtype=np.int32).reshape(-1, 1)})
        mols.append({"type": orc.ZTORQUE, "constants": np.array([0.8, 0.5, 0.7, zmid + 0.3]), "bonds": ch["bonds"].astype(np.int32)})
        mols.append({"type": orc.ZPOWER, "constants": np.array([-3.0 * 0.000085 / 4.0, 2.1]), "bonds": np.array([[0, 150], [300, 77]], np.int32)})
        mols.append({"type": orc.NANOCORE, "constants": np.array(spangler_row(2.0, 1.0, 2.0, 5.88, 100.0, -0.5) +
                                                                  spangler_row(2.0, 1.0, 3.0, 5.88, 100.0, -0.3)),
                     "bonds": np.array([[n], [n + 1]], np.int32)})
        mols.append({"type": orc.BALL, "constants": np.array([4.0, 15.0]), "bonds": np.array([[n, int(j)] for j in heads[::7]], np.int32)})
        mols.append({"type": orc.SOLID, "constants": np.array([1.0]), "bonds": np.array([[3], [4]], np.int32)})
        mols.append({"type": orc.OFFSET_BOUNDARY, "constants": np.array([2.0, wall, 1.0, 0.35]), "bonds": low[:10].reshape(-1, 1).astype(np.int32)})
        m["molecules"] = mols
        m["nMolecules"] = len(mols)
        orc.write_mpd(os.path.join(tmp, "fields.mpd"), m)
        t, _ = traj(tmp, m, "fields_run", 24)   # restart path (initialTime != 0) + MC trials
        g = harness(tmp, "fields", (1.0004, 1.0004, 1.0 / 1.0004 ** 2))
        for k, mol in enumerate(mols):
            print(k, mol["type"], "U", g["U_mol"][k], "dU", g["dU_mol"][k], "max|a|", np.abs(g[f"a_mol{k}"]).max())
        save("fields", pack(m, g, t))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def substrate():
    """The molecule kinds only MDsubstrate.cpp's switch evaluates (MDsubstrate.cpp:213-262): OFFSET_BOUNDARY, RIGIDBEND,
    PULLBEAD, next to CHAIN and BOUNDARY, on the equilibrated periodic bilayer of bilayer_eq.  MDsubstrate.cpp does not
    compile in the reference tree (it includes include/fileFormats/vmdOutput.h, which is not there), so there is no
    trajectory: the per-molecule forces come from the unmodified Blob::do*Force members through ref_harness."""
    tmp = tempfile.mkdtemp(prefix="golden_substrate_")
    try:
        m, _ = orc.load_golden(os.path.join(OUT, "bilayer_eq.npz"))
        m = dict(m)
        n = m["nParticles"]
        xyz, typ = m["xyz"], m["type"]
        z = xyz[:, 2]
        heads = np.where(typ == 2)[0]
        low = heads[np.argsort(z[heads])[:150]]
        wall = float(z[low].min()) - 2.05                 # |d| - offset from 1.05 upwards: some inside sqrt(2), none near 0
        ch = m["molecules"][0]
        st, nch, ln = [int(v) for v in ch["bonds"][0]]
        first = st + ln * np.arange(0, nch, 3)            # head and tail end of every third lipid
        pairs = np.stack([first, first + ln - 1], axis=1).astype(np.int32)
        mols = list(m["molecules"])
        mols.append({"type": orc.BOUNDARY, "constants": np.array([2.0, wall + 1.0, 0.0, 0.35]), "bonds": low[:40].reshape(-1, 1).astype(np.int32)})
        mols.append({"type": orc.OFFSET_BOUNDARY, "constants": np.array([2.0, wall, 1.0, 0.35]), "bonds": low.reshape(-1, 1).astype(np.int32)})
        # preferred direction z, onset angle 0.6 rad: lipids of both leaflets, some below and some beyond the onset
        mols.append({"type": orc.RIGIDBEND, "constants": np.array([0.0, 0.0, 1.0, 2.5, 0.6]), "bonds": pairs})
        mols.append({"type": orc.RIGIDBEND, "constants": np.array([0.6, 0.0, 0.8, 1.5, -0.4]), "bonds": pairs[::5]})
        mols.append({"type": orc.PULLBEAD, "constants": np.array([m["size"][0] * 0.95, 2.0, float(np.median(z)) + 4.0, 0.75]),
                     "bonds": np.array([[int(heads[5])], [int(heads[77])]], np.int32)})
        # kinds MDsubstrate ignores
        mols.append({"type": orc.ZPOWER, "constants": np.array([-0.0001, 2.1]), "bonds": np.array([[0, 150]], np.int32)})
        mols.append({"type": orc.BALL, "constants": np.array([4.0, 15.0]), "bonds": np.array([[0, int(j)] for j in heads[:9]], np.int32)})
        m["molecules"] = mols
        m["nMolecules"] = len(mols)
        orc.write_mpd(os.path.join(tmp, "substrate.mpd"), m)
        g = harness(tmp, "substrate", (1.0004, 1.0004, 1.0 / 1.0004 ** 2), substrate=True)
        for k, mol in enumerate(mols):
            print(k, mol["type"], "U", g["U_mol"][k], "dU", g["dU_mol"][k], "max|a|", np.abs(g[f"a_mol{k}"]).max())
        save("substrate", pack(m, g))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def stats():
    """long-run observables of the reference `MD` executable itself: a tensionless flat bilayer with box moves, 20 000
    steps, two independent runs (seeds 99 / 100, 8 OpenMP threads).  tests/golden/stat_bilayer.npz holds the input
    configuration and the reference's own kinetic_ / potential_ / size_ time series; the GPU test runs MD_b200 on the
    same input and compares ensemble means (temperature, potential energy per particle, area per lipid)."""
    tmp = tempfile.mkdtemp(prefix="golden_stat_")
    env8 = dict(os.environ, OMP_NUM_THREADS="8")
    try:
        subprocess.run([os.path.join(REF, "bilayer"), "sb", "99", "600", "3.11", "0", "0", "0", "0"], cwd=tmp, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        m = orc.read_mpd(os.path.join(tmp, "sb.mpd"))
        m.update(finalTime=400.0, storeInterval=400.0, measureInterval=1.0)
        out = {}
        for tag, seed in (("a", 99), ("b", 100)):
            d = os.path.join(tmp, tag)
            os.makedirs(d)
            mm = dict(m, seed=seed)
            orc.write_mpd(os.path.join(d, "sb.mpd"), mm)
            subprocess.run([os.path.join(REF, "MD"), "sb"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env8)
            for nm in ("kinetic", "potential", "size"):
                out[f"{nm}_{tag}"] = np.loadtxt(os.path.join(d, f"{nm}_sb.dat"))
        mm = dict(m, seed=99)
        orc.write_mpd(os.path.join(tmp, "in.mpd"), mm)
        out["mpd_text"] = np.frombuffer(open(os.path.join(tmp, "in.mpd"), "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, "stat_bilayer.npz"), **out)
        n = m["nParticles"]
        for tag in "ab":
            k, u, sz = out[f"kinetic_{tag}"], out[f"potential_{tag}"], out[f"size_{tag}"]
            h = len(k) // 2
            print(tag, "T", (2 * k[h:, 1] / (3 * n)).mean(), "U/N", (u[h:, 1] / n).mean(), "A", (sz[h:, 1] * sz[h:, 2]).mean())
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def bead24():
    """24 continuum-sphere beads in ONE BEAD molecule around a vesicle: more than 20 beads switch doBeadForce to its
    hash-cell branch (system.h:2105-2164), whose exclusion rule differs -- only the bead itself is skipped, so the beads
    also meet each other through the bead-particle term (the <= 20 branch :2167-2185 skips every bead of the molecule)."""
    tmp = tempfile.mkdtemp(prefix="golden_bead24_")
    try:
        run([os.path.join(REF, "continuumSphereAndLiposome"), "cs", "1234", "1200", "3.45", "3", "-6", "40", "5.88",
             "0", "1", "0", "2.0"], tmp)
        m = orc.read_mpd(os.path.join(tmp, "cs.mpd"))
        b = int(m["molecules"][1]["bonds"][0, 0])
        m["xyz"][b, 2] -= 2.2                                   # onto the outer leaflet, as for bead1
        lip = np.arange(m["nParticles"]) != b
        c = m["xyz"][lip].mean(0)
        rng = np.random.default_rng(24)
        extra = []
        while len(extra) < 23:                                  # the same distance from the vesicle's centre, other directions
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            x = c + q @ (m["xyz"][b] - c)
            others = np.array([m["xyz"][b]] + extra)
            d = np.linalg.norm(others - x, axis=1).min()
            if d > 5.0 and (len(extra) % 3 != 2 or d < 7.5):   # never overlapping; every third one within bead-bead range (2R + 2 = 8)
                extra.append(x)
        extra = np.array(extra)
        m2 = dict(m)
        m2["xyz"] = np.vstack([m["xyz"], extra])
        m2["vel"] = np.vstack([m["vel"], rng.normal(0, 0.1, (23, 3))])
        m2["type"] = np.append(m["type"], np.full(23, m["type"][b])).astype(np.int32)
        m2["nParticles"] = m["nParticles"] + 23
        idx = np.array([[b]] + [[m["nParticles"] + k] for k in range(23)], np.int32)
        m2["molecules"] = [m["molecules"][0], {"type": orc.BEAD, "constants": m["molecules"][1]["constants"], "bonds": idx}]
        assert np.all(m2["xyz"] > 0) and np.all(m2["xyz"] < np.array(m2["size"]))
        orc.write_mpd(os.path.join(tmp, "bead24.mpd"), m2)
        g = harness(tmp, "bead24")
        t, _ = traj(tmp, m2, "bead24_run", 8)
        save("bead24", pack(m2, g, t))
        print("bead molecule: |a| max", np.abs(g["a_mol1"]).max(), "U", g["U_mol"][1], "dU", g["dU_mol"][1])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def kat():
    """SURVEY.md 8(c)'s known-answer system at BASELINE size C1: `liposome kat5000 5000 5000 3.45` (N = 15 000, t = 0), the
    reference generator's own output and the reference's per-term quantities for the scaling (1.0005, 1.0005, 1.0005^-2)"""
    tmp = tempfile.mkdtemp(prefix="golden_kat_")
    try:
        run([os.path.join(REF, "liposome"), "kat5000", "5000", "5000", "3.45"], tmp)
        m = orc.read_mpd(os.path.join(tmp, "kat5000.mpd"))
        g = harness(tmp, "kat5000", (1.0005, 1.0005, 1.0 / 1.0005 ** 2))
        t, _ = traj(tmp, m, "kat_run", 16)   # + 16 steps of the reference `MD` binary at this size
        save("kat5000", pack(m, g, t))
        a = g["a_pair"].reshape(-1, 3)
        print("U_pair %.15e  dU_pair %.15e  dU_chain %.15e" % (g["U_pair"][0], g["dU_pair"][0], g["dU_mol"][0]))
        print("a_pair[0]", ["%.15e" % v for v in a[0]], "sum|F|^2 %.15e max|F| %.15e" % ((a * a).sum(), np.sqrt((a * a).sum(1)).max()))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    {"stats": stats, "ball": ball, "fields": fields, "substrate": substrate, "kat": kat, "bead24": bead24}.get(sys.argv[1] if len(sys.argv) > 1 else "", main)()
