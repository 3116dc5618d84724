"""MD_b200 (softmold_b200/csrc/md_main.cpp), the drop-in replacement of the reference's `MD <name>` executable, run
side by side with the UNMODIFIED reference binary (oracle/_ref/MD, prebuilt in the build container) on the same
`.mpd`: same files, same schedule, identical t = 0 observables, identical deterministic columns, statistically
consistent thermodynamics (the Langevin noise streams differ by design)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from softmold_b200 import workloads

pytestmark = pytest.mark.gpu

MD_B200 = os.path.join(ROOT, "softmold_b200", "MD_b200")
MD_REF = os.path.join(ROOT, "oracle", "_ref", "MD")


def table(path):
    rows = []
    for ln in open(path).read().splitlines():
        rows.append([float(x) if x not in ("-nan", "nan") else float("nan") for x in ln.split("\t")])
    return rows


def run_pair(tmp_path, orc, m, name):
    out = {}
    for tag, exe in (("ours", MD_B200), ("ref", MD_REF)):
        if not os.path.exists(exe):
            continue
        d = tmp_path / tag
        d.mkdir()
        orc.write_mpd(str(d / (name + ".mpd")), m)
        r = subprocess.run([exe, name], cwd=d, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="4"))
        assert r.returncode == 0, r.stderr[-2000:]
        out[tag] = (d, r)
    return out


def test_usage_and_missing_file(tmp_path):
    r = subprocess.run([MD_B200], capture_output=True, text=True)
    assert r.returncode == 0 and "usage:" in r.stderr             # MD.cpp:74-79
    r = subprocess.run([MD_B200, "nope"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "nope.mpd" in (r.stderr + r.stdout)


def test_liposome_run_matches_reference_outputs(tmp_path, orc):
    m = workloads.liposome(400, 3.45, 11)
    m.update(finalTime=4.0, storeInterval=2.0, measureInterval=0.4)      # 200 steps, 2 stores, 10 measures
    res = run_pair(tmp_path, orc, m, "lip")
    d, r = res["ours"]
    names = ["potential_", "size_", "lBond_", "bend_", "temp_", "kinetic_", "flicker_", "kEnergyDensity_"]
    for nm in names:
        assert (d / f"{nm}lip.dat").exists(), nm
    assert not (d / "resizeHist_lip.dat").exists()                     # deltaLXY absent
    pot = table(d / "potential_lip.dat")
    assert [round(x[0], 6) for x in pot] == [round(0.4 * k, 6) for k in range(11)]
    frames = open(d / "frames_lip.xyz").read().split("\n")
    assert frames[0] == "1200" and frames[1] == "test" and len(frames) == 3 * 1202 + 1
    chk = orc.read_mpd(str(d / "lip.mpd"))
    assert chk["initialTime"] == 4.0 and chk["nParticles"] == 1200 and chk["molecules"][0]["bonds"].tolist() == [[0, 400, 3]]
    assert "Resize acceptance ratio" in r.stderr and "starting main loop" in r.stderr
    if "ref" not in res:
        pytest.skip("oracle/_ref/MD not built: reference side of the comparison unavailable")
    dr, rr = res["ref"]
    assert sorted(os.listdir(d)) == sorted(os.listdir(dr))             # same set of files
    for nm in names[:-1]:
        a, b = table(d / f"{nm}lip.dat"), table(dr / f"{nm}lip.dat")
        assert len(a) == len(b), nm
        assert [x[0] for x in a] == [x[0] for x in b], nm              # same time stamps
        assert np.allclose(a[0], b[0], rtol=1e-5, equal_nan=True), nm  # t = 0: same configuration, 6 printed digits
    for nm in ("size_", "temp_", "lBond_", "bend_"):                   # deterministic columns agree throughout
        assert np.allclose(table(d / f"{nm}lip.dat"), table(dr / f"{nm}lip.dat"), rtol=1e-5, equal_nan=True), nm
    # thermodynamics: different noise streams, same ensemble
    ka, kb = np.array(table(d / "kinetic_lip.dat"))[3:, 1], np.array(table(dr / "kinetic_lip.dat"))[3:, 1]
    pa, pb = np.array(pot)[3:, 1], np.array(table(dr / "potential_lip.dat"))[3:, 1]
    assert abs(ka.mean() - kb.mean()) < 0.08 * kb.mean()
    assert abs(pa.mean() - pb.mean()) < 0.08 * abs(pb.mean())
    cr = orc.read_mpd(str(dr / "lip.mpd"))
    for k in ("initialTime", "finalTime", "nParticles", "seed", "storeInterval", "measureInterval", "size"):
        assert chk[k] == cr[k], k
    assert open(dr / "frames_lip.xyz").read().count("test\n") == 3


def test_bilayer_with_tension_box_moves_and_restart(tmp_path, orc):
    m = workloads.bilayer(600, 3.11, 5, tension=0.5)
    m.update(finalTime=1.6, storeInterval=0.8, measureInterval=0.4)      # 80 steps: MC trials at 8, 16, ..., 80
    res = run_pair(tmp_path, orc, m, "bl")
    d, r = res["ours"]
    hist = table(d / "resizeHist_bl.dat")
    assert abs(sum(h[1] + h[2] for h in hist) - 10.0) < 1e-9               # 10 trials, each lands in one bin
    ratio = float(r.stderr.split("Resize acceptance ratio:")[1].split()[0])
    assert 0.0 <= ratio <= 1.0
    size = table(d / "size_bl.dat")
    v0, v1 = size[0][1] * size[0][2] * size[0][3], size[-1][1] * size[-1][2] * size[-1][3]
    assert abs(v1 - v0) < 1e-4 * v0                                         # constant-volume moves (6 printed digits)
    assert any(row[1] != size[0][1] for row in size)                        # and the box did move
    if "ref" in res:
        dr, rr = res["ref"]
        assert len(table(dr / "resizeHist_bl.dat")) == len(hist)
        assert sorted(os.listdir(d)) == sorted(os.listdir(dr))
    # restart from the checkpoint the run left behind (MD.cpp:274-308): continues to a later finalTime
    chk = orc.read_mpd(str(d / "bl.mpd"))
    assert chk["initialTime"] == 1.6
    chk["finalTime"] = 2.0
    d2 = tmp_path / "restart"
    d2.mkdir()
    orc.write_mpd(str(d2 / "bl.mpd"), chk)
    r2 = subprocess.run([MD_B200, "bl"], cwd=d2, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr[-2000:]
    pot = table(d2 / "potential_bl.dat")
    assert [round(x[0], 6) for x in pot] == [2.0]                          # no t0 sample on restart, one measure at 2.0
    assert not (d2 / "frames_bl.xyz").exists() or open(d2 / "frames_bl.xyz").read().count("test\n") == 0


def test_resize_rate_of_the_anneal_variant(tmp_path, orc):
    """MDanneal.cpp is MD.cpp with `resizeRate 4` (and the temperature ramp every driver has): SMD_RESIZE_RATE=4 doubles the
    number of Metropolis trials of the same run; a temperature ramp (tempStepInterval) reaches the thermostat"""
    m = workloads.bilayer(600, 3.11, 5, tension=0.5)
    m.update(finalTime=1.6, storeInterval=1.6, measureInterval=0.4, finalTemp=1.0, tempStepInterval=0.4)
    trials = {}
    for rate in ("8", "4"):
        d = tmp_path / rate
        d.mkdir()
        orc.write_mpd(str(d / "bl.mpd"), m)
        r = subprocess.run([MD_B200, "bl"], cwd=d, capture_output=True, text=True, timeout=600, env=dict(os.environ, SMD_RESIZE_RATE=rate))
        assert r.returncode == 0, r.stderr[-2000:]
        trials[rate] = sum(h[1] + h[2] for h in table(d / "resizeHist_bl.dat"))
        temp = [row[1] for row in table(d / "temp_bl.dat")]
        assert temp[0] == 3.0 and temp[-1] < temp[0]              # ramped down towards finalTemp
    assert abs(trials["8"] - 10.0) < 1e-9 and abs(trials["4"] - 20.0) < 1e-9


def test_every_molecule_kind_of_the_md_switch(tmp_path, orc):
    """tests/golden/fields: BOUNDARY, FLOATING_BASE, ZTORQUE, ZPOWERPOTENTIAL, NANOCORE, BALL (+ SOLID and OFFSET_BOUNDARY,
    which `MD` parses and ignores) on a periodic bilayer with box moves.  Same files as the reference binary, same t = 0
    observables incl. beadPotential_ (NANOCORE counts as a bead, dataExtraction.h:971-978), checkpoint carries every
    molecule back."""
    from conftest import golden_path
    m, _ = orc.load_golden(golden_path("fields"))
    m = dict(m, initialTime=0.0, finalTime=0.8, storeInterval=0.4, measureInterval=0.2)
    res = run_pair(tmp_path, orc, m, "fld")
    d, r = res["ours"]
    assert (d / "beadPotential_fld.dat").exists()
    chk = orc.read_mpd(str(d / "fld.mpd"))
    assert [mol["type"] for mol in chk["molecules"]] == [mol["type"] for mol in m["molecules"]]
    for a, b in zip(chk["molecules"], m["molecules"]):
        assert np.array_equal(a["bonds"], b["bonds"]) and np.allclose(a["constants"], b["constants"], rtol=1e-14)
    assert chk["initialTime"] == 0.8
    if "ref" not in res:
        pytest.skip("oracle/_ref/MD not built: reference side of the comparison unavailable")
    dr, rr = res["ref"]
    assert sorted(os.listdir(d)) == sorted(os.listdir(dr))
    for nm in ("potential_", "beadPotential_", "kinetic_", "size_", "lBond_", "bend_", "flicker_"):
        a, b = table(d / f"{nm}fld.dat"), table(dr / f"{nm}fld.dat")
        assert [x[0] for x in a] == [x[0] for x in b], nm
        assert np.allclose(a[0], b[0], rtol=1e-5, equal_nan=True), (nm, a[0], b[0])


def test_per_type_friction_command(tmp_path, orc):
    """`gammaType` with gamma 0 (MD.cpp:129-138).  The reference's own .mpd parser writes the values into a vector it never
    sized (system.h:1269) and the binary dies on such a file, so there is no reference run to compare with; the semantics
    are those of Langevin::compute (langevin.h:236-281): every particle gets entry 0, the noise amplitude is fixed at the
    first temperature.  With gammaType[0] = gamma the run must therefore reproduce the plain-gamma run bit for bit."""
    m = workloads.liposome(300, 3.45, 3)
    m.update(finalTime=0.8, storeInterval=0.8, measureInterval=0.2)
    outs = {}
    for tag, extra in (("plain", {}), ("typed", {"gamma": 0.0, "gammaType": [m["gamma"]] + [0.125] * (m["nTypes"] - 1)})):
        d = tmp_path / tag
        d.mkdir()
        orc.write_mpd(str(d / "g.mpd"), dict(m, **extra))
        r = subprocess.run([MD_B200, "g"], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = orc.read_mpd(str(d / "g.mpd"))
    assert np.array_equal(outs["plain"]["xyz"], outs["typed"]["xyz"]) and np.array_equal(outs["plain"]["vel"], outs["typed"]["vel"])
    assert outs["typed"]["gammaType"][1] == 0.125 and outs["typed"]["gamma"] == 0.0     # echoed back as given
    d = tmp_path / "none"
    d.mkdir()
    orc.write_mpd(str(d / "g.mpd"), dict(m, gamma=0.0))
    r = subprocess.run([MD_B200, "g"], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "No gamma available" in r.stdout                        # MD.cpp:139-143


def test_long_run_observables_agree_with_the_reference_statistically(tmp_path):
    """north star: "long-run observables (area per lipid, membrane tension, temperature) agree statistically".
    tests/golden/stat_bilayer.npz (oracle/make_golden.py stats) holds the reference `MD` executable's own kinetic_ /
    potential_ / size_ series of two independent 20 000-step runs of a tensionless bilayer with box moves; MD_b200
    runs the same input (its own Philox noise).  Means over the second half must agree within the scatter the two
    reference runs show between themselves (plus the standard error of a 200-sample mean)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "stat_bilayer.npz"))
    d = tmp_path / "stat"
    d.mkdir()
    (d / "sb.mpd").write_bytes(g["mpd_text"].tobytes())
    r = subprocess.run([MD_B200, "sb"], cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    n = 1734

    def means(k, u, sz):
        h = len(k) // 2
        T, U, A = 2 * k[h:, 1] / (3 * n), u[h:, 1] / n, sz[h:, 1] * sz[h:, 2]
        blocks = lambda x: x[: len(x) // 10 * 10].reshape(10, -1).mean(axis=1)     # 10 block averages -> standard error
        return [(x.mean(), blocks(x).std(ddof=1) / np.sqrt(10)) for x in (T, U, A)]

    ours = means(np.loadtxt(d / "kinetic_sb.dat"), np.loadtxt(d / "potential_sb.dat"), np.loadtxt(d / "size_sb.dat"))
    ra = means(g["kinetic_a"], g["potential_a"], g["size_a"])
    rb = means(g["kinetic_b"], g["potential_b"], g["size_b"])
    assert len(np.loadtxt(d / "kinetic_sb.dat")) == len(g["kinetic_a"]) == 401
    for name, (mo, so), (ma, sa), (mb, sb) in zip(("temperature", "potential per particle", "box area"), ours, ra, rb):
        ref = 0.5 * (ma + mb)
        tol = 3.0 * abs(ma - mb) + 5.0 * max(so, sa, sb) + 2e-3 * abs(ref)
        assert abs(mo - ref) <= tol, (name, mo, ma, mb, tol)
    # the thermostat holds the set temperature in both codes (T = 3; dt = 0.02 gives the same small offset)
    assert abs(ours[0][0] - 3.0) < 0.06 and abs(ra[0][0] - 3.0) < 0.06
    ratio = float(r.stderr.split("Resize acceptance ratio:")[1].split()[0])
    assert 0.2 < ratio < 0.95
