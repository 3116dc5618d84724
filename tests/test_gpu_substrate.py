"""the molecule switch of MDsubstrate.cpp (MDsubstrate.cpp:213-262, :476-490) as a run-time option --
OFFSET_BOUNDARY, RIGIDBEND and PULLBEAD act, BALL / FLOATING_BASE / ZTORQUE / ZPOWERPOTENTIAL / NANOCORE are ignored.
MDsubstrate.cpp itself does not compile in the reference tree (it includes include/fileFormats/vmdOutput.h, which is not
there), so the three force routines are pinned through the unmodified Blob::do*Force members (golden fixture `substrate`,
oracle/make_golden.py substrate) and the schedule through MD's, which it shares."""
import os
import subprocess

import numpy as np
import pytest

import softmold_b200 as sm
from softmold_b200 import capi
from conftest import ROOT, golden_path

pytestmark = pytest.mark.gpu


def _only(m, keep):
    """the system with only the molecules whose index is in keep (the others dropped)"""
    return dict(m, molecules=[mol for k, mol in enumerate(m["molecules"]) if k in keep], nMolecules=len(keep))


def test_substrate_force_terms_match_the_reference(orc):
    m, ref = orc.load_golden(golden_path("substrate"))
    kinds = [mol["type"] for mol in m["molecules"]]
    for k, t in enumerate(kinds):
        if t not in (capi.MOL_OFFSET_BOUNDARY, capi.MOL_RIGIDBEND, capi.MOL_PULLBEAD, capi.MOL_BOUNDARY):
            continue
        ctx = sm.Context.from_dict(_only(m, [k]), driver="substrate")
        ctx.compute_forces(mask=1 << sm.TERM_FIELD)
        a = ctx.get_forces()
        r = ref[f"a_mol{k}"].reshape(-1, 3)
        assert np.abs(r).max() > 0
        # RIGIDBEND goes through asin / pow of the CUDA library (<= 2 ulp from glibc's); the others are +, *, / only
        tol = 1e-13 if t == capi.MOL_RIGIDBEND else 1e-15
        assert np.abs(a - r).max() <= tol * np.abs(r).max(), (k, t, np.abs(a - r).max() / np.abs(r).max())
        ctx.close()


def test_each_driver_ignores_the_other_drivers_kinds(orc):
    m, ref = orc.load_golden(golden_path("substrate"))
    kinds = [mol["type"] for mol in m["molecules"]]
    total = {}
    for driver in ("md", "substrate"):
        ctx = sm.Context.from_dict(m, driver=driver)
        ctx.compute_forces(mask=sm.MASK_ALL_MOLECULES)
        total[driver] = ctx.get_forces()
        ctx.close()
    acts_in = {"md": (capi.MOL_CHAIN, capi.MOL_BOUNDARY, capi.MOL_ZPOWERPOTENTIAL, capi.MOL_BALL),
               "substrate": (capi.MOL_CHAIN, capi.MOL_BOUNDARY, capi.MOL_OFFSET_BOUNDARY, capi.MOL_RIGIDBEND, capi.MOL_PULLBEAD)}
    # the substrate fixture's per-molecule forces come from the substrate switch; ZPOWERPOTENTIAL and BALL are zero there, so
    # only the substrate driver's total can be assembled from it
    want = sum(ref[f"a_mol{k}"].reshape(-1, 3) for k, t in enumerate(kinds) if t in acts_in["substrate"])
    assert np.abs(total["substrate"] - want).max() <= 1e-12 * np.abs(want).max()
    # the MD driver sees CHAIN + BOUNDARY + ZPOWERPOTENTIAL + BALL and none of the three substrate kinds
    sub_only = sum(ref[f"a_mol{k}"].reshape(-1, 3) for k, t in enumerate(kinds) if t in (capi.MOL_OFFSET_BOUNDARY, capi.MOL_RIGIDBEND, capi.MOL_PULLBEAD))
    common = sum(ref[f"a_mol{k}"].reshape(-1, 3) for k, t in enumerate(kinds) if t in (capi.MOL_CHAIN, capi.MOL_BOUNDARY))
    md_extra = total["md"] - common
    assert np.abs(md_extra).max() > 0                                     # ZPOWERPOTENTIAL / BALL act under MD ...
    assert np.abs(total["substrate"] - common - sub_only).max() <= 1e-12 * np.abs(want).max()
    touched = np.abs(md_extra).sum(axis=1) > 1e-13
    assert touched.sum() < 200                                            # ... on their own few particles only (PULLBEAD would touch all)


def test_md_b200_substrate_driver_runs_and_writes_its_files(orc, tmp_path):
    """SMD_DRIVER=substrate: MD_b200 follows MDsubstrate.cpp's switch and writes dAcceptFile.dat / dRejectFile.dat / resizeHist.dat
    (MDsubstrate.cpp:710-734, :753); the first steps' positions equal a library run with the substrate switch."""
    exe = os.path.join(ROOT, "softmold_b200", "MD_b200")
    if not os.path.exists(exe):
        pytest.skip("MD_b200 not built")
    m, _ = orc.load_golden(golden_path("substrate"))
    nsteps = 40
    m = dict(m, initialTime=0.0, finalTime=(nsteps - 0.5) * m["deltaT"], storeInterval=32 * m["deltaT"], measureInterval=16 * m["deltaT"],
             deltaLXY=0.01, tension=0.4)
    orc.write_mpd(str(tmp_path / "sub.mpd"), m)
    env = dict(os.environ, SMD_DRIVER="substrate")
    r = subprocess.run([exe, "sub"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = []
    for f in ("dAcceptFile.dat", "dRejectFile.dat"):
        if (tmp_path / f).exists():
            rows += [l.split() for l in open(tmp_path / f).read().splitlines()]
    assert len(rows) == 4 and sorted(float(x[0]) for x in rows) == [8 * m["deltaT"] * k for k in (1, 2, 3, 4)]     # a trial after steps 8, 16, 24, 32
    assert (tmp_path / "resizeHist.dat").exists() and not (tmp_path / "resizeHist_sub.dat").exists()
    # the same run under MD's switch differs: the substrate kinds really acted (frames of the final state)
    last = lambda p: np.loadtxt(open(p).read().splitlines()[-m["nParticles"]:], usecols=(1, 2, 3))
    x_sub = last(tmp_path / "frames_sub.xyz")
    for f in tmp_path.iterdir():
        if f.name != "sub.mpd":
            f.unlink()
    orc.write_mpd(str(tmp_path / "sub.mpd"), m)
    r = subprocess.run([exe, "sub"], cwd=tmp_path, env=dict(os.environ, SMD_DRIVER="md"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not (tmp_path / "dAcceptFile.dat").exists() and (tmp_path / "resizeHist_sub.dat").exists()
    x_md = last(tmp_path / "frames_sub.xyz")
    assert np.abs(x_sub - x_md).max() > 1e-6
