"""dataExtraction::compute's geometric observables reduced on the device (smd_observe, smd_msd_start,
smd_ke_histogram) against a numpy restatement of /root/reference/dataExtraction.h:861-937 (bond / bend means), :1457-1487
(flicker), :1511-1520 (kinetic-energy histogram) and :1525-1663 (mean square displacement per molecule) evaluated on the
arrays read back from the same context.  The files MD_b200 writes from these numbers are compared with the reference
binary's in tests/test_md_driver.py."""
import numpy as np
import pytest

import softmold_b200 as sm
from softmold_b200 import capi
from conftest import golden_path

pytestmark = pytest.mark.gpu


def image(d, box):
    d = d.copy()
    for a in range(3):
        d[:, a] = np.where(d[:, a] >= box[a] / 2.0, d[:, a] - box[a], d[:, a])
        d[:, a] = np.where(d[:, a] <= -box[a] / 2.0, d[:, a] + box[a], d[:, a])
    return d


def host_observables(m, xyz, vel, unw, unw0, box):
    out = {"lbond": 0.0, "nbond": 0, "cos": 0.0, "la": 0.0, "lb": 0.0, "nbend": 0, "msd": [], "cnt": []}
    for mol in m["molecules"]:
        t, rec = int(mol["type"]), np.asarray(mol["bonds"], dtype=np.int64)
        msd, cnt = 0.0, 0
        if t == capi.MOL_BOND:
            r = rec.reshape(-1, 2)
            d = image(xyz[r[:, 0]] - xyz[r[:, 1]], box)
            out["lbond"] += np.sqrt((d * d).sum(1)).sum()
            out["nbond"] += len(r)
        elif t == capi.MOL_BEND:
            r = rec.reshape(-1, 3)
            da, db = image(xyz[r[:, 0]] - xyz[r[:, 1]], box), image(xyz[r[:, 1]] - xyz[r[:, 2]], box)
            ra, rb = np.sqrt((da * da).sum(1)), np.sqrt((db * db).sum(1))
            out["cos"] += ((da * db).sum(1) / (ra * rb)).sum()
            out["la"] += ra.sum()
            out["lb"] += rb.sum()
            out["nbend"] += len(r)
        if t in (capi.MOL_BOND, capi.MOL_BEND, capi.MOL_BEAD):
            idx = rec.ravel()
        elif t == capi.MOL_CHAIN:
            idx = np.concatenate([np.arange(s, s + n * l) for s, n, l in rec.reshape(-1, 3)])
        else:
            idx = np.zeros(0, dtype=np.int64)
        if len(idx):
            msd, cnt = float((((unw[idx] - unw0[idx]) ** 2).sum(1)).sum()), len(idx)
        out["msd"].append(msd)
        out["cnt"].append(cnt)
    out["lo"] = np.minimum(np.asarray(box), xyz.min(0))
    out["hi"] = np.maximum(0.0, xyz.max(0))
    out["bins"] = (0.5 * (vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1] + vel[:, 2] * vel[:, 2]) / 0.0001).astype(np.int64)
    return out


@pytest.mark.parametrize("case", ["bondbend", "lipocyto_eq", "bead2", "fields", "lipo_eq"])
def test_device_observables_match_the_host_restatement(orc, case):
    m, _ = orc.load_golden(golden_path(case))
    m = dict(m, initialTime=0.0)
    ctx = sm.Context.from_dict(m, track_unwrapped=True)
    ctx.compute_forces(step=0)
    ctx.step(0, 3)
    ctx.msd_start()
    unw0 = ctx.get_unwrapped()
    hist = np.zeros(0, dtype=np.int64)
    nst = 2 if case.startswith("bead") else 10
    for rep in range(3):
        ctx.step(3 + nst * rep, nst)
        o = ctx.observe(capi.OBS_BONDS | capi.OBS_EXTENT | capi.OBS_KE_HIST | capi.OBS_MSD)
        xyz, _, vel = ctx.get_particles()
        h = host_observables(m, xyz, vel, ctx.get_unwrapped(), unw0, ctx.get_box())
        assert o["n_bond"] == h["nbond"] and o["n_bend"] == h["nbend"]
        assert abs(o["lbond_sum"] - h["lbond"]) <= 1e-12 * max(abs(h["lbond"]), 1.0)
        assert abs(o["cos_bend_sum"] - h["cos"]) <= 1e-12 * max(h["nbend"], 1)
        assert np.allclose(o["lbend_sum"], [h["la"], h["lb"]], rtol=1e-12, atol=0)
        assert np.array_equal(o["lo"], h["lo"]) and np.array_equal(o["hi"], h["hi"])      # min / max: exact
        assert list(o["msd_count"]) == h["cnt"]
        assert np.allclose(o["msd_sum"], h["msd"], rtol=1e-12, atol=1e-18)
        b = np.bincount(h["bins"])
        n = max(len(b), len(hist))
        hist = np.pad(hist, (0, n - len(hist))) + np.pad(b, (0, n - len(b)))
        assert np.array_equal(ctx.ke_histogram(), hist)                                     # integer bins: exact, accumulated
    ctx.close()


def test_observables_are_loud_about_misuse(orc):
    m, _ = orc.load_golden(golden_path("lipo_eq"))
    ctx = sm.Context.from_dict(m)            # unwrapped positions not tracked
    with pytest.raises(sm.SoftMoldError):
        ctx.msd_start()
    with pytest.raises(sm.SoftMoldError):
        ctx.observe(capi.OBS_MSD)
    assert len(ctx.ke_histogram()) == 0      # nothing observed yet
    ctx.close()
