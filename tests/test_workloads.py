"""The synthetic-input builders of softmold_b200/workloads.py against configurations written by the reference's own
generator programs (golden fixtures lipo_t0 = `liposome lipo 42 300 3.45`, bilayer_t0 = `bilayer bl 99 600 3.11 0 0 0 0`,
see oracle/make_golden.py).  The fixtures went through the reference's 15-significant-digit text format."""
import numpy as np

from conftest import golden_path
from softmold_b200 import workloads


def _same(m, g, ref_has_tension=False):
    assert m["nParticles"] == g["nParticles"] and m["nTypes"] == g["nTypes"]
    np.testing.assert_allclose(m["size"], g["size"], rtol=1e-14)
    assert np.array_equal(m["type"], g["type"])
    np.testing.assert_allclose(m["xyz"], g["xyz"], rtol=0, atol=2e-12)
    np.testing.assert_allclose(m["vel"], g["vel"], rtol=0, atol=1e-13)
    np.testing.assert_array_equal(m["twoBodyFconst"], g["twoBodyFconst"])
    np.testing.assert_array_equal(m["twoBodyUconst"], g["twoBodyUconst"])
    assert len(m["molecules"]) == len(g["molecules"]) == 1
    assert m["molecules"][0]["type"] == g["molecules"][0]["type"]
    assert np.array_equal(m["molecules"][0]["bonds"], g["molecules"][0]["bonds"])
    np.testing.assert_array_equal(m["molecules"][0]["constants"], g["molecules"][0]["constants"])
    for k in ("cutoff", "deltaT", "gamma", "initialTemp", "seed"):
        assert m[k] == g[k], k


def test_liposome_matches_reference_generator(orc):
    g, _ = orc.load_golden(golden_path("lipo_t0"))
    _same(workloads.liposome(300, 3.45, 42), g)


def test_bilayer_matches_reference_generator(orc):
    g, _ = orc.load_golden(golden_path("bilayer_t0"))
    m = workloads.bilayer(600, 3.11, 99, tension=0.5)
    _same(m, g)
    assert m["deltaLXY"] == g["deltaLXY"] == 0.01 and m["tension"] == g["tension"]


def test_headline_workload_shape():
    m = workloads.liposome(80000, 3.45, 777)
    assert m["nParticles"] == 240000 and m["size"] == [400.0, 400.0, 400.0]
    assert np.all(m["xyz"] > 0) and np.all(m["xyz"] < 400)
    ke = 0.5 * (m["vel"] ** 2).sum()
    assert abs(ke - 240000 * 4.5) < 1e-6 * ke     # |v|^2 = 3T = 9 for every particle
