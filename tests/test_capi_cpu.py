"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the `.mpd` reader/writer reproduces the reference's own files, and the product refuses to run without a GPU
(no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import softmold_b200 as sm
from conftest import ROOT, golden_path


def test_library_exports_every_declared_symbol():
    L = sm.lib()
    header = open(os.path.join(ROOT, "include", "softmold_b200.h")).read()
    declared = set(re.findall(r"\b(smd_[a-z0-9_]+)\s*\(", header))
    declared -= {"smd_ctx", "smd_mpd", "smd_desc"}
    assert declared == set(sm.SYMBOLS), declared ^ set(sm.SYMBOLS)
    for s in declared:
        assert hasattr(L, s)
    assert L.smd_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    if sm.lib().smd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sm.SoftMoldError) as e:
        sm.Context(10, 2, [10, 10, 10], 2.0, 0.02, 1.0, 3.0, 1)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "softmold_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/oracle.c", "").replace("oracle.orc.read_mpd", "").replace("restated in oracle", ""), f


def test_mpd_roundtrip_is_byte_identical_to_reference_written_file(tmp_path):
    src = os.path.join(ROOT, "tests", "golden", "ref_written_lipo20")   # written by the reference's Script::write
    m = sm.Mpd(src)
    out = str(tmp_path / "rt")
    m.write(out)
    assert open(src + ".mpd", "rb").read() == open(out + ".mpd", "rb").read()
    d = m.to_dict()
    assert d["nParticles"] == 60 and d["nTypes"] == 6 and d["molecules"][0]["type"] == sm.MOL_CHAIN
    assert list(d["molecules"][0]["bonds"][0]) == [0, 20, 3]
    assert "deltaLXY" not in d and "tension" not in d


def test_mpd_reader_matches_independent_python_reader(orc, tmp_path):
    for case in ("bead2", "lipocyto_eq", "bilayer_t0"):
        m, _ = orc.load_golden(golden_path(case))
        m["finalTime"], m["storeInterval"], m["measureInterval"] = 100.0, 10.0, 1.0
        path = str(tmp_path / case)
        orc.write_mpd(path + ".mpd", m)
        d = sm.Mpd(path).to_dict()
        assert np.array_equal(d["xyz"], m["xyz"]) and np.array_equal(d["vel"], m["vel"]) and np.array_equal(d["type"], m["type"])
        assert np.array_equal(d["twoBodyFconst"], m["twoBodyFconst"]) and np.array_equal(d["twoBodyUconst"], m["twoBodyUconst"])
        assert len(d["molecules"]) == len(m["molecules"])
        for a, b in zip(d["molecules"], m["molecules"]):
            assert a["type"] == b["type"] and np.array_equal(a["bonds"], b["bonds"]) and np.array_equal(a["constants"], b["constants"])
        for k in ("deltaLXY", "tension"):
            assert d.get(k) == m.get(k)


def test_mpd_errors_like_the_reference(tmp_path):
    p = tmp_path / "bad.mpd"
    p.write_text("gamma 1 nTypes 2 frobnicate 3\n")
    with pytest.raises(sm.SoftMoldError, match="not a recognized command"):
        sm.Mpd(str(tmp_path / "bad"))
    p.write_text("nParticles 1 size 4 4 4 positions 1 5.0 1 1 velocities 0 0 0\n")
    with pytest.raises(sm.SoftMoldError, match="X position of particle 0 is out of bounds"):
        sm.Mpd(str(tmp_path / "bad"))
    p.write_text("positions 1 1 1 1\n")
    with pytest.raises(sm.SoftMoldError, match="nParticles was not present before positions"):
        sm.Mpd(str(tmp_path / "bad"))
    with pytest.raises(sm.SoftMoldError, match="Could not open"):
        sm.Mpd(str(tmp_path / "missing"))


def test_documented_switches_exist_in_the_sources():
    """README.md's table of run-time switches against the sources: a documented variable that nothing reads is a lie"""
    import re
    readme = open(os.path.join(ROOT, "README.md")).read()
    table = readme[readme.index("Run-time switches"):]
    names = set(re.findall(r"`(SMD_[A-Z0-9_]+)", table))
    assert len(names) >= 10
    src = ""
    for f in ("smd_core.cu", "md_main.cpp", "smd_kernels.cuh"):
        src += open(os.path.join(ROOT, "softmold_b200", "csrc", f)).read()
    missing = [n for n in sorted(names) if f'"{n}"' not in src]
    assert not missing, missing
