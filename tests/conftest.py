import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["lipo_t0", "lipo_eq", "bondbend", "bilayer_t0", "bilayer_eq", "lipocyto_eq", "bead1", "bead2", "ball", "fields",
         "kat5000", "bead24"]   # bead24: > 20 beads in one BEAD molecule (the reference's hash-cell branch); kat5000: BASELINE config C1 at its real size (15 000 particles), SURVEY.md 8(c)'s known-answer system

# known answers of `liposome kat5000 5000 5000 3.45` at t = 0, produced by the reference code (1 thread)
KAT5000 = {"U_pair": -3.883210139620534e+05, "sumF2": 4.654216448205777e+07, "maxF": 1.880067818892180e+02,
           "a_pair0": (-5.327910857253163e+01, 2.741678778829078e+01, -5.409139740921876e+01),
           "scale": (1.0005, 1.0005, 1.0 / 1.0005 ** 2), "dU_pair": 2.854100577515757e+01, "dU_chain": -4.894054924078784e-02}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    return o


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")
