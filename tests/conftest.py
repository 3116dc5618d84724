import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["lipo_t0", "lipo_eq", "bondbend", "bilayer_t0", "bilayer_eq", "lipocyto_eq", "bead1", "bead2", "ball", "fields"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    return o


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")
