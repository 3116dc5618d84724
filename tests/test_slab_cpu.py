"""Host-side logic of the slab decomposition, no GPU: the partition functions of the C ABI and, with two gloo ranks,
partition completeness and the replicated Metropolis decision."""
import os
import subprocess
import sys

import numpy as np

import softmold_b200 as sm
from softmold_b200 import capi
from conftest import ROOT


def test_slab_columns_tile_the_grid():
    for nc in (10, 163, 1309, 2047):
        for g in (1, 2, 3, 4, 8):
            edges = [capi.slab_columns(nc, g, r) for r in range(g)]
            assert edges[0][0] == 0 and edges[-1][1] == nc
            assert all(edges[r][1] == edges[r + 1][0] for r in range(g - 1))
            w = [hi - lo for lo, hi in edges]
            assert max(w) - min(w) <= 1


def test_slab_select_uses_the_reference_cell_column():
    rng = np.random.default_rng(0)
    box = [41.3, 30.0, 20.0]
    xyz = rng.random((5000, 3)) * box
    xyz[0, 0] = box[0]            # p == L survives the wrap (verlet.h:337); CellOpt::build clamps it (cellOpt.h:537)
    xyz[1, 0] = 0.0
    nc = int(box[0] / 2.0)
    col = np.minimum((xyz[:, 0] / (box[0] / nc)).astype(int), nc - 1)
    owners = np.zeros(len(xyz), int)
    for r in range(4):
        f = capi.slab_select(box, 2.0, 4, r, xyz)
        lo, hi = capi.slab_columns(nc, 4, r)
        assert np.array_equal(f == 1, (col >= lo) & (col < hi))
        owners += (f == 1)
        # ghosts: the two columns on either side, periodic
        g = np.isin((col - lo) % nc, [nc - 2, nc - 1]) | np.isin((col - hi) % nc, [0, 1])
        assert np.array_equal(f == 2, g & (f != 1))
    assert np.all(owners == 1)
    assert np.array_equal(capi.slab_select(box, 2.0, 1, 0, xyz), np.ones(len(xyz), np.int32))


def test_mc_helpers_follow_md_cpp():
    box = np.array([30.0, 20.0, 40.0])
    new_box, scale = capi.mc_propose(box, 0.01, 0.75)
    fl = 0.01 * (2.0 * 0.75 - 1.0)
    assert new_box[0] == box[0] + fl and new_box[1] == box[1] + fl
    assert new_box[2] == box[2] * ((box[0] * box[1]) / ((box[0] + fl) * (box[1] + fl)))
    assert np.array_equal(scale, new_box / box)
    assert abs(np.prod(new_box) - np.prod(box)) < 1e-9           # volume preserving (MD.cpp:594-606)
    dA = new_box[0] * new_box[1] - box[0] * box[1]
    acc, dU = capi.mc_accept(-1.0, 0.5, box, new_box, 3.0, 0.99)
    assert dU == -1.0 + 0.5 * dA and acc == (np.exp(dU / 3.0) >= 0.99)
    assert capi.mc_accept(+1.0, 0.0, box, new_box, 3.0, 0.999999)[0]      # dU >= 0 always accepted (MD.cpp:695)
    assert not capi.mc_accept(-50.0, 0.0, box, new_box, 3.0, 0.5)[0]


def test_two_rank_gloo_partition_and_replicated_decisions():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29572", os.path.join(ROOT, "tests", "slab_dist_worker.py"), "cpu"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SLAB_CPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
