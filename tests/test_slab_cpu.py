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


def test_step_mc_host_sequence_arms_the_call_that_holds_the_last_step():
    """softmold_b200/slab.py step_mc without a GPU: recording stand-ins for the contexts.  The scaling is proposed from the
    box BEFORE the steps, every rank is armed with it right before the smd_step call that contains the last step (the
    batched driver cuts a run into calls of four steps), dpotential is asked for that same scaling, the all-reduced sum
    decides, and an accepted move rescales every rank."""
    from softmold_b200.slab import LocalSlabGroup

    class FakeCtx:
        def __init__(self, share):
            self.calls, self.share, self.box = [], share, np.array([30.0, 20.0, 40.0])

        def get_box(self):
            return self.box.copy()

        def arm_dpotential(self, scale):
            self.calls.append(("arm", tuple(scale)))

        def step(self, first, n):
            self.calls.append(("step", first, n))

        def dpotential(self, scale):
            self.calls.append(("dU", tuple(scale)))
            t = np.zeros(capi.NTERMS)
            t[capi.TERM_PAIR] = self.share
            return t

        def rescale(self, scale, new_box):
            self.calls.append(("rescale", tuple(scale)))
            self.box = np.array(new_box)

    grp = LocalSlabGroup.__new__(LocalSlabGroup)
    grp.nranks, grp.temperature, grp.batched_default = 3, 3.0, True
    grp.ctx = [FakeCtx(0.5), FakeCtx(0.25), FakeCtx(0.25)]        # partial sums: +1.0 in total -> always accepted
    box0 = grp.get_box()
    new_box, scale = capi.mc_propose(box0, 0.01, 0.9)
    acc, dU, box = grp.step_mc(16, 10, 0.01, 0.5, 0.9, 0.3)
    assert acc and np.array_equal(box, new_box)
    assert dU == 1.0 + 0.5 * (new_box[0] * new_box[1] - box0[0] * box0[1])
    for c in grp.ctx:
        assert c.calls == [("step", 16, 4), ("step", 20, 4), ("arm", tuple(scale)), ("step", 24, 2), ("dU", tuple(scale)),
                           ("rescale", tuple(scale))]
    # a rejected move rescales nobody
    grp.ctx = [FakeCtx(-40.0), FakeCtx(-40.0)]
    grp.nranks = 2
    acc, _, box = grp.step_mc(0, 4, 0.01, 0.0, 0.9, 0.999)
    assert not acc and all(c.calls[-1][0] == "dU" and c.calls[0][0] == "arm" for c in grp.ctx)


def test_two_rank_gloo_partition_and_replicated_decisions():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29572", os.path.join(ROOT, "tests", "slab_dist_worker.py"), "cpu"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SLAB_CPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
