"""Slab decomposition on the GPU: nranks slab contexts against ONE single-GPU context of the same
system (which tests/test_gpu_parity.py pins against the reference), through the C ABI.

LocalSlabGroup runs all ranks inside this process on one device: same kernels, same message protocol and per-rank
state as a multi-GPU run, plain device pointers instead of CUDA IPC / NVLink as transport (the multi-process
transport is exercised by tests/test_gpu_slab_dist.py when more than one GPU is visible).

Because every particle gathers its own force over its neighbour list in cell order and the Philox noise is keyed on
the global particle index, a slab run reproduces the single-GPU trajectory to rounding: the bounds below are the
north star's (1e-5 forces, 1e-7 energies) with a much tighter design check next to them."""
import numpy as np
import pytest

import softmold_b200 as sm
from softmold_b200 import workloads
from softmold_b200.slab import LocalSlabGroup
from conftest import golden_path

pytestmark = pytest.mark.gpu


def rel_err(a, ref):
    s = np.abs(ref).max()
    return np.abs(a - ref).max() / (s if s > 0 else 1.0)


@pytest.fixture(scope="module")
def bilayer():
    m = workloads.bilayer(5000, 3.11, seed=11)     # 15 000 particles, 40.1 x 40.1 x 40 box: 20 cell columns
    # thermalise a little on one GPU so that the cells are not lattice-aligned
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces()
    ctx.step(0, 60)
    m = dict(m)
    m["xyz"], m["type"], m["vel"] = ctx.get_particles()
    ctx.close()
    return m


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_slab_forces_and_energies_match_single_gpu(bilayer, nranks):
    m = bilayer
    n = m["nParticles"]
    one = sm.Context.from_dict(m)
    one.compute_forces(mask=sm.MASK_ALL, step=77)
    a1 = one.get_forces()
    U1, K1 = one.potential(), one.kinetic()
    scale = [1.0005, 1.0005, 1.0 / (1.0005 * 1.0005)]
    dU1 = one.dpotential(scale)
    pairs1, _ = one.count_pairs(per_particle=False)
    one.close()

    grp = LocalSlabGroup(m, nranks)
    grp.compute_forces(mask=sm.MASK_ALL, step=77)
    xyz, typ, vel, acc, owner = grp.gather(n)
    assert np.array_equal(xyz, m["xyz"]) and np.array_equal(typ, m["type"]) and np.array_equal(vel, m["vel"])
    # ownership = the cell column ranges of the C ABI's own partition function
    for r in range(nranks):
        flags = sm.capi.slab_select(m["size"], m["cutoff"], nranks, r, m["xyz"])
        assert np.array_equal(flags == 1, owner == r)
    assert len(set(owner)) == nranks
    err = rel_err(acc, a1)
    assert err <= 1e-5
    assert err <= 1e-12, err
    U, K, dU = grp.potential(), grp.kinetic(), grp.dpotential(scale)
    assert grp.count_pairs() == pairs1                             # neighbour membership: every pair exactly once
    for t in (sm.TERM_PAIR, sm.TERM_CHAIN):
        assert abs(U[t] - U1[t]) <= 1e-7 * abs(U1[t])
        assert abs(U[t] - U1[t]) <= 1e-12 * abs(U1[t])
        assert abs(dU[t] - dU1[t]) <= 1e-12 * abs(U1[t]) + 1e-9 * abs(dU1[t])
    assert abs(K - K1) <= 1e-13 * K1
    grp.close()


@pytest.mark.parametrize("nranks", [2, 4])
def test_slab_trajectory_with_box_moves_matches_single_gpu(bilayer, orc, nranks):
    """60 steps of the MD.cpp loop incl. a Metropolis box move with tension every 8 steps: particles migrate between
    slabs, the halo is renewed every step, the box (and with it every column boundary) changes"""
    m = bilayer
    n, K = m["nParticles"], 60
    mc = orc.mt_rand53(m["seed"], 2 * (K // 8 + 2))
    tension = 0.5

    def run(sim, box_move):
        sim.compute_forces(mask=sm.MASK_ALL, step=0)
        trials = acc = 0
        for i in range(K):
            sim.step(i, 1)
            if i % 8 == 0 and i != 0:
                ok = box_move(m["deltaLXY"], tension, mc[2 * trials], mc[2 * trials + 1])[0]
                trials += 1
                acc += bool(ok)
        return trials, acc

    one = sm.Context.from_dict(m)
    t1, acc1 = run(one, one.mc_box_move)
    x1, _, v1 = one.get_particles()
    box1 = one.get_box()
    one.close()

    grp = LocalSlabGroup(m, nranks)
    t2, acc2 = run(grp, grp.mc_box_move)
    x2, _, v2, _, owner = grp.gather(n)
    box2 = grp.get_box()
    moved = sum(int(np.sum((sm.capi.slab_select(m["size"], m["cutoff"], nranks, r, m["xyz"]) == 1) & (owner != r))) for r in range(nranks))
    grp.close()
    assert (t1, acc1) == (t2, acc2) and acc1 > 0
    np.testing.assert_allclose(box2, box1, rtol=1e-14)
    assert moved > 0                                    # the run did exercise migration
    assert np.abs(x2 - x1).max() <= 1e-9
    assert np.abs(v2 - v1).max() <= 1e-8


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_step_mc_sums_the_dpotential_in_the_force_pass(bilayer, orc, nranks):
    """step_mc in slab mode: every rank arms the proposed scaling (smd_arm_dpotential), the pair kernel of its last step sums
    its share of the pair dPotential, the all-reduced total and the decisions equal those of step + mc_box_move, and the
    dPotential kernel is not launched"""
    m = bilayer
    mc = orc.mt_rand53(3, 32)
    out = []
    for fused in (False, True):
        grp = LocalSlabGroup(m, nranks)
        grp.batched_default = True
        grp.compute_forces(mask=sm.MASK_ALL, step=0)
        log = []
        for t in range(4):
            if fused:
                acc, dU, box = grp.step_mc(8 * t, 8, m["deltaLXY"], 0.5, mc[2 * t], mc[2 * t + 1])
            else:
                grp.step(8 * t, 8)
                acc, dU, box = grp.mc_box_move(m["deltaLXY"], 0.5, mc[2 * t], mc[2 * t + 1])
            log.append((acc, dU, tuple(box)))
        x, _, v, _, _ = grp.gather(m["nParticles"])
        launches = sum(c.stats()[0] for c in grp.ctx)
        U = abs(grp.potential()[sm.TERM_PAIR])
        grp.close()
        out.append((log, x, v, launches, U))
    (l0, x0, v0, n0, U), (l1, x1, v1, n1, _) = out
    assert [a for a, _, _ in l0] == [a for a, _, _ in l1] and any(a for a, _, _ in l0)
    for (_, d0, b0), (_, d1, b1) in zip(l0, l1):
        assert abs(d0 - d1) <= 1e-12 * U and b0 == b1
    assert np.array_equal(x0, x1) and np.array_equal(v0, v1)      # forces of the fused pass are bit-identical
    assert n1 < n0                                                # four dPotential pair kernels per rank fewer


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_exchange_packed_by_the_step_seam_is_bit_identical(bilayer, nranks):
    """smd_step in slab mode: the fused seam kernel is also the send side of the exchange (every owned particle is packed
    into the neighbour's buffer as it gets its new position).  Against SMD_NO_SEAM_PACK=1 (a pack kernel of its own after
    the seam): same particles on the same ranks, bit for bit, after 40 steps with migration; one launch per step fewer."""
    import os
    m = bilayer
    out = []
    for env in ("1", "0"):
        os.environ["SMD_NO_SEAM_PACK"] = env
        try:
            grp = LocalSlabGroup(m, nranks)
        finally:
            del os.environ["SMD_NO_SEAM_PACK"]
        grp.compute_forces(mask=sm.MASK_ALL, step=0)
        grp.step(0, 40, batched=True)
        x, t, v, a, owner = grp.gather(m["nParticles"])
        launches = sum(c.stats()[0] for c in grp.ctx)
        grp.close()
        out.append((x, v, owner, launches))
    (x0, v0, o0, l0), (x1, v1, o1, l1) = out
    assert np.array_equal(x0, x1) and np.array_equal(v0, v1) and np.array_equal(o0, o1)
    assert l1 < l0


@pytest.mark.parametrize("nranks", [2, 4])
def test_slab_vesicle_across_the_seam_and_empty_ranks(orc, nranks):
    """a small vesicle in the middle of a 400^3 box: the seam of a 2-rank split cuts it in half; with 4 ranks two of
    them own nothing but still take part in every exchange"""
    m, ref = orc.load_golden(golden_path("lipo_eq"))
    m["initialTime"] = 0.0
    n = m["nParticles"]
    one = sm.Context.from_dict(m)
    one.compute_forces(step=3)
    one.step(3, 25)
    x1, _, v1 = one.get_particles()
    one.close()
    grp = LocalSlabGroup(m, nranks, capacity=4096, msg_capacity=4096)
    grp.compute_forces(step=3)
    grp.step(3, 13)
    grp.step(16, 12, batched=True)      # smd_step batches: the fused step kernel + exchange inside one call
    x2, _, v2, _, owner = grp.gather(n)
    counts = [c.slab_counts() for c in grp.ctx]
    grp.close()
    assert np.abs(x2 - x1).max() <= 1e-10 and np.abs(v2 - v1).max() <= 1e-9
    assert sum(o for _, o in counts) == n
    if nranks == 2:
        assert min(o for _, o in counts) > 100              # really split
        assert all(l > o for l, o in counts)                # ghosts present on both sides


def test_slab_errors_are_loud(bilayer):
    m = bilayer
    # too many ranks for the box: a slab must be at least 2 * halo + 1 columns wide
    with pytest.raises(sm.SoftMoldError, match="narrower"):
        sm.Context.from_dict(m, rank=0, nranks=5)
    # a particle that jumps over the halo in one step
    grp = LocalSlabGroup(m, 2)
    grp.compute_forces()
    g, x, t, v, a = grp.ctx[0].slab_get_local()
    bad = dict(m)
    bad["vel"] = m["vel"].copy()
    bad["vel"][g[np.argmax(x[:, 0])]] = [600.0, 0.0, 0.0]          # 12 sigma = 6 columns per step, from the slab's last column
    grp.close()
    grp = LocalSlabGroup(bad, 2)
    grp.compute_forces()
    with pytest.raises(sm.SoftMoldError, match="halo"):    # reported at the first synchronisation point
        grp.step(0, 1)
        grp.synchronize()
    for c in grp.ctx:
        c.close()
    # slab contexts refuse what they do not support instead of computing something else
    ctx = sm.Context.from_dict(m, rank=0, nranks=2)
    with pytest.raises(sm.SoftMoldError, match="CHAIN, BOND and BEND molecules only"):
        ctx.add_molecule(sm.MOL_BEAD, np.array([[0]], np.int32), np.zeros(22 * m["nTypes"] ** 2))
    with pytest.raises(sm.SoftMoldError, match="connect both neighbours"):
        ctx.step(0, 1)
    ctx.close()


def test_slab_restart_from_per_rank_particles(bilayer):
    """smd_slab_set_local: every rank re-loads only its own particles (as a per-rank restart would); ghosts and
    strays travel through the ordinary exchange.  Includes particles handed to the 'wrong' neighbour rank."""
    m = bilayer
    n = m["nParticles"]
    grp = LocalSlabGroup(m, 3)
    grp.compute_forces(mask=sm.MASK_ALL, step=9)
    ref = grp.gather(n)
    parts = [c.slab_get_local() for c in grp.ctx]
    # move the 50 right-most particles of rank 0 into rank 1's hands and vice versa: each is at most 2 columns off
    g0, x0, t0, v0, _ = parts[0]
    g1, x1, t1, v1, _ = parts[1]
    a = np.argsort(x0[:, 0])[-50:]
    b = np.argsort(x1[:, 0])[:50]
    k0 = np.setdiff1d(np.arange(len(g0)), a)
    k1 = np.setdiff1d(np.arange(len(g1)), b)
    new0 = tuple(np.concatenate([p[k0], q[b]]) for p, q in ((g0, g1), (x0, x1), (t0, t1), (v0, v1)))
    new1 = tuple(np.concatenate([p[k1], q[a]]) for p, q in ((g1, g0), (x1, x0), (t1, t0), (v1, v0)))
    loads = [new0, new1, parts[2][:4]]
    for c, (g, x, t, v) in zip(grp.ctx, loads):
        c.slab_set_local(g, x, t, v)
    grp.compute_forces(mask=sm.MASK_ALL, step=9)
    got = grp.gather(n)
    for u, w in zip(ref, got):
        assert np.array_equal(u, w)
    grp.step(9, 3)
    grp.close()


def test_dpotential_left_on_the_device_equals_the_host_read_back(orc):
    """smd_dpotential_device (the buffer the multi-GPU driver all-reduces in place with NCCL): the same terms, bit for bit,
    as smd_dpotential's host read-back; nothing is copied by the call itself"""
    import torch
    from softmold_b200.slab import _DeviceDoubles
    for case in ("bilayer_eq", "lipocyto_eq", "fields"):
        m, _ = orc.load_golden(golden_path(case))
        ctx = sm.Context.from_dict(m)
        ctx.compute_forces(step=0)
        ctx.step(0, 4)
        scale = [1.002, 1.002, 1.0 / 1.002 ** 2]
        host = ctx.dpotential(scale)
        ptr = ctx.dpotential_device(scale)
        ctx.synchronize()
        dev = torch.as_tensor(_DeviceDoubles(ptr, sm.NTERMS), device="cuda").cpu().numpy()
        assert np.array_equal(dev, host), (dev, host)
        ctx.close()


@pytest.mark.parametrize("batched", [False, True])
def test_slab_histogram_filled_by_the_seam_and_the_unpack_is_bit_identical(bilayer, batched):
    """slab mode: the histogram of the cell build is filled by the kernel that moves the particles (owned particles and the
    migrants that stay behind as ghosts) and by the unpack kernel (what arrives), instead of a pass of its own over every
    slot (k_bin; SMD_NO_SLAB_PREBIN=1 keeps it).  Same sort, same forces: trajectories with box moves identical bit for bit,
    and the index-table entries of dropped ghosts are still given back (gather finds every particle exactly once)."""
    import os
    m = bilayer
    n = m["nParticles"]
    mc = np.random.RandomState(5).random_sample(16)
    out = []
    for env, push in (("1", "0"), ("0", "0"), ("0", "1")):
        # (third run: SMD_SLAB_PULL=1 -- the receiver reads the sender's buffer through the link -- against the default, where
        # the sender writes into the neighbour's buffer: same messages, same results; measured slower, kept for A/B)
        os.environ["SMD_NO_SLAB_PREBIN"], os.environ["SMD_SLAB_PULL"] = env, push
        try:
            grp = LocalSlabGroup(m, 3)
        finally:
            del os.environ["SMD_NO_SLAB_PREBIN"], os.environ["SMD_SLAB_PULL"]
        grp.batched_default = batched
        grp.compute_forces(mask=sm.MASK_ALL, step=0)
        boxes = []
        for t in range(4):
            boxes.append(grp.step_mc(8 * t, 8, 0.01, 0.4, mc[2 * t], mc[2 * t + 1])[2]) if batched else \
                (grp.step(8 * t, 8), boxes.append(grp.mc_box_move(0.01, 0.4, mc[2 * t], mc[2 * t + 1])[2]))
        xyz, typ, vel, acc, owner = grp.gather(n)
        launches = sum(c.stats()[0] for c in grp.ctx)
        out.append((xyz, vel, acc, np.array(boxes), launches))
        grp.close()
    (x0, v0, a0, b0, l0), (x1, v1, a1, b1, l1), (x2, v2, a2, b2, l2) = out
    assert np.array_equal(b0, b1) and np.array_equal(x0, x1) and np.array_equal(v0, v1) and np.array_equal(a0, a1)
    assert np.array_equal(b0, b2) and np.array_equal(x0, x2) and np.array_equal(v0, v2) and np.array_equal(a0, a2)
    assert l1 < l0 and l2 == l1          # the histogram passes are gone


def _with_bond_and_bend_lists(m, seed=3):
    """the bilayer plus a BOND list between head groups of neighbouring lipids (a cytoskeleton-like mesh, bonds up to ~2 sigma:
    well inside the two-column halo) and a BEND list over the lipids' own triplets"""
    rng = np.random.RandomState(seed)
    xyz, typ, box = m["xyz"], m["type"], np.array(m["size"])
    st, nch, ln = [int(v) for v in m["molecules"][0]["bonds"][0]]
    heads = st + ln * np.arange(nch)
    # nearest head in +x direction within 2 sigma (minimum image), brute force on a subset
    pick = heads[rng.choice(len(heads), 600, replace=False)]
    bonds = []
    for i in pick:
        d = xyz[heads] - xyz[i]
        d -= box * np.round(d / box)
        r = np.sqrt((d * d).sum(axis=1))
        j = heads[np.argsort(r)[1]]            # nearest other head
        if r[np.argsort(r)[1]] < 2.0:
            bonds.append((int(i), int(j)))
    bends = [(int(h), int(h) + 1, int(h) + 2) for h in heads[::5]]
    mols = list(m["molecules"]) + [
        {"type": sm.MOL_BOND, "constants": np.array([1.1, 40.0]), "bonds": np.array(bonds, np.int32)},
        {"type": sm.MOL_BEND, "constants": np.array([-0.7, 30.0]), "bonds": np.array(bends, np.int32)}]
    return dict(m, molecules=mols, nMolecules=len(mols))


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_bond_and_bend_lists_match_single_gpu(bilayer, nranks):
    """BOND / BEND lists in slab mode (round 2): every rank walks the whole list with global indices, evaluates the records
    that have an owned member (the partners sit in its halo), adds forces to owned members only and counts the energy of a
    record where its first member is owned.  Against one GPU: forces <= 1e-12, every energy term, dPotential, a trajectory with
    box moves; a bond longer than the halo is reported, not skipped."""
    m = _with_bond_and_bend_lists(bilayer)
    n = m["nParticles"]
    assert len(m["molecules"][1]["bonds"]) > 300
    scale = [1.0005, 1.0005, 1.0 / (1.0005 * 1.0005)]
    mc = np.random.RandomState(9).random_sample(8)

    def run(sim, gather):
        sim.compute_forces(mask=sm.MASK_ALL, step=5)
        a = gather(sim)[3]
        U, dU = sim.potential(), sim.dpotential(scale)
        boxes = []
        for t in range(3):
            sim.step(8 * t, 8)
            boxes.append(sim.mc_box_move(0.01, 0.4, mc[2 * t], mc[2 * t + 1])[2])
        x = gather(sim)[0]
        return a, U, dU, np.array(boxes), x

    one = sm.Context.from_dict(m)
    a1, U1, dU1, b1, x1 = run(one, lambda c: (c.get_particles()[0], None, None, c.get_forces()))
    one.close()
    grp = LocalSlabGroup(m, nranks)
    a2, U2, dU2, b2, x2 = run(grp, lambda g: g.gather(n))
    grp.close()
    assert rel_err(a2, a1) <= 1e-12
    for t in (sm.TERM_PAIR, sm.TERM_CHAIN, sm.TERM_BOND, sm.TERM_BEND):
        assert U1[t] != 0 and abs(U2[t] - U1[t]) <= 1e-12 * abs(U1[t]), t
        assert abs(dU2[t] - dU1[t]) <= 1e-12 * abs(U1[t]) + 1e-9 * abs(dU1[t]), t
    assert np.allclose(b2, b1, rtol=1e-13) and np.abs(x2 - x1).max() <= 1e-9

    # the production call sequence: steps + trial in one call (smd_step_mc: fused seam, force + dPotential pass armed on every rank)
    one = sm.Context.from_dict(m)
    one.compute_forces(mask=sm.MASK_ALL, step=5)
    grp = LocalSlabGroup(m, nranks)
    grp.batched_default = True
    grp.compute_forces(mask=sm.MASK_ALL, step=5)
    for t in range(3):
        r1 = one.step_mc(8 * t, 8, 0.01, 0.4, mc[2 * t], mc[2 * t + 1])
        r2 = grp.step_mc(8 * t, 8, 0.01, 0.4, mc[2 * t], mc[2 * t + 1])
        assert r1[0] == r2[0] and np.allclose(r1[2], r2[2], rtol=1e-13)
    assert np.abs(grp.gather(n)[0] - one.get_particles()[0]).max() <= 1e-9
    one.close()
    grp.close()

    # a bond across half the box: its partner is in nobody's halo
    dx = m["xyz"][:, 0] - m["xyz"][0, 0]
    dx -= m["size"][0] * np.round(dx / m["size"][0])
    far_j = int(np.argmax(np.abs(dx)))          # half a box away along x: ten cell columns
    far = dict(m)
    far["molecules"] = list(m["molecules"][:1]) + [{"type": sm.MOL_BOND, "constants": np.array([1.0, 1.0]),
                                                    "bonds": np.array([[0, far_j]], np.int32)}]
    far["nMolecules"] = 2
    grp = LocalSlabGroup(far, 4)
    with pytest.raises(sm.SoftMoldError, match="neither owned nor inside the halo"):
        grp.compute_forces(mask=sm.MASK_ALL, step=0)
        grp.synchronize()
    for c in grp.ctx:
        c.close()


def test_slab_set_local_checks_run_on_the_device_and_name_the_fault(bilayer):
    """smd_slab_set_local: global index, type and position of every uploaded particle are checked by the import kernel (no
    host loop on the upload path); a refused upload leaves the rank usable"""
    m = bilayer
    grp = LocalSlabGroup(m, 2)
    grp.compute_forces()
    c = grp.ctx[0]
    g, x, t, v, _ = c.slab_get_local()
    g, x, t, v = g.copy(), x.copy(), t.copy(), v.copy()
    for what, match in (("gid", "global particle index out of range"), ("type", "particle type out of range"),
                        ("pos", "out of bounds")):
        g2, x2, t2 = g.copy(), x.copy(), t.copy()
        if what == "gid":
            g2[3] = m["nParticles"] + 5
        elif what == "type":
            t2[4] = m["nTypes"]
        else:
            x2[5, 1] = -0.5
        with pytest.raises(sm.SoftMoldError, match=match):
            c.slab_set_local(g2, x2, t2, v)
    for r, cc in enumerate(grp.ctx):                  # reload every rank with its own particles: the run goes on
        gg, xx, tt, vv, _ = (g, x, t, v, None) if r == 0 else cc.slab_get_local()
        cc.slab_set_local(gg.copy(), xx.copy(), tt.copy(), vv.copy())
    grp.compute_forces()
    grp.step(0, 4)
    grp.gather(m["nParticles"])
    grp.close()
