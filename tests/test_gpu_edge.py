"""Edge cases of the hot path on the GPU, against the oracle (pinned to the unmodified reference by the golden
fixtures, tests/test_oracle_golden.py) on fresh seeded inputs, against plain numpy where the reference itself cannot
run (more than MAX_CELL_SIZE = 512 particles in a cell, cellOpt.h:26-28), and -- at BASELINE.json's full size --
through size-independent properties.

Covered: a periodic gas whose every pair class occurs (all cells are boundary cells, non-cubic box, cells larger
than the cutoff), asymmetric constant tables, type-0 particles that are never integrated (Q4),
particles exactly on the faces of the box (Q5), several CHAIN blocks of different lengths, one- and two-particle
systems, an empty molecule list, an overfull cell, the 240 000-particle liposome."""
import numpy as np
import pytest

import softmold_b200 as sm
from softmold_b200 import workloads

pytestmark = pytest.mark.gpu


def gas(seed, n, box, n_types=3, symmetric=True, type0=0.1, chains=()):
    rng = np.random.default_rng(seed)
    umax = rng.uniform(50.0, 200.0, (n_types, n_types))
    umin = np.where(rng.random((n_types, n_types)) < 0.5, rng.uniform(-6.0, -1.0, (n_types, n_types)), 0.0)
    if symmetric:
        umax, umin = (umax + umax.T) / 2, np.minimum(umin, umin.T)
    fC, uC = workloads.pair_tables(n_types, umax.T.ravel(), umin.T.ravel())
    xyz = rng.random((n, 3)) * np.array(box)
    typ = rng.integers(1, n_types, n).astype(np.int32)
    typ[rng.random(n) < type0] = 0
    vel = rng.normal(0.0, 1.0, (n, 3))
    mols, at = [], 0
    for nch, ln in chains:      # chain members close together: a random walk of 0.7 steps
        for c in range(nch):
            for l in range(1, ln):
                step = rng.normal(0, 1, 3)
                xyz[at + c * ln + l] = np.mod(xyz[at + c * ln + l - 1] + 0.7 * step / np.linalg.norm(step), box)
        typ[at:at + nch * ln] = np.maximum(typ[at:at + nch * ln], 1)
        mols.append({"type": sm.MOL_CHAIN, "constants": np.array([0.7, 100.0, 1.0, 100.0]), "bonds": np.array([[at, nch, ln]], np.int32)})
        at += nch * ln
    return {"gamma": 1.0, "initialTemp": 3.0, "finalTemp": 3.0, "seed": seed, "nTypes": n_types, "nMolecules": len(mols),
            "nParticles": n, "periodic": 1, "cutoff": 2.0, "size": [float(b) for b in box], "initialTime": 0.0, "finalTime": 1.0,
            "deltaT": 0.02, "storeInterval": 1.0, "measureInterval": 1.0, "twoBodyFconst": fC, "twoBodyUconst": uC,
            "xyz": xyz, "type": typ, "vel": vel, "molecules": mols}


def check_against_oracle(orc, m, steps=5, scale=(1.0007, 0.9991, 1.0004)):
    ctx = sm.Context.from_dict(m)
    n, nT = m["nParticles"], m["nTypes"]
    # membership: cells and in-range pairs, bit-exact
    _, key, _ = ctx.get_cell_ids()
    assert np.array_equal(key, orc.cell_ids(m["xyz"], m["size"], m["cutoff"]))
    tot, per = ctx.count_pairs()
    otot, oper = orc.pair_count(m["xyz"], m["type"], nT, m["size"], m["cutoff"], per_particle=True)
    assert tot == otot and np.array_equal(per, oper)
    # pair terms
    ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    a, ra = ctx.get_forces(), orc.pair_force(m["xyz"], m["type"], nT, m["twoBodyFconst"], m["size"], m["cutoff"])
    assert np.abs(a - ra).max() <= 1e-12 * max(np.abs(ra).max(), 1e-300)
    U, rU = ctx.potential()[sm.TERM_PAIR], orc.pair_potential(m["xyz"], m["type"], nT, m["twoBodyUconst"], m["size"], m["cutoff"])
    assert abs(U - rU) <= 1e-12 * max(abs(rU), 1.0)
    dU, rdU = ctx.dpotential(scale)[sm.TERM_PAIR], orc.pair_dpotential(m["xyz"], m["type"], nT, m["twoBodyUconst"], m["size"], m["cutoff"], scale)
    assert abs(dU - rdU) <= 1e-12 * max(abs(rU), 1.0) + 1e-9 * abs(rdU)
    # the whole loop with the product's noise
    S = orc.System(m, noise="philox")
    S.init(0)
    ctx.compute_forces(mask=sm.MASK_ALL, step=0)
    assert np.abs(ctx.get_forces() - S.acc).max() <= 1e-11 * max(np.abs(S.acc).max(), 1e-300)
    ctx.step(0, steps)
    for i in range(steps):
        S.step(i)
    xyz, typ, vel = ctx.get_particles()
    assert np.array_equal(typ, m["type"])
    assert np.abs(xyz - S.xyz).max() <= 1e-9 and np.abs(vel - S.vel).max() <= 1e-8
    frozen = m["type"] == 0
    if frozen.any():   # Q4: type 0 is never integrated (verlet.h:296,308,470)
        assert np.array_equal(xyz[frozen], m["xyz"][frozen]) and np.array_equal(vel[frozen], m["vel"][frozen])
    ctx.close()


@pytest.mark.parametrize("symmetric", [True, False])
def test_periodic_gas_noncubic_box(orc, symmetric):
    # nc = (7, 4, 6): cells of 2.04 x 2.28 x 2.12 > rc; every cell touches its own periodic image's neighbours in y
    m = gas(11 if symmetric else 12, 3000, (14.3, 9.1, 12.7), symmetric=symmetric, chains=((150, 4), (100, 3), (40, 7)))
    check_against_oracle(orc, m)


def test_three_cells_per_axis(orc):
    # the smallest periodic grid the reference's 27-cell stencil can use: every neighbour is reached through an image
    m = gas(13, 700, (6.5, 6.2, 7.9), n_types=2, chains=((60, 3),))
    check_against_oracle(orc, m, steps=6)


def test_particles_on_the_faces_of_the_box(orc):
    m = gas(14, 1500, (12.0, 10.0, 8.0), type0=0.0)
    box = np.array(m["size"])
    rng = np.random.default_rng(5)
    for k in range(120):      # exactly 0 and exactly L on every axis (Q5: the wrap is strict, build() clamps the index)
        m["xyz"][k, k % 3] = 0.0 if (k // 3) % 2 == 0 else box[k % 3]
    m["vel"][:120] = rng.normal(0, 0.2, (120, 3))
    check_against_oracle(orc, m, steps=4)


def test_tiny_systems(orc):
    base = gas(15, 2, (8.0, 8.0, 8.0), n_types=2, type0=0.0)
    for xyz in ([[1.0, 1.0, 1.0], [2.2, 1.4, 0.8]],          # in range
                [[0.3, 4.0, 4.0], [7.6, 4.1, 3.9]],          # in range through the periodic image
                [[1.0, 1.0, 1.0], [5.0, 5.0, 5.0]]):         # out of range
        m = dict(base, xyz=np.array(xyz))
        check_against_oracle(orc, m, steps=5)
    one = dict(base, nParticles=1, xyz=base["xyz"][:1].copy(), vel=base["vel"][:1].copy(), type=base["type"][:1].copy())
    ctx = sm.Context.from_dict(one)
    ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    assert np.array_equal(ctx.get_forces(), np.zeros((1, 3)))
    assert ctx.potential()[sm.TERM_PAIR] == 0.0 and ctx.count_pairs()[0] == 0
    ctx.step(0, 3)
    ctx.synchronize()
    ctx.close()


def test_overfull_cell(orc):
    """4 600 particles in ONE cell: more than the 12-bit offset of a list entry addresses (the excess is taken one by
    one) and nine times the reference's own MAX_CELL_SIZE, so the check is a numpy all-pairs sum"""
    n, box = 4600, np.array([12.0, 12.0, 12.0])
    rng = np.random.default_rng(16)
    m = gas(16, n, box, n_types=2, type0=0.0)
    m["xyz"] = 4.02 + rng.random((n, 3)) * 1.96           # cell (2,2,2) of the 6 x 6 x 6 grid
    m["xyz"][:40] = rng.random((40, 3)) * box             # and a few elsewhere
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    a = ctx.get_forces()
    U = ctx.potential()[sm.TERM_PAIR]
    ctx.close()
    fC, uC, nT, t = m["twoBodyFconst"].reshape(-1, 6), m["twoBodyUconst"].reshape(-1, 6), 2, m["type"]
    ra, rU = np.zeros((n, 3)), 0.0
    for i0 in range(0, n, 400):
        d = m["xyz"][i0:i0 + 400, None, :] - m["xyz"][None, :, :]
        d -= box * np.round(d / box)
        r = np.sqrt((d ** 2).sum(-1))
        row = t[i0:i0 + 400, None] * nT + t[None, :]
        inr = (r < 2.0) & (r > 0)
        rs = np.where(inr, r, 1.0)
        core = rs < fC[row, 0]
        mm = np.where(core, fC[row, 0], fC[row, 3]) - rs
        mag = np.where(core, fC[row, 1] - fC[row, 2] * mm, fC[row, 4] - fC[row, 5] * mm) * mm / rs
        ra[i0:i0 + 400] = (d * np.where(inr, mag, 0.0)[..., None]).sum(1)
        uc = uC[row, 1] * (uC[row, 0] - rs) ** 2 + uC[row, 2]
        ut = (uC[row, 3] - rs) ** 2 * (uC[row, 4] - (uC[row, 3] - rs) * uC[row, 5])
        rU += 0.5 * np.where(inr, np.where(rs <= uC[row, 0], uc, ut), 0.0).sum()
    assert np.abs(a - ra).max() <= 1e-10 * np.abs(ra).max()
    assert abs(U - rU) <= 1e-10 * abs(rU)


def test_full_size_properties():
    """BASELINE configs[1] (liposome, 80 000 lipids, N = 240 000): Newton's third law, the two energy kernels, an
    identity box move, the neighbour census, and a thermostatted run that stays at its temperature"""
    m = workloads.liposome(80000)
    n = m["nParticles"]
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces(mask=sm.MASK_ALL, step=0)
    ctx.step(0, 300)
    ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    a = ctx.get_forces()
    assert np.abs(a.sum(0)).max() <= 1e-9 * np.abs(a).sum()          # every pair term appears once with each sign
    ctx.compute_forces(mask=1 << sm.TERM_CHAIN)
    c = ctx.get_forces()
    assert np.abs(c.sum(0)).max() <= 1e-9 * np.abs(c).sum()
    tot, per = ctx.count_pairs()
    assert per.sum() == 2 * tot and 20 < tot / n < 40
    U = ctx.potential()
    assert ctx.dpotential([1.0, 1.0, 1.0])[sm.TERM_PAIR] == 0.0 and ctx.dpotential([1.0, 1.0, 1.0])[sm.TERM_CHAIN] == 0.0
    s = 1.0 + 1e-6
    dU = ctx.dpotential([s, s, s])   # a uniform dilation: dU = U(x) - U(s x); to first order the virial -(s-1) sum r.dU/dr
    assert abs(dU[sm.TERM_PAIR]) < 1e-3 * abs(U[sm.TERM_PAIR]) and dU[sm.TERM_PAIR] != 0.0
    T = 2.0 * ctx.kinetic() / (3.0 * n)
    assert abs(T - 3.0) < 0.1, T
    xyz, _, _ = ctx.get_particles()
    assert (xyz >= 0).all() and (xyz <= np.array(m["size"])).all()
    ctx.close()


def test_hundred_beads_through_the_cell_grid(orc):
    """A vesicle studded with 100 continuum-sphere beads in one BEAD molecule plus 30 in a second one (bead lists assembled
    from molecules j >= i, system.h:2053-2070), in a box small enough that bead cubes wrap around it: k_bead visits only the
    cells within R + rc of each bead.  Forces, energies, dPotential and a short trajectory against the oracle, whose > 20
    bead rule is pinned to the reference by tests/golden/bead24.npz."""
    from conftest import golden_path
    m, _ = orc.load_golden(golden_path("bead1"))
    b = int(m["molecules"][1]["bonds"][0, 0])
    lip = np.arange(m["nParticles"]) != b
    c = m["xyz"][lip].mean(0)
    rng = np.random.default_rng(100)
    extra = []
    for _ in range(100000):     # (bounded: 129 spheres 3.6 apart fit between 0.97 and 1.5 vesicle radii in ~1 300 draws)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        x = c + (q @ (m["xyz"][b] - c)) * rng.uniform(0.97, 1.5)
        if np.linalg.norm(np.array([m["xyz"][b]] + extra) - x, axis=1).min() > 3.6 and np.all(x > 1) and np.all(x < np.array(m["size"]) - 1):
            extra.append(x)
        if len(extra) == 129:
            break
    assert len(extra) == 129
    n0 = m["nParticles"]
    m = dict(m, xyz=np.vstack([m["xyz"], extra]), vel=np.vstack([m["vel"], rng.normal(0, 0.1, (129, 3))]),
             type=np.append(m["type"], np.full(129, m["type"][b])).astype(np.int32), nParticles=n0 + 129)
    own1 = np.array([[b]] + [[n0 + k] for k in range(99)], np.int32)
    own2 = np.array([[n0 + k] for k in range(99, 129)], np.int32)
    C = m["molecules"][1]["constants"]
    m["molecules"] = [m["molecules"][0], {"type": sm.MOL_BEAD, "constants": C, "bonds": own1},
                      {"type": sm.MOL_BEAD, "constants": C, "bonds": own2}]
    m["nMolecules"] = 3
    # shrink the box around the vesicle so that the cubes of cells around outer beads wrap across the periodic faces
    lo, hi = m["xyz"].min(0) - 1.5, m["xyz"].max(0) + 1.5
    m["xyz"] = m["xyz"] - lo
    m["size"] = [float(v) for v in (hi - lo)]
    ctx = sm.Context.from_dict(m)
    S = orc.System(m, noise="philox")
    S.init(7)
    ctx.compute_forces(mask=sm.MASK_ALL, step=7)
    a = ctx.get_forces()
    assert np.abs(a - S.acc).max() <= 1e-11 * np.abs(S.acc).max()
    ctx.compute_forces(mask=1 << sm.TERM_BEAD)
    ab = ctx.get_forces()
    assert np.abs(ab).max() > 10.0                     # the beads really press on the membrane and on each other
    U, Uo = ctx.potential(), S.potential()
    assert abs(U.sum() - Uo) <= 1e-11 * abs(Uo) and U[sm.TERM_BEAD] != 0.0
    scale = (0.9991, 0.9991, 1.0 / 0.9991 ** 2)
    dU, dUo = ctx.dpotential(scale), S.dpotential(scale)
    assert abs(dU.sum() - dUo) <= 1e-11 * abs(Uo)
    ctx.compute_forces(mask=sm.MASK_ALL, step=7)       # (a[] held the bead term alone)
    ctx.step(7, 6)
    for i in range(6):
        S.step(7 + i)
    xyz, _, vel = ctx.get_particles()
    assert np.abs(xyz - S.xyz).max() <= 1e-9 and np.abs(vel - S.vel).max() <= 1e-7
    ctx.close()


def test_particles_that_outrun_the_occupied_window(orc):
    """The kernel that moves the particles also fills the histogram of the next cell build, under a window published by the
    previous build (occupied box + 1 cell).  Particles that leave that window in one step -- here a few fired at 150 sigma
    per time unit out of a small cluster in a large box, some of them across the periodic faces -- mark it dirty and the
    build redoes its histogram (k_scan's fallback): trajectory equal to the oracle's, cells bit-exact."""
    rng = np.random.default_rng(5)
    m = gas(21, 1500, (60.0, 60.0, 60.0), n_types=3, type0=0.0, chains=((100, 3),))
    m["xyz"] = np.mod(m["xyz"] * 0.2 + np.array([1.0, 24.0, 47.5]), 60.0)     # a 12-sigma cluster touching the x = 0 face region
    m["vel"][:] = rng.normal(0, 1.0, m["vel"].shape)
    fast = rng.choice(1500, 12, replace=False)
    m["vel"][fast] = rng.normal(0, 1.0, (12, 3)) * 150.0
    ctx = sm.Context.from_dict(m)
    S = orc.System(m, noise="philox")
    S.init(3)
    ctx.compute_forces(mask=sm.MASK_ALL, step=3)
    for k in range(3):       # the fused multi-step path and the step_begin / step_end path
        ctx.step(3 + 4 * k, 3)
        ctx.step_begin(6 + 4 * k); ctx.step_end(6 + 4 * k)
        for i in range(4):
            S.step(3 + 4 * k + i)
        xyz, _, vel = ctx.get_particles()
        assert np.abs(xyz - S.xyz).max() <= 1e-9 and np.abs(vel - S.vel).max() <= 1e-8
        _, key, _ = ctx.get_cell_ids()
        assert np.array_equal(key, orc.cell_ids(xyz, m["size"], m["cutoff"]))
    ctx.close()


def test_device_timeline_of_a_step(orc):
    """smd_timeline / smd_timeline_read: %globaltimer stamps inside the kernels of one step.  The order of the build kernels, the
    pair kernel starting after the reorder ended, the seam ending last; switching it off again stamps nothing new."""
    from softmold_b200 import workloads
    m = workloads.liposome(3000, 3.45, 2)
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces(step=0)
    ctx.step(0, 20)
    ctx.timeline(True)
    ctx.step(20, 6)
    t = ctx.timeline_read()
    assert t["k_scan"][0] == 0.0
    for k in t:
        assert 0.0 <= t[k][0] <= t[k][1] <= t[k][2] < 1e4, (k, t[k])
    assert t["k_scan"][2] <= t["k_place"][2] <= t["k_reorder"][2] <= t["k_pair_force2"][2] <= t["k_chain_kick"][2]
    assert t["k_pair_force2"][0] >= t["k_reorder"][2] - 0.5           # (stamped after its griddepcontrol.wait)
    ctx.timeline(False)
    ctx.step(26, 6)
    assert ctx.timeline_read() == t
    ctx.close()


@pytest.mark.parametrize("engine", ["0", "1"])
def test_many_particle_types_need_the_shared_memory_opt_in(orc, engine):
    """26 particle types: the staged pair tables (80 B per ordered type pair) push the dynamic shared memory of both pair
    engines past 48 KB, the size that needs cudaFuncAttributeMaxDynamicSharedMemorySize -- with one (SMD_PAIR3=0) and with
    three threads per particle"""
    import os
    m = gas(23, 2500, (14.3, 11.1, 12.7), n_types=26, chains=((100, 3),))
    os.environ["SMD_PAIR3"] = engine
    try:
        check_against_oracle(orc, m, steps=4)
    finally:
        del os.environ["SMD_PAIR3"]
