"""GPU parity tests proper: the CUDA path, called through the C ABI, against
  (1) the golden vectors produced by the UNMODIFIED reference (tests/golden/*.npz), and
  (2) the oracle (oracle/oracle.c) on the same seeded inputs.
Tolerances are the north star's: forces <= 1e-5 relative, energies <= 1e-7 relative (FP64), cell / neighbour
membership bit-exact.  In practice the pair terms are bit-identical and only the summation order differs, so the
observed errors sit around 1e-14; the asserts below use the stated bounds with an extra tight check where the
design promises more."""
import numpy as np
import pytest

import softmold_b200 as sm
from conftest import CASES, golden_path

pytestmark = pytest.mark.gpu

F_RTOL = 1e-5   # per-particle force, relative to max|F| of the term
E_RTOL = 1e-7   # energies


def rel_force_err(a, ref):
    scale = np.abs(ref).max()
    return np.abs(a - ref).max() / (scale if scale > 0 else 1.0)


def close_energy(x, ref, scale=None):
    s = max(abs(ref), abs(scale) if scale is not None else 0.0, 1e-300)
    return abs(x - ref) / s


@pytest.fixture(scope="module")
def contexts(orc):
    cache = {}

    def get(case, **kw):
        key = (case, tuple(sorted(kw.items())))
        if key not in cache:
            m, ref = orc.load_golden(golden_path(case))
            cache[key] = (sm.Context.from_dict(m, **kw), m, ref)
        return cache[key]

    yield get
    for ctx, _, _ in cache.values():
        ctx.close()


@pytest.mark.parametrize("case", CASES)
def test_cell_membership_bit_exact(contexts, case):
    ctx, m, ref = contexts(case)
    nc, key, rank = ctx.get_cell_ids()
    assert list(nc) == list(ref["nCells_nFull"][:3])
    assert np.array_equal(key, ref["cell_id"])                       # cellOpt.h:530-556
    # list order of every cell = the reference's head-inserted linked list (descending index)
    nxt = ref["cell_next"]
    n = len(key)
    expect_rank = np.zeros(n, np.int32)
    heads = {}
    for i in range(n):          # head of a cell = largest index in it
        heads[key[i]] = i
    for c, h in heads.items():
        r, i = 0, h
        while i != -1:
            expect_rank[i] = r
            r += 1
            i = nxt[i]
    assert np.array_equal(rank, expect_rank)
    assert len(heads) == ref["nCells_nFull"][3]


@pytest.mark.parametrize("case", CASES)
def test_neighbour_membership_bit_exact(contexts, orc, case):
    ctx, m, ref = contexts(case)
    tot, per = ctx.count_pairs()
    otot, oper = orc.pair_count(m["xyz"], m["type"], m["nTypes"], m["size"], m["cutoff"], per_particle=True)
    assert tot == otot
    assert np.array_equal(per, oper)


@pytest.mark.parametrize("case", CASES)
def test_pair_force_potential_dpotential(contexts, case):
    ctx, m, ref = contexts(case)
    ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    a = ctx.get_forces()
    ra = ref["a_pair"].reshape(-1, 3)
    err = rel_force_err(a, ra)
    assert err <= F_RTOL
    assert err <= 1e-12, err            # identical pair terms, only the summation order differs
    U = ctx.potential()
    assert close_energy(U[sm.TERM_PAIR], ref["U_pair"][0]) <= E_RTOL
    assert close_energy(U[sm.TERM_PAIR], ref["U_pair"][0]) <= 1e-12
    dU = ctx.dpotential(ref["scale"])
    # dU is a sum of many cancelling terms: the natural scale is the potential it is a difference of
    assert close_energy(dU[sm.TERM_PAIR], ref["dU_pair"][0]) <= 1e-7
    assert abs(dU[sm.TERM_PAIR] - ref["dU_pair"][0]) <= 1e-12 * abs(ref["U_pair"][0]) + 1e-9 * abs(ref["dU_pair"][0])
    assert close_energy(ctx.kinetic(), ref["kinetic"][0]) <= 1e-13


def test_known_answers_of_the_survey_at_baseline_size_c1(orc):
    """SURVEY.md 8(c)'s KAT table (`liposome kat5000 5000 5000 3.45`, N = 15 000 = BASELINE config C1 at full size), FP64"""
    from conftest import KAT5000 as K
    m, _ = orc.load_golden(golden_path("kat5000"))
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces(mask=1 << sm.TERM_PAIR)
    a = ctx.get_forces()
    assert np.abs(a[0] - np.array(K["a_pair0"])).max() <= 1e-12 * K["maxF"]
    assert abs((a * a).sum() - K["sumF2"]) <= 1e-12 * K["sumF2"]
    assert abs(np.sqrt((a * a).sum(1)).max() - K["maxF"]) <= 1e-12 * K["maxF"]
    U = ctx.potential()
    assert abs(U[sm.TERM_PAIR] - K["U_pair"]) <= 1e-12 * abs(K["U_pair"])
    dU = ctx.dpotential(K["scale"])
    assert abs(dU[sm.TERM_PAIR] - K["dU_pair"]) <= 1e-12 * abs(K["U_pair"])
    assert abs(dU[sm.TERM_CHAIN] - K["dU_chain"]) <= 1e-9 * abs(K["dU_chain"])
    ctx.close()


TERM_OF = {sm.MOL_CHAIN: sm.TERM_CHAIN, sm.MOL_BOND: sm.TERM_BOND, sm.MOL_BEND: sm.TERM_BEND, sm.MOL_BEAD: sm.TERM_BEAD,
           sm.MOL_BALL: sm.TERM_BALL, sm.MOL_BOUNDARY: sm.TERM_FIELD, sm.MOL_FLOATING_BASE: sm.TERM_FIELD,
           sm.MOL_ZTORQUE: sm.TERM_FIELD, sm.MOL_ZPOWERPOTENTIAL: sm.TERM_FIELD, sm.MOL_NANOCORE: sm.TERM_NANOCORE}


@pytest.mark.parametrize("case", CASES)
def test_molecule_terms(contexts, case):
    ctx, m, ref = contexts(case)
    U, dU = ctx.potential(), ctx.dpotential(ref["scale"])
    by_term_a, by_term_U, by_term_dU = {}, {}, {}
    for k, mol in enumerate(m["molecules"]):
        if mol["type"] in sm.MOL_IGNORED:      # SOLID, OFFSET_BOUNDARY, ...: `MD` does nothing with them, nor does the reference harness
            assert not ref[f"a_mol{k}"].any() and ref["U_mol"][k] == 0
            continue
        t = TERM_OF[mol["type"]]
        by_term_a[t] = by_term_a.get(t, 0) + ref[f"a_mol{k}"].reshape(-1, 3)
        by_term_U[t] = by_term_U.get(t, 0) + ref["U_mol"][k]
        by_term_dU[t] = by_term_dU.get(t, 0) + ref["dU_mol"][k]
    for t in by_term_a:
        ctx.compute_forces(mask=1 << t)
        a = ctx.get_forces()
        err = rel_force_err(a, by_term_a[t])
        assert err <= F_RTOL, (case, t, err)
        assert err <= 1e-11, (case, t, err)
        assert close_energy(U[t], by_term_U[t], scale=ref["U_pair"][0] * 1e-6) <= E_RTOL, (case, t, U[t], by_term_U[t])
        assert abs(dU[t] - by_term_dU[t]) <= 1e-7 * max(abs(by_term_dU[t]), 1e-6 * abs(by_term_U[t]), 1e-12), (case, t)


@pytest.mark.parametrize("case", CASES)
def test_total_force_matches_oracle_with_philox_noise(contexts, orc, case):
    """all terms + Langevin with the counter-based noise, against the oracle's restatement of the same spec"""
    ctx, m, ref = contexts(case)
    step = 12345
    ctx.compute_forces(mask=sm.MASK_ALL, step=step)
    a = ctx.get_forces()
    S = orc.System(m, noise="philox")
    S.init(step)
    err = rel_force_err(a, S.acc)
    assert err <= 1e-11, err


def test_philox_uniforms_bit_exact(orc):
    """a = -g v + sigma (2u-1) with v = 0 and only the Langevin term isolates the uniforms"""
    n = 4096
    rng = np.random.default_rng(3)
    box = [40.0, 40.0, 40.0]
    xyz = rng.random((n, 3)) * 40
    typ = np.ones(n, np.int32)
    gamma, dt, T, seed = 1.0, 0.02, 3.0, 0x123456789ABCDEF
    ctx = sm.Context(n, 2, box, 2.0, dt, gamma, T, seed)
    ctx.set_pair_tables(np.zeros(24), np.zeros(24))
    ctx.set_particles(xyz, typ, np.zeros((n, 3)))
    for step in (0, 7, 2 ** 33 + 5):
        ctx.compute_forces(mask=sm.MASK_LANGEVIN, step=step)
        a = ctx.get_forces()
        u = orc.philox_uniforms(seed, step, n)
        sigma = np.sqrt((6.0 * T * gamma) / dt)
        expect = -gamma * 0.0 + sigma * (2.0 * u - 1.0)
        assert np.array_equal(a, expect)
    ctx.close()
    # sanity of the stream itself
    u = orc.philox_uniforms(seed, 1, 200000)
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 1e-3 and u.min() >= 0 and u.max() < 1


@pytest.mark.parametrize("case", ["lipo_t0", "lipo_eq", "bilayer_t0", "bilayer_eq", "bead1", "bead2", "bead24", "ball", "fields", "kat5000"])
def test_trajectory_matches_reference_md(contexts, orc, case):
    """The whole loop of MD.cpp for K steps against the state the reference `MD` executable wrote (one thread):
    Langevin noise = the reference's MT19937 stream fed through smd_set_noise, MC box moves with tension driven by
    the reference's second MT19937 stream, bead mass quirk, restart path."""
    m, ref = orc.load_golden(golden_path(case))
    ctx = sm.Context.from_dict(m, noise=sm.NOISE_EXTERNAL)
    n, K = m["nParticles"], int(ref["traj_steps"])
    lang = orc.mt_rand53(m["seed"], 3 * n * (K + 2)).reshape(-1, n, 3)
    mc = orc.mt_rand53(m["seed"], 2 * (K // 8 + 2))
    start = int(m["initialTime"] / m["deltaT"] + 1e-7)
    draw = 0
    ctx.set_noise(lang[draw]); draw += 1
    ctx.compute_forces(mask=sm.MASK_ALL, step=start)
    if m["initialTime"] != 0:
        ctx.resume()
    trials = 0
    for i in range(start, start + K):
        ctx.step_begin(i)
        ctx.set_noise(lang[draw]); draw += 1
        ctx.step_end(i)
        if i % 8 == 0 and i != 0 and m.get("deltaLXY", 0) != 0:
            ctx.mc_box_move(m["deltaLXY"], m.get("tension", 0.0), mc[2 * trials], mc[2 * trials + 1])
            trials += 1
    ctx.step_begin(start + K)
    xyz, _, vel = ctx.get_particles()
    box = ctx.get_box()
    ctx.close()
    np.testing.assert_allclose(box, ref["traj_box"], rtol=1e-13)
    # 15 significant digits of text in the reference's file + chaotic growth of the 1e-16 summation-order noise
    assert np.abs(xyz - ref["traj_xyz"]).max() <= 1e-9
    assert np.abs(vel - ref["traj_vel"]).max() <= 1e-8


def test_trajectory_vs_oracle_philox_long(orc):
    """100 steps with the product's own noise against the oracle running the same Philox spec"""
    m, _ = orc.load_golden(golden_path("bilayer_eq"))
    ctx = sm.Context.from_dict(m)
    S = orc.System(m, noise="philox")
    start = 0
    m["initialTime"] = 0.0
    S.init(start)
    ctx.compute_forces(mask=sm.MASK_ALL, step=start)
    K = 100
    ctx.step(start, K)
    for i in range(start, start + K):
        S.s.deltaLXY = 0.0
        S.step(i)
    xyz, _, vel = ctx.get_particles()
    assert np.abs(xyz - S.xyz).max() <= 1e-7
    assert np.abs(vel - S.vel).max() <= 1e-6
    launches, rebuilds = ctx.stats()
    assert rebuilds >= K and launches > 0
    ctx.close()


def test_errors_are_loud(orc):
    m, _ = orc.load_golden(golden_path("lipo_t0"))
    bad = dict(m)
    bad["xyz"] = m["xyz"].copy()
    bad["xyz"][5, 0] = 401.0
    with pytest.raises(sm.SoftMoldError, match="X position of particle 5 is out of bounds"):
        sm.Context.from_dict(bad)
    ctx = sm.Context.from_dict(m)
    # (the load-time checks run inside the import kernel: the FIRST offender is named, a refused upload leaves the context
    # usable, and the next good upload works)
    bad["xyz"][2, 2] = float("nan")
    with pytest.raises(sm.SoftMoldError, match="Z position of particle 2 is out of bounds"):
        ctx.set_particles(bad["xyz"], m["type"], m["vel"])
    t = m["type"].copy()
    t[7] = m["nTypes"]
    with pytest.raises(sm.SoftMoldError, match="particle type out of range"):
        ctx.set_particles(m["xyz"], t, m["vel"])
    with pytest.raises(sm.SoftMoldError, match="smd_set_particles first"):
        ctx.compute_forces()
    ctx.set_particles(m["xyz"], m["type"], m["vel"])
    ctx.compute_forces()
    ctx.synchronize()
    with pytest.raises(sm.SoftMoldError):
        ctx.add_molecule(sm.MOL_BOND, np.array([[0, 10 ** 6]], np.int32), [1.0, 1.0])
    # a particle crossing many cells in one step is fine (cells are rebuilt from scratch, as in the reference) ...
    xyz, typ, vel = ctx.get_particles()
    vel[0] = [3000.0, 0, 0]
    ctx.set_particles(xyz, typ, vel)
    ctx.compute_forces()
    ctx.step(0, 1)
    ctx.synchronize()
    # ... but one that leaves the box even after the single wrap of Verlet::first makes the reference index out of
    # bounds (cellOpt.h:541-552); here it is a loud error
    vel[0] = [1e5, 0, 0]
    ctx.set_particles(xyz, typ, vel)
    ctx.compute_forces()
    ctx.step(0, 2)
    with pytest.raises(sm.SoftMoldError, match="outside the box"):
        ctx.synchronize()
    ctx.close()


def test_two_phase_energy_kernels_match_one_phase(orc):
    """potential / dPotential through the two-phase kernel (every unordered pair once, FP32 prefilter widened by the
    scaling) against the one-phase half-stencil kernels, on a membrane that spans the periodic box, for box moves far
    larger than the 0.01 steps of MD.cpp"""
    import os
    from softmold_b200 import workloads
    m = workloads.bilayer(4000, 3.11, seed=3)
    ctx = sm.Context.from_dict(m)
    ctx.compute_forces()
    ctx.step(0, 50)
    m = dict(m)
    m["xyz"], m["type"], m["vel"] = ctx.get_particles()
    scales = [[1.0005, 1.0005, 1.0 / 1.0005 ** 2], [0.99, 0.99, 1.0 / 0.99 ** 2], [1.03, 1.03, 1.0 / 1.03 ** 2], [1.0, 1.0, 1.0]]
    two = [ctx.potential()] + [ctx.dpotential(sc) for sc in scales]
    ctx.close()
    os.environ["SMD_ENERGY_ONEPHASE"] = "1"
    try:
        ctx = sm.Context.from_dict(m)
        one = [ctx.potential()] + [ctx.dpotential(sc) for sc in scales]
        ctx.close()
    finally:
        del os.environ["SMD_ENERGY_ONEPHASE"]
    U = abs(one[0][sm.TERM_PAIR])
    for a, b in zip(two, one):
        assert abs(a[sm.TERM_PAIR] - b[sm.TERM_PAIR]) <= 1e-12 * U, (a[sm.TERM_PAIR], b[sm.TERM_PAIR])
    assert two[-1][sm.TERM_PAIR] == 0.0


@pytest.mark.parametrize("case", ["lipo_eq", "bilayer_eq", "lipocyto_chains"])
def test_fused_step_kernel_is_bit_identical_to_separate_kernels(orc, case):
    """smd_step fuses chain forces + Verlet::second + the next Verlet::first into one per-particle kernel for CHAIN-only
    systems (SMD_NO_FUSE=1: separate kernels; SMD_PDL: how the kernels of the step are chained); the trajectory must
    not change by a single bit"""
    import os
    if case == "lipocyto_chains":      # two CHAIN molecules of different length (3 and 10), explicit BOND list dropped
        m, _ = orc.load_golden(golden_path("lipocyto_eq"))
        m = dict(m, molecules=[mol for mol in m["molecules"] if mol["type"] == sm.MOL_CHAIN])
    else:
        m, _ = orc.load_golden(golden_path(case))
    out = []
    # (SMD_PDL=0: plain stream order; 2: programmatic dependent launches along every kernel of the step, not only the seam)
    envs = ({"SMD_PDL": "1"}, {"SMD_PDL": "0"}, {"SMD_PDL": "2"}, {"SMD_NO_FUSE": "1"})
    for env in envs:
        os.environ.update(env)
        try:
            ctx = sm.Context.from_dict(m, track_unwrapped=True)
        finally:
            for k in env:
                del os.environ[k]
        ctx.compute_forces(step=5)
        ctx.step(5, 33)
        ctx.step(38, 1)
        ctx.step(39, 7)
        out.append(ctx.get_particles() + (ctx.get_forces(), ctx.get_unwrapped(), ctx.stats()[0]))
        ctx.close()
    x1, _, v1, a1, u1, l1 = out[-1]                 # separate kernels
    for x, _, v, a, u, l in out[:-1]:
        assert np.array_equal(x, x1) and np.array_equal(v, v1) and np.array_equal(a, a1) and np.array_equal(u, u1)
    assert out[0][5] < l1                           # launches: one seam kernel < separate kernels


@pytest.mark.parametrize("case", ["bondbend", "lipocyto_eq", "ball", "fields", "bead1", "bead2"])
def test_fused_step_kernel_with_every_molecule_kind(orc, case):
    """systems with BOND / BEND / BALL / BEAD / NANOCORE molecules or one-body fields next to (or without) CHAIN blocks also
    take the fused step seam: their kernels scatter into a[] before it, and the seam divides the continuum-sphere particles
    by their mass (once before each half kick for BEAD, quirk Q3; once for NANOCORE).  Only the order in which a
    particle's terms are added differs from the separate kernels: the trajectories agree to rounding."""
    import os
    m, _ = orc.load_golden(golden_path(case))
    m = dict(m, initialTime=0.0)
    out = []
    for env in ("0", "1"):
        os.environ["SMD_NO_FUSE"] = env
        try:
            ctx = sm.Context.from_dict(m, track_unwrapped=True)
        finally:
            del os.environ["SMD_NO_FUSE"]
        ctx.compute_forces(step=0)
        # the bead fixtures start with the sphere pressed into the membrane and run away after ~9 steps (in the oracle too)
        k1 = 3 if case.startswith("bead") else 24
        ctx.step(0, k1)
        ctx.step(k1, 1)
        ctx.step(k1 + 1, 3 if case.startswith("bead") else 6)
        out.append(ctx.get_particles() + (ctx.get_forces(), ctx.get_unwrapped(), ctx.stats()[0]))
        ctx.close()
    (x0, _, v0, a0, u0, l0), (x1, _, v1, a1, u1, l1) = out
    assert l0 < l1                       # the fused path did run
    assert np.abs(x0 - x1).max() <= 1e-9 and np.abs(v0 - v1).max() <= 1e-8 and np.abs(u0 - u1).max() <= 1e-9
    assert rel_force_err(a0, a1) <= 1e-9


@pytest.mark.parametrize("case", ["bilayer_eq", "lipo_eq", "lipocyto_eq", "fields"])
def test_step_mc_equals_step_then_box_move(orc, case):
    """smd_step_mc: the pair kernel of the last step also sums the dPotential of the proposed box move (k_pair_force2 EMODE
    3) instead of a second pass over the pairs.  Against smd_step + smd_mc_box_move and against SMD_NO_DU_FUSE=1: the same
    decisions, dU equal to rounding, forces of the fused pass bit-identical -- so the trajectories stay identical bit for
    bit as long as the decisions agree."""
    import os
    m, _ = orc.load_golden(golden_path(case))
    m = dict(m, initialTime=0.0)
    dl, tension = 0.01, 0.4
    mc = orc.mt_rand53(7, 64)
    runs = []
    for mode in ("two_calls", "fused", "unfused_env"):
        if mode == "unfused_env":
            os.environ["SMD_NO_DU_FUSE"] = "1"
        try:
            ctx = sm.Context.from_dict(m)
        finally:
            os.environ.pop("SMD_NO_DU_FUSE", None)
        ctx.compute_forces(step=0)
        log = []
        for t in range(6):
            if mode == "two_calls":
                ctx.step(8 * t, 8)
                acc, dU, box = ctx.mc_box_move(dl, tension, mc[2 * t], mc[2 * t + 1])
            else:
                acc, dU, box = ctx.step_mc(8 * t, 8, dl, tension, mc[2 * t], mc[2 * t + 1])
            log.append((acc, dU, tuple(box)))
        runs.append((log, ctx.get_particles(), ctx.get_forces(), ctx.stats()[0], abs(ctx.potential()[sm.TERM_PAIR])))
        ctx.close()
    (l0, p0, a0, n0, U0), (l1, p1, a1, n1, _), (l2, p2, a2, n2, _) = runs
    assert [x[0] for x in l0] == [x[0] for x in l1] == [x[0] for x in l2] and any(x[0] for x in l0)
    exact = case in ("bilayer_eq", "lipo_eq")     # list molecules add their forces with FP64 atomics: not reproducible to the bit
    for (_, d0, b0), (_, d1, b1), (_, d2, b2) in zip(l0, l1, l2):
        assert abs(d0 - d1) <= 1e-12 * U0 and abs(d0 - d2) <= 1e-12 * U0 and np.allclose(b0, b1, rtol=1e-15) and np.allclose(b0, b2, rtol=1e-15)
        assert not exact or (d0 == d2 and b0 == b1 == b2)
    if exact:
        assert np.array_equal(p0[0], p1[0]) and np.array_equal(p0[2], p1[2]) and np.array_equal(a0, a1)
        assert np.array_equal(p0[0], p2[0]) and np.array_equal(a0, a2)
    else:
        assert np.abs(p0[0] - p1[0]).max() <= 1e-9 and np.abs(p0[0] - p2[0]).max() <= 1e-9 and rel_force_err(a1, a0) <= 1e-9
    if case != "fields":                 # (BEAD / NANOCORE-free systems take the fused step path, which smd_step_mc needs)
        assert n1 < n0                   # fewer launches: the dPotential kernel is gone


def test_x_sliced_sort_changes_nothing_but_the_summation_order(orc):
    """Geom::xs: the sort key splits every reference cell into x slices so that phase 1 of the pair kernel reads only the
    slices within reach.  Against SMD_XSUB=1 (cell-sorted only): identical cell ids and linked-list ranks, identical
    in-range pair counts per particle, forces / energies / a trajectory equal to rounding."""
    import os
    from softmold_b200 import workloads
    out = []
    for m in (workloads.bilayer(4000, 3.11, seed=5), workloads.liposome(3000, 3.45, 2)):
        res = []
        for xs in ("1", "2", "4", "8"):
            os.environ["SMD_XSUB"] = xs
            try:
                ctx = sm.Context.from_dict(m)
            finally:
                del os.environ["SMD_XSUB"]
            ctx.compute_forces(step=2)
            ctx.step(2, 30)
            a = ctx.get_forces()
            nc, key, rank = ctx.get_cell_ids()
            tot, per = ctx.count_pairs(per_particle=True)
            U = ctx.potential()
            dU = ctx.dpotential([1.001, 1.001, 1.0 / 1.001 ** 2])
            res.append((ctx.get_particles()[0], a, key, rank, per, U[sm.TERM_PAIR], dU[sm.TERM_PAIR]))
            ctx.close()
        x0, a0, k0, r0, p0, U0, dU0 = res[0]
        for x, a, k, r, p, U, dU in res[1:]:
            assert np.array_equal(k, k0) and np.array_equal(r, r0) and np.array_equal(p, p0)
            assert np.abs(x - x0).max() <= 1e-9 and rel_force_err(a, a0) <= 1e-9
            assert abs(U - U0) <= 1e-11 * abs(U0) and abs(dU - dU0) <= 1e-11 * abs(U0)


def test_async_snapshot_equals_blocking_readback(orc):
    """smd_snapshot / smd_snapshot_wait (the writer thread's read-back) against smd_get_particles / smd_get_unwrapped,
    with the device stepping on between the snapshot and the wait"""
    m, _ = orc.load_golden(golden_path("bilayer_eq"))
    ctx = sm.Context.from_dict(m, track_unwrapped=True)
    ctx.compute_forces(mask=sm.MASK_ALL, step=0)
    ctx.step(0, 5)
    xyz, _, vel = ctx.get_particles()
    unw = ctx.get_unwrapped()
    w1 = ctx.snapshot(unwrapped=True)
    ctx.step(5, 7)                       # overwrites the state the snapshot was taken from
    w2 = ctx.snapshot()
    a = w1()
    assert np.array_equal(a[0], xyz) and np.array_equal(a[1], vel) and np.array_equal(a[2], unw)
    xyz2, _, vel2 = ctx.get_particles()
    b = w2()
    assert np.array_equal(b[0], xyz2) and np.array_equal(b[1], vel2)
    assert not np.array_equal(xyz, xyz2)
    ctx.close()


@pytest.mark.parametrize("case", ["lipo_eq", "bilayer_eq", "lipocyto_eq", "bead24", "fields", "gas"])
def test_split_pair_engine_changes_nothing_but_the_summation_order(orc, case):
    """k_pair_force2<.., SPLIT = 3> (SMD_PAIR3=1; the default for systems of at most SMD_PAIR3_MAX particles): three threads
    per particle, one per z plane of the stencil, three partial sums added in plane order.  Against the one-thread engine
    (SMD_PAIR3=0): the same pairs -- forces equal to rounding --, the same Metropolis decisions and dU (the force + dPotential
    pass runs through it too), trajectories equal to rounding; a periodic gas exercises the image rows of every plane."""
    import os
    if case == "gas":
        from test_gpu_edge import gas
        m = gas(31, 3000, (14.3, 9.1, 12.7), chains=((150, 4), (100, 3)))
    else:
        m, _ = orc.load_golden(golden_path(case))
    m = dict(m, initialTime=0.0)
    mc = orc.mt_rand53(9, 16)
    runs = []
    for eng in ("0", "1"):
        os.environ["SMD_PAIR3"] = eng
        try:
            ctx = sm.Context.from_dict(m)
        finally:
            del os.environ["SMD_PAIR3"]
        ctx.compute_forces(step=0)
        a_first = ctx.get_forces()
        log = []
        nst = 2 if case == "bead24" else 8    # (the bead fixtures run away after ~9 steps, in the oracle too)
        for t in range(3):
            log.append(ctx.step_mc(nst * t, nst, 0.01, 0.4, mc[2 * t], mc[2 * t + 1]))
        runs.append((a_first, log, ctx.get_particles()[0], ctx.get_forces(), abs(ctx.potential()[sm.TERM_PAIR])))
        ctx.close()
    f0, l0, x0, a0, U0 = runs[0]
    for f1, l1, x1, a1, _ in runs[1:]:
        assert rel_force_err(f1, f0) <= 1e-13
        assert [x[0] for x in l0] == [x[0] for x in l1]
        for (_, d0, b0), (_, d1, b1) in zip(l0, l1):
            assert abs(d0 - d1) <= 1e-12 * U0 and np.allclose(b0, b1, rtol=1e-15)
        assert np.abs(x1 - x0).max() <= 1e-9 and rel_force_err(a1, a0) <= 1e-8
