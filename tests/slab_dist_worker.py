"""worker of tests/test_gpu_slab_dist.py and tests/test_slab_cpu.py, started under torch.distributed.run:
    mode "gpu":  one DistSlab rank per process (CUDA IPC transport), checked against a single-GPU context on rank 0
    mode "cpu":  host logic only, gloo: partition completeness and the replicated Metropolis decision"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def mt_rand53(seed, n):
    return np.random.RandomState(seed).random_sample(n)     # MT19937 genrand_res53 = MTRand::rand53


def cpu_mode():
    import torch
    import torch.distributed as dist
    from softmold_b200 import capi, workloads
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    m = workloads.bilayer(3000, 3.11, seed=5)
    n = m["nParticles"]
    flags = capi.slab_select(m["size"], m["cutoff"], world, rank, m["xyz"])
    allf = [None] * world
    dist.all_gather_object(allf, flags)
    allf = np.stack(allf)
    assert np.all((allf == 1).sum(axis=0) == 1), "every particle is owned by exactly one rank"
    # ghosts of rank r = particles of the neighbours within SLAB_HALO columns of r's range (brute force)
    nc = int(m["size"][0] / m["cutoff"])
    cs = m["size"][0] / nc
    col = np.minimum((m["xyz"][:, 0] / cs).astype(np.int64), nc - 1)
    lo, hi = capi.slab_columns(nc, world, rank)
    rel = (col - lo) % nc
    expect = np.where(rel < hi - lo, 1, np.where((rel < hi - lo + capi.SLAB_HALO) | (rel >= nc - capi.SLAB_HALO), 2, 0))
    assert np.array_equal(flags, expect)
    # replicated Metropolis decision from all-reduced partial sums (MD.cpp:589-721)
    box = np.array(m["size"])
    draws = mt_rand53(m["seed"], 40)
    rng = np.random.default_rng(100 + rank)
    decisions = []
    for k in range(20):
        new_box, scale = capi.mc_propose(box, 0.01, draws[2 * k])
        part = torch.tensor([rng.normal() * 0.5])          # this rank's share of dU
        dist.all_reduce(part)
        acc, dU = capi.mc_accept(float(part[0]), 0.5, box, new_box, 3.0, draws[2 * k + 1])
        decisions.append((acc, dU, tuple(new_box)))
        if acc:
            box = new_box
    alld = [None] * world
    dist.all_gather_object(alld, decisions)
    assert all(d == alld[0] for d in alld), "ranks disagree on the accept sequence"
    assert 0 < sum(a for a, _, _ in decisions) < 20
    dist.barrier()
    if rank == 0:
        print("SLAB_CPU_OK")
    dist.destroy_process_group()


def gpu_mode():
    import torch
    import torch.distributed as dist
    import softmold_b200 as sm
    from softmold_b200 import workloads
    from softmold_b200.slab import DistSlab
    ndev = torch.cuda.device_count()
    local = int(os.environ.get("LOCAL_RANK", "0")) % ndev
    torch.cuda.set_device(local)
    backend = "nccl" if ndev >= int(os.environ["WORLD_SIZE"]) else "gloo"   # several ranks on one device: scalars over gloo
    dist.init_process_group(backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    m = workloads.bilayer(5000, 3.11, seed=11)
    n, K, tension = m["nParticles"], 24, 0.5
    mc = mt_rand53(m["seed"], 2 * (K // 8 + 2))

    def run(sim, box_move):
        sim.compute_forces(mask=sm.MASK_ALL, step=0)
        trials = 0
        for i in range(K):
            sim.step(i, 1)
            if i % 8 == 0 and i != 0:
                box_move(m["deltaLXY"], tension, mc[2 * trials], mc[2 * trials + 1])
                trials += 1

    slab = DistSlab(m, local)
    run(slab, slab.mc_box_move)
    U = slab.potential()
    x2, _, v2, _, owner = slab.gather(n)
    box2 = slab.get_box()
    if rank == 0:
        one = sm.Context.from_dict(m, device=local)
        run(one, one.mc_box_move)
        x1, _, v1 = one.get_particles()
        U1 = one.potential()
        assert np.allclose(one.get_box(), box2, rtol=1e-14)
        one.close()
        assert np.abs(x2 - x1).max() <= 1e-9, np.abs(x2 - x1).max()
        assert np.abs(v2 - v1).max() <= 1e-8
        assert abs(U.sum() - U1.sum()) <= 1e-10 * abs(U1.sum())
        assert len(set(owner)) == world
        print("SLAB_DIST_OK ranks=%d devices=%d" % (world, ndev))
    slab.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    cpu_mode() if sys.argv[1] == "cpu" else gpu_mode()
