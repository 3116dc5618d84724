"""Slab decomposition of one MD system over several GPUs: host-side plumbing.

The reference is a single shared-memory process (its MPI attempt, MPImem.h / runMPI.cpp, no longer compiles); large flat
bilayers are split here into slabs of cell columns along x, one C-ABI context (= one GPU) per slab.  Everything on
the data path is in libsoftmold_b200.so: every step each rank's pack kernel writes its migrants and its ghost-column
halo straight into the neighbour's receive buffer through peer memory (NVLink P2P, CUDA IPC between processes) and
the unpack kernel waits on the message header -- no host round trip, no library collective on the per-step path.
This module only
  * wires the buffers once (IPC handles travel through torch.distributed, or plain pointers inside one process),
  * all-reduces the handful of scalars of an energy evaluation / a Metropolis box move (MD.cpp:589-721), whose
    accept decision is then taken identically on every rank from the shared MT19937 draws,
  * gathers particles for read-back.

    DistSlab        one rank per process under torchrun (bench.py --gpus N, NCCL or gloo for the scalars)
    LocalSlabGroup  all ranks inside one process on one device -- the single-GPU test harness of the very same kernels
"""
import numpy as np

from . import capi
from .capi import Context, NTERMS


class _DeviceDoubles:
    """__cuda_array_interface__ of n doubles at a device address the library owns"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def neighbours(rank, nranks):
    """(left, right) ranks of a slab in the periodic ring along x"""
    return (rank - 1) % nranks, (rank + 1) % nranks


class _SlabBase:
    """energy sums, box moves and read-back shared by both drivers; subclasses provide allreduce / the step loop"""

    def potential(self):
        return self._allreduce(self._local_terms(lambda c: c.potential()))

    def dpotential(self, scale):
        return self._allreduce(self._local_terms(lambda c: c.dpotential(scale)))

    def kinetic(self):
        return float(self._allreduce(self._local_terms(lambda c: np.array([c.kinetic()])))[0])

    def count_pairs(self):
        # every rank returns the sum of its owned particles' neighbour counts: each pair is seen from both ends
        return int(self._allreduce(self._local_terms(lambda c: np.array([float(c.count_pairs(per_particle=False)[0])])))[0]) // 2

    def mc_box_move(self, deltaLXY, tension, u_fluct, u_accept):
        """one Metropolis box-move trial, MD.cpp:589-721: every rank proposes the same box, the dPotential terms are
        all-reduced, and the identical decision is applied everywhere"""
        box = self.get_box()
        new_box, scale = capi.mc_propose(box, deltaLXY, u_fluct)
        terms = self.dpotential(scale)
        accepted, dU = capi.mc_accept(self._mc_sum(terms), tension, box, new_box, self.temperature, u_accept)
        if accepted:
            self._each(lambda c: c.rescale(scale, new_box))
        return accepted, dU, (new_box if accepted else box)

    @staticmethod
    def _mc_sum(terms):
        """the terms MD.cpp:642-669 adds into the Metropolis sum: doBallDPotential / doNanoCoreDPotential are evaluated there
        but their results are dropped (the single-GPU trial, smd_mc_box_move, leaves them out the same way)"""
        return float(sum(terms[t] for t in range(NTERMS) if t not in (capi.TERM_BALL, capi.TERM_NANOCORE)))

    def set_temperature(self, T):
        """temperature ramp (MD.cpp:366-371): the thermostat of every context and the acceptance test follow it"""
        self.temperature = float(T)
        self._each(lambda c: c.set_temperature(T))

    def step_mc(self, first_step, nsteps, deltaLXY, tension, u_fluct, u_accept):
        """step(first_step, nsteps) then mc_box_move(...): every rank arms the proposed scaling first, so that the pair kernel
        of its last step also sums its share of the pair dPotential (smd_arm_dpotential) and the trial needs no second pass"""
        box = self.get_box()
        new_box, scale = capi.mc_propose(box, deltaLXY, u_fluct)
        self.step(first_step, nsteps, arm_scale=scale)
        terms = self.dpotential(scale)
        accepted, dU = capi.mc_accept(self._mc_sum(terms), tension, box, new_box, self.temperature, u_accept)
        if accepted:
            self._each(lambda c: c.rescale(scale, new_box))
        return accepted, dU, (new_box if accepted else box)


class LocalSlabGroup(_SlabBase):
    """nranks slab contexts in ONE process on one device, wired to each other with plain device pointers.  The
    kernels, the message protocol and the per-rank state are exactly those of a multi-GPU run; only the transport
    (same-device stores instead of NVLink) differs.  Steps are issued phase by phase (all sends, then all receives)
    so that no rank's wait kernel is enqueued before the kernel it waits for."""

    def __init__(self, m, nranks, device=0, **kw):
        self.nranks = nranks
        self.temperature = float(m["initialTemp"])
        self.ctx = [Context.from_dict(m, device=device, rank=r, nranks=nranks, **kw) for r in range(nranks)]
        for r, c in enumerate(self.ctx):
            left, right = neighbours(r, nranks)
            c.slab_connect_ptr(0, self.ctx[left].slab_recv_buffer(1)[0])
            c.slab_connect_ptr(1, self.ctx[right].slab_recv_buffer(0)[0])

    def close(self):
        for c in self.ctx:
            c.synchronize()
        for c in self.ctx:
            c.close()

    def _each(self, f):
        return [f(c) for c in self.ctx]

    def _local_terms(self, f):
        return self._each(f)

    def _allreduce(self, parts):
        return np.sum(np.stack(parts), axis=0)

    def get_box(self):
        return self.ctx[0].get_box()

    def compute_forces(self, mask=capi.MASK_ALL, step=0):
        self._each(lambda c: c.compute_forces(mask, step))

    batched_default = False

    def step(self, first_step, nsteps=1, batched=None, arm_scale=None):
        """arm_scale: smd_arm_dpotential before the smd_step call that holds the last step (batched mode only).
        batched=False: one step at a time, all sends before all receives.  batched=True: every context enqueues
        a whole smd_step batch (the production call, with the fused step kernel) one after the other; the wait
        kernels of the first contexts then spin until the host has enqueued the later contexts' sends, which is fine
        for the short batches used here (nothing may fill a launch queue while it waits)"""
        if batched is None:
            batched = self.batched_default
        if batched:
            for k in range(0, nsteps, 4):
                if arm_scale is not None and k + 4 >= nsteps:
                    self._each(lambda c: c.arm_dpotential(arm_scale))
                self._each(lambda c: c.step(first_step + k, min(4, nsteps - k)))
            return
        for k in range(nsteps):
            self._each(lambda c: c.step_begin(first_step + k))
            self._each(lambda c: c.synchronize())      # every message is out before any rank starts waiting for one
            self._each(lambda c: c.step_end(first_step + k))

    def synchronize(self):
        self._each(lambda c: c.synchronize())

    def gather(self, n):
        """(xyz, type, vel, acc, owner) of all n particles in original order"""
        xyz, vel, acc = np.full((n, 3), np.nan), np.full((n, 3), np.nan), np.full((n, 3), np.nan)
        typ, owner = np.full(n, -1, np.int32), np.full(n, -1, np.int32)
        for r, c in enumerate(self.ctx):
            g, x, t, v, a = c.slab_get_local()
            assert np.all(owner[g] == -1), "a particle is owned by two ranks"
            xyz[g], typ[g], vel[g], acc[g], owner[g] = x, t, v, a, r
        assert np.all(owner >= 0), "a particle is owned by no rank"
        return xyz, typ, vel, acc, owner


class DistSlab(_SlabBase):
    """one slab rank per process (torchrun): `group` is a torch.distributed process group used for the one-time IPC
    handle exchange and for the scalar all-reduces only"""

    def __init__(self, m, device, rank=None, nranks=None, group=None, **kw):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.nranks = dist.get_world_size(group) if nranks is None else nranks
        self.temperature = float(m["initialTemp"])
        self.device = device
        self.c = Context.from_dict(m, device=device, rank=self.rank, nranks=self.nranks, **kw)
        handles = [None] * self.nranks
        dist.all_gather_object(handles, (self.c.slab_ipc_handle(0), self.c.slab_ipc_handle(1)), group=group)
        left, right = neighbours(self.rank, self.nranks)
        self.c.slab_connect_ipc(0, handles[left][1])
        self.c.slab_connect_ipc(1, handles[right][0])
        dist.barrier(group=group)

    def close(self):
        self.c.synchronize()
        self.dist.barrier(group=self.group)   # nobody unmaps a buffer a neighbour may still write
        self.c.close()

    def _each(self, f):
        return [f(self.c)]

    def _local_terms(self, f):
        return f(self.c)

    def dpotential(self, scale):
        """NCCL: the terms never visit the host before they are reduced -- smd_dpotential_device leaves them in device memory,
        the all-reduce runs in place on that buffer, ordered on the library's stream, and ONE device-to-host copy of the
        total is the trial's only host wait (was: a copy + wait per rank, a staged upload, the all-reduce, a second copy)"""
        if self.dist.get_backend(self.group) != "nccl":
            return super().dpotential(scale)
        import torch
        dev = torch.device("cuda", self.device)
        ptr = self.c.dpotential_device(scale)
        t = torch.as_tensor(_DeviceDoubles(ptr, NTERMS), device=dev)      # zero-copy view of the library's buffer
        with torch.cuda.stream(torch.cuda.ExternalStream(self.c.stream(), device=dev)):
            self.dist.all_reduce(t, group=self.group)
            out = t.cpu()
        return out.numpy()

    def _allreduce(self, part):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64).copy())
        if self.dist.get_backend(self.group) == "nccl":
            t = t.to(torch.device("cuda", self.device))
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def get_box(self):
        return self.c.get_box()

    def compute_forces(self, mask=capi.MASK_ALL, step=0):
        self.c.compute_forces(mask, step)

    def step(self, first_step, nsteps=1, arm_scale=None):
        if arm_scale is not None:
            self.c.arm_dpotential(arm_scale)
        self.c.step(first_step, nsteps)

    def synchronize(self):
        self.c.synchronize()

    def gather(self, n):
        local = self.c.slab_get_local()
        parts = [None] * self.nranks
        self.dist.all_gather_object(parts, local, group=self.group)
        xyz, vel, acc = np.full((n, 3), np.nan), np.full((n, 3), np.nan), np.full((n, 3), np.nan)
        typ, owner = np.full(n, -1, np.int32), np.full(n, -1, np.int32)
        for r, (g, x, t, v, a) in enumerate(parts):
            assert np.all(owner[g] == -1), "a particle is owned by two ranks"
            xyz[g], typ[g], vel[g], acc[g], owner[g] = x, t, v, a, r
        assert np.all(owner >= 0), "a particle is owned by no rank"
        return xyz, typ, vel, acc, owner
