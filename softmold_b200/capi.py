"""ctypes binding of libsoftmold_b200.so (C ABI: include/softmold_b200.h).

This is plumbing for tests and bench.py; the product is the shared library and the `MD_b200` host driver.  There is
no CPU fallback anywhere: if the library is missing this module raises at import of the symbols, and every compute
call raises SoftMoldError when no B200-class device is usable."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOFTMOLD_B200_LIB") or os.path.join(HERE, "libsoftmold_b200.so")   # override: A/B builds

SMD_OK, SMD_ERR_ARG, SMD_ERR_CUDA, SMD_ERR_CELL, SMD_ERR_IO, SMD_ERR_UNSUPPORTED = range(6)
MOL_BOND, MOL_BEND, MOL_CHAIN, MOL_BEAD, MOL_BALL = 6, 7, 8, 9, 19
MOL_SOLID, MOL_BOUNDARY, MOL_RIGIDBEND, MOL_PULLBEAD, MOL_OFFSET_BOUNDARY = 10, 11, 12, 13, 14
MOL_FLOATING_BASE, MOL_ZTORQUE, MOL_ZPOWERPOTENTIAL, MOL_NANOCORE = 15, 16, 17, 18
MOL_IGNORED = (MOL_SOLID, MOL_RIGIDBEND, MOL_PULLBEAD, MOL_OFFSET_BOUNDARY)   # `MD` parses them and does nothing (MD.cpp:414-478)
TERM_PAIR, TERM_CHAIN, TERM_BOND, TERM_BEND, TERM_BEAD, TERM_BALL, TERM_FIELD, TERM_NANOCORE, NTERMS = 0, 1, 2, 3, 4, 5, 6, 7, 8
MASK_LANGEVIN = 1 << 16
MASK_ALL_MOLECULES = sum(1 << t for t in (TERM_CHAIN, TERM_BOND, TERM_BEND, TERM_BEAD, TERM_BALL, TERM_FIELD, TERM_NANOCORE))
MASK_ALL = (1 << TERM_PAIR) | MASK_ALL_MOLECULES | MASK_LANGEVIN
NOISE_PHILOX, NOISE_EXTERNAL = 0, 1
ABI_VERSION = 2
SLAB_HALO = 2

# every symbol include/softmold_b200.h declares
SYMBOLS = [
    "smd_abi_version", "smd_last_error", "smd_device_count", "smd_create", "smd_destroy", "smd_set_pair_tables",
    "smd_set_particles", "smd_add_chain", "smd_add_bonds", "smd_add_bends", "smd_add_beads", "smd_add_ball",
    "smd_add_boundary", "smd_add_floating_base", "smd_add_ztorque", "smd_add_zpower", "smd_add_nanocore", "smd_set_gamma_type",
    "smd_set_temperature", "smd_set_noise", "smd_build_cells", "smd_compute_forces", "smd_resume", "smd_step",
    "smd_step_begin", "smd_step_end", "smd_potential", "smd_kinetic", "smd_dpotential", "smd_rescale",
    "smd_mc_box_move", "smd_step_mc", "smd_arm_dpotential", "smd_get_particles", "smd_get_forces", "smd_get_unwrapped", "smd_get_box", "smd_get_cell_ids",
    "smd_count_pairs", "smd_synchronize", "smd_device_ptr", "smd_stream", "smd_stats", "smd_profile", "smd_profile_read", "smd_fp64_peak", "smd_mpd_read", "smd_mpd_write",
    "smd_mpd_free", "smd_mpd_get_scalar", "smd_mpd_set_scalar", "smd_mpd_get_size", "smd_mpd_set_size",
    "smd_mpd_particles", "smd_mpd_pair_tables", "smd_mpd_n_molecules", "smd_mpd_molecule", "smd_create_from_mpd",
    "smd_slab_columns", "smd_slab_select", "smd_slab_recv_buffer", "smd_slab_ipc_handle", "smd_slab_connect_ipc",
    "smd_slab_connect_ptr", "smd_slab_exchange_send", "smd_slab_exchange_recv", "smd_slab_counts", "smd_slab_capacity",
    "smd_slab_get_local", "smd_slab_set_local", "smd_mc_propose", "smd_mc_accept",
    "smd_host_alloc", "smd_host_free", "smd_snapshot", "smd_snapshot_wait",
    "smd_add_offset_boundary", "smd_add_rigidbend", "smd_add_pullbead", "smd_create_from_mpd_driver",
    "smd_timeline", "smd_timeline_read",
    "smd_add_inert", "smd_observe", "smd_msd_start", "smd_ke_histogram", "smd_dpotential_device",
]
OBS_BONDS, OBS_EXTENT, OBS_KE_HIST, OBS_MSD = 1, 2, 4, 8


class SoftMoldError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"softmold_b200 error {code}: {msg}")
        self.code = code


class Observables(C.Structure):
    _fields_ = [("lbond_sum", C.c_double), ("n_bond", C.c_int64), ("cos_bend_sum", C.c_double), ("lbend_sum", C.c_double * 2),
                ("n_bend", C.c_int64), ("lo", C.c_double * 3), ("hi", C.c_double * 3), ("n_molecules", C.c_int32)]


class Desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_particles", C.c_int32), ("n_types", C.c_int32), ("device", C.c_int32),
                ("box", C.c_double * 3), ("cutoff", C.c_double), ("dt", C.c_double), ("gamma", C.c_double),
                ("temperature", C.c_double), ("seed", C.c_uint64), ("noise", C.c_int32), ("track_unwrapped", C.c_int32),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("reserved", C.c_int32 * 8)]


_lib = None


def lib():
    """load the shared library; raises (loudly) when it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(softmold_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for s in SYMBOLS:
            getattr(L, s)
        L.smd_last_error.restype = C.c_char_p
        L.smd_last_error.argtypes = [C.c_void_p]
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.smd_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.smd_host_free.argtypes = [vp]
        L.smd_snapshot.argtypes = [vp, vp, vp, vp, C.POINTER(i64)]
        L.smd_snapshot_wait.argtypes = [vp, i64]
        L.smd_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
        L.smd_destroy.argtypes = [vp]
        L.smd_set_pair_tables.argtypes = [vp, vp, vp]
        L.smd_set_particles.argtypes = [vp, vp, vp, vp]
        L.smd_add_chain.argtypes = [vp, i32, vp, vp]
        L.smd_add_bonds.argtypes = [vp, i32, vp, vp]
        L.smd_add_bends.argtypes = [vp, i32, vp, vp]
        L.smd_add_beads.argtypes = [vp, i32, vp, vp]
        L.smd_add_ball.argtypes = [vp, i32, vp, vp]
        for nm in ("boundary", "floating_base", "ztorque", "zpower", "nanocore"):
            getattr(L, "smd_add_" + nm).argtypes = [vp, i32, vp, vp]
        L.smd_set_gamma_type.argtypes = [vp, i32, vp]
        L.smd_set_temperature.argtypes = [vp, dbl]
        L.smd_set_noise.argtypes = [vp, vp]
        L.smd_build_cells.argtypes = [vp]
        L.smd_compute_forces.argtypes = [vp, C.c_uint32, i64]
        L.smd_resume.argtypes = [vp]
        L.smd_step.argtypes = [vp, i64, i32]
        L.smd_step_begin.argtypes = [vp, i64]
        L.smd_step_end.argtypes = [vp, i64]
        L.smd_potential.argtypes = [vp, vp]
        L.smd_kinetic.argtypes = [vp, dp]
        L.smd_dpotential.argtypes = [vp, vp, vp]
        L.smd_rescale.argtypes = [vp, vp, vp]
        L.smd_mc_box_move.argtypes = [vp, dbl, dbl, dbl, dbl, ip, dp, vp]
        L.smd_step_mc.argtypes = [vp, i64, i32, dbl, dbl, dbl, dbl, ip, dp, vp]
        L.smd_arm_dpotential.argtypes = [vp, vp]
        L.smd_get_particles.argtypes = [vp, vp, vp, vp]
        L.smd_get_forces.argtypes = [vp, vp]
        L.smd_get_unwrapped.argtypes = [vp, vp]
        L.smd_get_box.argtypes = [vp, vp]
        L.smd_get_cell_ids.argtypes = [vp, vp, vp, vp]
        L.smd_count_pairs.argtypes = [vp, C.POINTER(i64), vp]
        L.smd_synchronize.argtypes = [vp]
        L.smd_device_ptr.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.smd_stream.argtypes = [vp, C.POINTER(vp)]
        L.smd_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.smd_profile.argtypes = [vp, C.c_uint32]
        L.smd_profile_read.argtypes = [vp, vp, vp]
        L.smd_fp64_peak.argtypes = [vp, dp, dp]
        L.smd_mpd_read.argtypes = [C.c_char_p, C.POINTER(vp), C.c_char_p, C.c_size_t]
        L.smd_mpd_write.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t]
        L.smd_mpd_free.argtypes = [vp]
        L.smd_mpd_free.restype = None
        L.smd_mpd_get_scalar.argtypes = [vp, C.c_char_p, dp, ip]
        L.smd_mpd_set_scalar.argtypes = [vp, C.c_char_p, dbl]
        L.smd_mpd_get_size.argtypes = [vp, vp]
        L.smd_mpd_set_size.argtypes = [vp, vp]
        L.smd_mpd_particles.argtypes = [vp, ip, C.POINTER(dp), C.POINTER(ip), C.POINTER(dp)]
        L.smd_mpd_pair_tables.argtypes = [vp, ip, C.POINTER(dp), C.POINTER(dp)]
        L.smd_mpd_n_molecules.argtypes = [vp]
        L.smd_mpd_molecule.argtypes = [vp, i32, ip, ip, ip, C.POINTER(ip), ip, C.POINTER(dp)]
        L.smd_create_from_mpd.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
        L.smd_slab_columns.argtypes = [i32, i32, i32, ip, ip]
        L.smd_slab_select.argtypes = [vp, dbl, i32, i32, i32, vp, vp]
        L.smd_slab_recv_buffer.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.smd_slab_ipc_handle.argtypes = [vp, i32, vp]
        L.smd_slab_connect_ipc.argtypes = [vp, i32, vp]
        L.smd_slab_connect_ptr.argtypes = [vp, i32, vp]
        L.smd_slab_exchange_send.argtypes = [vp]
        L.smd_slab_exchange_recv.argtypes = [vp]
        L.smd_slab_counts.argtypes = [vp, ip, ip]
        L.smd_slab_capacity.argtypes = [vp, ip]
        L.smd_slab_get_local.argtypes = [vp, ip, vp, vp, vp, vp, vp]
        L.smd_slab_set_local.argtypes = [vp, i32, vp, vp, vp, vp]
        L.smd_mc_propose.argtypes = [vp, dbl, dbl, vp, vp]
        L.smd_mc_accept.argtypes = [dbl, dbl, vp, vp, dbl, dbl, ip, dp]
        L.smd_dpotential_device.argtypes = [vp, vp, C.POINTER(vp)]
        for nm in ("offset_boundary", "rigidbend", "pullbead"):
            getattr(L, "smd_add_" + nm).argtypes = [vp, i32, vp, vp]
        L.smd_create_from_mpd_driver.argtypes = [vp, i32, i32, i32, i32, C.POINTER(vp)]
        L.smd_timeline.argtypes = [vp, i32]
        L.smd_timeline_read.argtypes = [vp, vp]
        L.smd_add_inert.argtypes = [vp, i32]
        L.smd_observe.argtypes = [vp, C.c_uint32, C.POINTER(Observables), vp, vp, i32]
        L.smd_msd_start.argtypes = [vp]
        L.smd_ke_histogram.argtypes = [vp, vp, i64, C.POINTER(i64)]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One simulation resident on one GPU.  Method names follow the C ABI (smd_*), whose entry points in turn replace
    the reference's CellOpt / Verlet / Langevin / Blob::do* seams (see include/softmold_b200.h)."""

    def __init__(self, n_particles, n_types, box, cutoff, dt, gamma, temperature, seed, device=0,
                 noise=NOISE_PHILOX, track_unwrapped=False, _handle=None, rank=0, nranks=1, capacity=0, msg_capacity=0):
        self.L = lib()
        self.n = int(n_particles)
        self.n_types = int(n_types)
        if _handle is not None:
            self.h = _handle
            return
        d = Desc()
        d.abi_version = ABI_VERSION
        d.n_particles, d.n_types, d.device = self.n, self.n_types, device
        d.box = (C.c_double * 3)(*[float(x) for x in box])
        d.cutoff, d.dt, d.gamma, d.temperature = float(cutoff), float(dt), float(gamma), float(temperature)
        d.seed, d.noise, d.track_unwrapped = int(seed), int(noise), int(bool(track_unwrapped))
        d.rank, d.nranks = int(rank), int(nranks)   # nranks > 1: slab rank (n_particles, box are the global ones)
        d.reserved[0], d.reserved[1] = int(capacity), int(msg_capacity)
        self.rank, self.nranks = int(rank), int(nranks)
        self.temperature = float(temperature)
        h = C.c_void_p()
        rc = self.L.smd_create(C.byref(d), C.byref(h))
        if rc:
            raise SoftMoldError(rc, self.L.smd_last_error(None).decode())
        self.h = h

    # -- lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.L.smd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SoftMoldError(rc, self.L.smd_last_error(self.h).decode())

    @classmethod
    def from_dict(cls, m, device=0, noise=NOISE_PHILOX, track_unwrapped=False, driver="md", **slab):
        """m: dict with the .mpd fields (as oracle.orc.read_mpd / load_golden produce them); slab: rank, nranks,
        capacity, msg_capacity for a slab rank of a multi-GPU run (m is always the GLOBAL system)"""
        ctx = cls(m["nParticles"], m["nTypes"], m["size"], m["cutoff"], m["deltaT"], m["gamma"], m["initialTemp"],
                  m["seed"], device=device, noise=noise, track_unwrapped=track_unwrapped, **slab)
        ctx.set_pair_tables(m["twoBodyFconst"], m["twoBodyUconst"])
        ctx.set_particles(m["xyz"], m["type"], m["vel"])
        for mol in m["molecules"]:
            ctx.add_molecule(mol["type"], mol["bonds"], mol["constants"], driver=driver)
        return ctx

    # -- setup
    def set_pair_tables(self, fC, uC):
        fC, uC = _f64(fC), _f64(uC)
        assert fC.size == 6 * self.n_types ** 2 and uC.size == fC.size
        self._ck(self.L.smd_set_pair_tables(self.h, _ptr(fC), _ptr(uC)))

    def set_particles(self, xyz, typ, vel=None):
        xyz, typ = _f64(xyz), _i32(typ)
        assert xyz.shape == (self.n, 3) and typ.shape == (self.n,)
        vel = None if vel is None else _f64(vel)
        self._ck(self.L.smd_set_particles(self.h, _ptr(xyz), _ptr(typ), _ptr(vel)))

    # -- slab decomposition (include/softmold_b200.h, "slab decomposition over several GPUs")
    def slab_recv_buffer(self, side):
        p, b = C.c_void_p(), C.c_size_t()
        self._ck(self.L.smd_slab_recv_buffer(self.h, side, C.byref(p), C.byref(b)))
        return p.value, b.value

    def slab_ipc_handle(self, side):
        buf = C.create_string_buffer(64)
        self._ck(self.L.smd_slab_ipc_handle(self.h, side, buf))
        return buf.raw

    def slab_connect_ipc(self, direction, handle):
        self._ck(self.L.smd_slab_connect_ipc(self.h, direction, C.create_string_buffer(handle, 64)))

    def slab_connect_ptr(self, direction, ptr):
        self._ck(self.L.smd_slab_connect_ptr(self.h, direction, C.c_void_p(ptr)))

    def slab_counts(self, owned=True):
        a, b = C.c_int32(), C.c_int32()
        self._ck(self.L.smd_slab_counts(self.h, C.byref(a), C.byref(b) if owned else None))
        return a.value, b.value

    def slab_capacity(self):
        a = C.c_int32()
        self._ck(self.L.smd_slab_capacity(self.h, C.byref(a)))
        return a.value

    def slab_set_local(self, gid, xyz, typ, vel=None):
        """load this rank's own particles only; every rank calls it, then compute_forces (ghosts arrive by exchange)"""
        gid, xyz, typ = _i32(gid), _f64(xyz), _i32(typ)
        vel = None if vel is None else _f64(vel)
        self._ck(self.L.smd_slab_set_local(self.h, len(gid), _ptr(gid), _ptr(xyz), _ptr(typ), _ptr(vel)))

    def slab_get_local(self, out=None):
        """owned particles of this rank: (gid, xyz, type, vel, acc), arbitrary order.  out: optional preallocated
        (gid[cap], xyz[cap][3], type[cap], vel[cap][3], acc[cap][3] or None) arrays, e.g. page-locked ones"""
        cap = self.slab_capacity()
        n = C.c_int32()
        if out is None:
            gid, typ = np.empty(cap, np.int32), np.empty(cap, np.int32)
            xyz, vel, acc = np.empty((cap, 3)), np.empty((cap, 3)), np.empty((cap, 3))
        else:
            gid, xyz, typ, vel, acc = out
        self._ck(self.L.smd_slab_get_local(self.h, C.byref(n), _ptr(gid), _ptr(xyz), _ptr(typ), _ptr(vel), _ptr(acc)))
        k = n.value
        if acc is None:
            return gid[:k], xyz[:k], typ[:k], vel[:k], None
        return gid[:k], xyz[:k], typ[:k], vel[:k], acc[:k]

    def add_molecule(self, mtype, records, constants, driver="md"):
        """driver: whose molecule switch decides what acts -- "md" (MD.cpp:414-478) or "substrate" (MDsubstrate.cpp:213-262);
        the same rule as smd_create_from_mpd_driver"""
        r, c = _i32(records), _f64(constants)
        n = len(r)
        both = {MOL_CHAIN: self.L.smd_add_chain, MOL_BOND: self.L.smd_add_bonds, MOL_BEND: self.L.smd_add_bends,
                MOL_BEAD: self.L.smd_add_beads, MOL_BOUNDARY: self.L.smd_add_boundary}
        md_only = {MOL_BALL: self.L.smd_add_ball, MOL_FLOATING_BASE: self.L.smd_add_floating_base, MOL_ZTORQUE: self.L.smd_add_ztorque,
                   MOL_ZPOWERPOTENTIAL: self.L.smd_add_zpower, MOL_NANOCORE: self.L.smd_add_nanocore}
        sub_only = {MOL_OFFSET_BOUNDARY: self.L.smd_add_offset_boundary, MOL_RIGIDBEND: self.L.smd_add_rigidbend,
                    MOL_PULLBEAD: self.L.smd_add_pullbead}
        assert driver in ("md", "substrate")
        t = int(mtype)
        f = both.get(t) or (md_only if driver == "md" else sub_only).get(t)
        self.n_molecules = getattr(self, "n_molecules", 0) + 1
        if f is None:
            if t == MOL_SOLID or t in md_only or t in sub_only:     # parsed and ignored by this driver
                self._ck(self.L.smd_add_inert(self.h, t))
                return
            raise SoftMoldError(SMD_ERR_UNSUPPORTED, f"molecule type {mtype} is outside the hot path")
        self._ck(f(self.h, n, _ptr(r), _ptr(c)))

    def set_gamma_type(self, gamma_type):
        g = _f64(gamma_type)
        self._ck(self.L.smd_set_gamma_type(self.h, len(g), _ptr(g)))

    def set_temperature(self, T):
        self.temperature = float(T)
        self._ck(self.L.smd_set_temperature(self.h, float(T)))

    def set_noise(self, u):
        u = _f64(u)
        assert u.shape == (self.n, 3)
        self._ck(self.L.smd_set_noise(self.h, _ptr(u)))

    # -- compute
    def build_cells(self):
        self._ck(self.L.smd_build_cells(self.h))

    def compute_forces(self, mask=MASK_ALL, step=0):
        self._ck(self.L.smd_compute_forces(self.h, mask, step))

    def resume(self):
        self._ck(self.L.smd_resume(self.h))

    def step(self, first_step, nsteps=1):
        self._ck(self.L.smd_step(self.h, first_step, nsteps))

    def step_begin(self, step):
        self._ck(self.L.smd_step_begin(self.h, step))

    def step_end(self, step):
        self._ck(self.L.smd_step_end(self.h, step))

    def potential(self):
        out = np.zeros(NTERMS)
        self._ck(self.L.smd_potential(self.h, _ptr(out)))
        return out

    def kinetic(self):
        v = C.c_double()
        self._ck(self.L.smd_kinetic(self.h, C.byref(v)))
        return v.value

    def dpotential_device(self, scale):
        """smd_dpotential_device: the dPotential terms stay on the device; returns the device address of NTERMS doubles
        (written in stream order on self.stream(), nothing waits)"""
        s, p = _f64(scale), C.c_void_p()
        self._ck(self.L.smd_dpotential_device(self.h, _ptr(s), C.byref(p)))
        return p.value

    TIMELINE_KERNELS = ("k_scan", "k_place", "k_reorder", "k_pair_force2", "k_chain_kick")

    def timeline(self, enable=True):
        self._ck(self.L.smd_timeline(self.h, 1 if enable else 0))

    def timeline_read(self):
        """{kernel: (first block starts working, last block starts, last block ends)} in us after the scan's first block"""
        t = np.zeros(15)
        self._ck(self.L.smd_timeline_read(self.h, _ptr(t)))
        return {k: tuple(float(x) for x in t[3 * i:3 * i + 3]) for i, k in enumerate(self.TIMELINE_KERNELS)}

    def msd_start(self):
        self._ck(self.L.smd_msd_start(self.h))

    def observe(self, what=OBS_BONDS | OBS_EXTENT):
        """dataExtraction::compute's geometric observables reduced on the device (smd_observe): a dict with lbond_sum, n_bond,
        cos_bend_sum, lbend_sum[2], n_bend, lo[3], hi[3] and, with OBS_MSD, msd_sum / msd_count per molecule"""
        o = Observables()
        nm = getattr(self, "n_molecules", 0)
        msd, cnt = np.zeros(max(nm, 1)), np.zeros(max(nm, 1), dtype=np.int64)
        self._ck(self.L.smd_observe(self.h, int(what), C.byref(o), _ptr(msd), _ptr(cnt), nm))
        return {"lbond_sum": o.lbond_sum, "n_bond": o.n_bond, "cos_bend_sum": o.cos_bend_sum, "lbend_sum": np.array(o.lbend_sum[:]),
                "n_bend": o.n_bend, "lo": np.array(o.lo[:]), "hi": np.array(o.hi[:]), "msd_sum": msd[:nm], "msd_count": cnt[:nm]}

    def ke_histogram(self):
        n = C.c_int64()
        self._ck(self.L.smd_ke_histogram(self.h, None, 0, C.byref(n)))
        h = np.zeros(max(n.value, 1), dtype=np.int64)
        self._ck(self.L.smd_ke_histogram(self.h, _ptr(h), n.value, C.byref(n)))
        return h[:n.value]

    def dpotential(self, scale):
        out, s = np.zeros(NTERMS), _f64(scale)
        self._ck(self.L.smd_dpotential(self.h, _ptr(s), _ptr(out)))
        return out

    def rescale(self, scale, new_box):
        s, b = _f64(scale), _f64(new_box)
        self._ck(self.L.smd_rescale(self.h, _ptr(s), _ptr(b)))

    def mc_box_move(self, deltaLXY, tension, u_fluct, u_accept):
        acc, dU, box = C.c_int32(), C.c_double(), np.zeros(3)
        self._ck(self.L.smd_mc_box_move(self.h, deltaLXY, tension, u_fluct, u_accept, C.byref(acc), C.byref(dU), _ptr(box)))
        return bool(acc.value), dU.value, box

    def arm_dpotential(self, scale):
        """the next step() call also sums the pair dPotential of `scale` in its last force pass; dpotential(scale) then reuses it"""
        sc = _f64(scale)
        self._ck(self.L.smd_arm_dpotential(self.h, _ptr(sc)))

    def step_mc(self, first_step, nsteps, deltaLXY, tension, u_fluct, u_accept):
        """step(first_step, nsteps) then mc_box_move(...) in one call: the last step's pair kernel also sums the dPotential"""
        acc, dU, box = C.c_int32(), C.c_double(), np.zeros(3)
        self._ck(self.L.smd_step_mc(self.h, first_step, nsteps, deltaLXY, tension, u_fluct, u_accept, C.byref(acc), C.byref(dU), _ptr(box)))
        return bool(acc.value), dU.value, box

    # -- read back
    def get_particles(self, out=None):
        """out: optional (xyz, type, vel) preallocated C-contiguous arrays (e.g. pinned host memory); an entry may be
        None to skip that read-back"""
        if out is None:
            xyz, typ, vel = np.zeros((self.n, 3)), np.zeros(self.n, np.int32), np.zeros((self.n, 3))
        else:
            xyz, typ, vel = out
        self._ck(self.L.smd_get_particles(self.h, _ptr(xyz), _ptr(typ), _ptr(vel)))
        return xyz, typ, vel

    def snapshot(self, unwrapped=False):
        """asynchronous read-back (smd_snapshot): returns wait() -> (xyz, vel[, unwrapped]) in page-locked buffers"""
        n = self.n
        bufs, ptrs = [], []
        for _ in range(3 if unwrapped else 2):
            p = C.c_void_p()
            self._ck(self.L.smd_host_alloc(C.byref(p), 24 * n))
            ptrs.append(p)
            bufs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n, 3)))
        t = C.c_int64()
        self._ck(self.L.smd_snapshot(self.h, ptrs[0], ptrs[1], ptrs[2] if unwrapped else None, C.byref(t)))

        def wait():
            self._ck(self.L.smd_snapshot_wait(self.h, t))
            out = tuple(b.copy() for b in bufs)
            for p in ptrs:
                self.L.smd_host_free(p)
            return out
        return wait

    def get_forces(self):
        a = np.zeros((self.n, 3))
        self._ck(self.L.smd_get_forces(self.h, _ptr(a)))
        return a

    def get_unwrapped(self):
        a = np.zeros((self.n, 3))
        self._ck(self.L.smd_get_unwrapped(self.h, _ptr(a)))
        return a

    def get_box(self):
        b = np.zeros(3)
        self._ck(self.L.smd_get_box(self.h, _ptr(b)))
        return b

    def get_cell_ids(self):
        nc, key, rank = np.zeros(3, np.int32), np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        self._ck(self.L.smd_get_cell_ids(self.h, _ptr(nc), _ptr(key), _ptr(rank)))
        return nc, key, rank

    def count_pairs(self, per_particle=True):
        tot = C.c_int64()
        per = np.zeros(self.n, np.int32) if per_particle else None
        self._ck(self.L.smd_count_pairs(self.h, C.byref(tot), _ptr(per)))
        return tot.value, per

    def synchronize(self):
        self._ck(self.L.smd_synchronize(self.h))

    def stream(self):
        s = C.c_void_p()
        self._ck(self.L.smd_stream(self.h, C.byref(s)))
        return s.value

    def profile(self, phases=None):
        """enable device timing of the named phases (PHASES), all when None, none when []"""
        mask = 0
        for p in (PHASES if phases is None else phases):
            mask |= 1 << PHASES.index(p)
        self._ck(self.L.smd_profile(self.h, mask))

    def profile_read(self):
        ms, cnt = np.zeros(16), np.zeros(16, np.int64)
        self._ck(self.L.smd_profile_read(self.h, _ptr(ms), _ptr(cnt)))
        return {p: (float(ms[i]), int(cnt[i])) for i, p in enumerate(PHASES)}

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.smd_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stats(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.L.smd_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


PHASES = ["integrate1", "build", "pair", "molecules", "langevin", "integrate2", "step", "exchange", "fused", "pair_du",
          "build_hist", "build_scan", "build_place", "build_reorder"]


class Mpd:
    """`.mpd` file object (host side, no GPU needed): smd_mpd_* of the C ABI."""

    SCALARS = ["gamma", "initialTemp", "finalTemp", "seed", "nTypes", "nMolecules", "nParticles", "periodic", "cutoff",
               "initialTime", "finalTime", "deltaT", "storeInterval", "measureInterval", "deltaLXY", "removeSolvent",
               "tempStepInterval", "tension"]

    def __init__(self, name):
        self.L = lib()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = self.L.smd_mpd_read(os.fsencode(name), C.byref(h), err, len(err))
        if rc:
            raise SoftMoldError(rc, err.value.decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.smd_mpd_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def write(self, name):
        err = C.create_string_buffer(1024)
        rc = self.L.smd_mpd_write(self.h, os.fsencode(name), err, len(err))
        if rc:
            raise SoftMoldError(rc, err.value.decode())

    def scalar(self, cmd):
        v, p = C.c_double(), C.c_int32()
        if self.L.smd_mpd_get_scalar(self.h, cmd.encode(), C.byref(v), C.byref(p)):
            raise KeyError(cmd)
        return v.value, bool(p.value)

    def set_scalar(self, cmd, value):
        if self.L.smd_mpd_set_scalar(self.h, cmd.encode(), float(value)):
            raise KeyError(cmd)

    def size(self):
        s = np.zeros(3)
        self.L.smd_mpd_get_size(self.h, _ptr(s))
        return s

    def to_dict(self):
        """same shape as oracle.orc.read_mpd (copies)"""
        m = {"molecules": []}
        for k in self.SCALARS:
            v, present = self.scalar(k)
            if present:
                m[k] = int(v) if k in ("seed", "nTypes", "nMolecules", "nParticles", "periodic") else v
        m["size"] = list(self.size())
        n = C.c_int32()
        xyz, typ, vel = C.POINTER(C.c_double)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
        self.L.smd_mpd_particles(self.h, C.byref(n), C.byref(xyz), C.byref(typ), C.byref(vel))
        n = n.value
        m["xyz"] = np.ctypeslib.as_array(xyz, (n, 3)).copy() if n else np.zeros((0, 3))
        m["type"] = np.ctypeslib.as_array(typ, (n,)).copy() if n else np.zeros(0, np.int32)
        m["vel"] = np.ctypeslib.as_array(vel, (n, 3)).copy() if n else np.zeros((0, 3))
        nT = C.c_int32()
        fC, uC = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        self.L.smd_mpd_pair_tables(self.h, C.byref(nT), C.byref(fC), C.byref(uC))
        if fC:
            m["twoBodyFconst"] = np.ctypeslib.as_array(fC, (6 * nT.value ** 2,)).copy()
            m["twoBodyUconst"] = np.ctypeslib.as_array(uC, (6 * nT.value ** 2,)).copy()
        for k in range(self.L.smd_mpd_n_molecules(self.h)):
            t, nr, w, nc = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            rec, cst = C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
            self.L.smd_mpd_molecule(self.h, k, C.byref(t), C.byref(nr), C.byref(w), C.byref(rec), C.byref(nc), C.byref(cst))
            r = np.ctypeslib.as_array(rec, (nr.value, w.value)).copy() if nr.value else np.zeros((0, w.value), np.int32)
            c = np.ctypeslib.as_array(cst, (nc.value,)).copy() if nc.value else np.zeros(0)
            m["molecules"].append({"type": t.value, "bonds": r, "constants": c})
        return m

    def create_context(self, device=0, noise=NOISE_PHILOX, track_unwrapped=False):
        h = C.c_void_p()
        rc = self.L.smd_create_from_mpd(self.h, device, noise, int(bool(track_unwrapped)), C.byref(h))
        if rc:
            msg = self.L.smd_last_error(h if h else None).decode()
            if h:
                self.L.smd_destroy(h)
            raise SoftMoldError(rc, msg)
        n, _ = self.scalar("nParticles")
        nT, _ = self.scalar("nTypes")
        return Context(int(n), int(nT), None, None, None, None, None, None, _handle=h)


def slab_columns(n_cols, nranks, rank):
    lo, hi = C.c_int32(), C.c_int32()
    if lib().smd_slab_columns(n_cols, nranks, rank, C.byref(lo), C.byref(hi)):
        raise ValueError("bad slab arguments")
    return lo.value, hi.value


def slab_select(box, cutoff, nranks, rank, xyz):
    """flags per particle: 0 not on this rank, 1 owned, 2 ghost"""
    xyz, box = _f64(xyz), _f64(box)
    flags = np.zeros(len(xyz), np.int32)
    if lib().smd_slab_select(_ptr(box), float(cutoff), nranks, rank, len(xyz), _ptr(xyz), _ptr(flags)):
        raise ValueError("bad slab arguments")
    return flags


def mc_propose(box, deltaLXY, u_fluct):
    box, new_box, scale = _f64(box), np.zeros(3), np.zeros(3)
    lib().smd_mc_propose(_ptr(box), float(deltaLXY), float(u_fluct), _ptr(new_box), _ptr(scale))
    return new_box, scale


def mc_accept(dU_terms_sum, tension, box, new_box, temperature, u_accept):
    acc, dU = C.c_int32(), C.c_double()
    box, new_box = _f64(box), _f64(new_box)
    lib().smd_mc_accept(float(dU_terms_sum), float(tension), _ptr(box), _ptr(new_box), float(temperature), float(u_accept),
                        C.byref(acc), C.byref(dU))
    return bool(acc.value), dU.value
