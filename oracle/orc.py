"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (the plain-C restatement of the
reference `MD` timestep, oracle/oracle.c) plus readers for the reference harness' dump container and for
`.mpd` files.  Nothing under softmold_b200/ imports this module."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF = os.path.join(HERE, "_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    """compile oracle.c -> liboracle.so (gcc only; always possible)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "oracle.c")):
            build()
        L = C.CDLL(LIB)
        d3 = C.c_double * 3
        L.orc_cell_ids.argtypes = [C.c_int, _dp, d3, C.c_double, _ip]
        L.orc_build.argtypes = [C.c_int, _dp, d3, C.c_double, _ip, _ip, _ip]
        L.orc_build.restype = C.c_int
        L.orc_pair_force.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, d3, C.c_double, _dp]
        L.orc_pair_force.restype = C.c_int
        L.orc_pair_potential.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, d3, C.c_double]
        L.orc_pair_potential.restype = C.c_double
        L.orc_pair_dpotential.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, d3, C.c_double, d3]
        L.orc_pair_dpotential.restype = C.c_double
        L.orc_pair_count.argtypes = [C.c_int, _dp, _ip, C.c_int, d3, C.c_double, C.c_void_p]
        L.orc_pair_count.restype = C.c_longlong
        d4 = C.c_double * 4
        d2 = C.c_double * 2
        L.orc_chain_force.argtypes = [_dp, d3, C.c_int, C.c_int, C.c_int, d4, _dp]
        L.orc_chain_potential.argtypes = [_dp, d3, C.c_int, C.c_int, C.c_int, d4]
        L.orc_chain_potential.restype = C.c_double
        L.orc_chain_dpotential.argtypes = [_dp, d3, C.c_int, C.c_int, C.c_int, d4, d3]
        L.orc_chain_dpotential.restype = C.c_double
        for nm in ("bond", "bend", "ball"):
            getattr(L, f"orc_{nm}_force").argtypes = [_dp, d3, C.c_int, _ip, d2, _dp]
            getattr(L, f"orc_{nm}_potential").argtypes = [_dp, d3, C.c_int, _ip, d2]
            getattr(L, f"orc_{nm}_potential").restype = C.c_double
            getattr(L, f"orc_{nm}_dpotential").argtypes = [_dp, d3, C.c_int, _ip, d2, d3]
            getattr(L, f"orc_{nm}_dpotential").restype = C.c_double
        L.orc_boundary_force.argtypes = [_dp, d3, C.c_int, _ip, d4, _dp]
        L.orc_offset_boundary_force.argtypes = [_dp, d3, C.c_int, _ip, d4, _dp]
        L.orc_rigidbend_force.argtypes = [_dp, d3, C.c_int, _ip, C.c_double * 5, _dp]
        L.orc_pullbead_force.argtypes = [C.c_int, _dp, d3, C.c_int, _ip, d4, _dp]
        L.orc_boundary_potential.argtypes = [_dp, d3, C.c_int, _ip, d4]
        L.orc_boundary_potential.restype = C.c_double
        L.orc_floating_base_force.argtypes = [_dp, _ip, C.c_int, _ip, _dp, _dp]
        L.orc_floating_base_potential.argtypes = [_dp, _ip, C.c_int, _ip, _dp]
        L.orc_floating_base_potential.restype = C.c_double
        L.orc_ztorque_force.argtypes = [_dp, d3, C.c_int, _ip, d4, _dp]
        L.orc_ztorque_potential.argtypes = [_dp, d3, C.c_int, _ip, d4]
        L.orc_ztorque_potential.restype = C.c_double
        L.orc_zpower_force.argtypes = [_dp, C.c_int, _ip, d2, _dp]
        L.orc_zpower_potential.argtypes = [_dp, C.c_int, _ip, d2]
        L.orc_zpower_potential.restype = C.c_double
        L.orc_nanocore_force.argtypes = [C.c_int, _dp, d3, C.c_int, _ip, _dp, _dp]
        L.orc_nanocore_potential.argtypes = [C.c_int, _dp, d3, C.c_int, _ip, _dp]
        L.orc_nanocore_potential.restype = C.c_double
        L.orc_nanocore_dpotential.argtypes = [C.c_int, _dp, d3, C.c_int, _ip, _dp, d3]
        L.orc_nanocore_dpotential.restype = C.c_double
        L.orc_bead_force.argtypes = [C.c_int, _dp, _ip, C.c_int, d3, C.c_int, C.c_int, _ip, _dp, _dp]
        L.orc_bead_potential.argtypes = [C.c_int, _dp, _ip, C.c_int, d3, C.c_int, C.c_int, _ip, _dp]
        L.orc_bead_potential.restype = C.c_double
        L.orc_bead_dpotential.argtypes = [C.c_int, _dp, _ip, C.c_int, d3, C.c_int, C.c_int, _ip, _dp, d3]
        L.orc_bead_dpotential.restype = C.c_double
        L.orc_verlet_first.argtypes = [C.c_int, _dp, _dp, _dp, _ip, d3, C.c_double, C.c_void_p]
        L.orc_verlet_second.argtypes = [C.c_int, _dp, _dp, _ip, C.c_double]
        L.orc_langevin.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_kinetic.argtypes = [C.c_int, _dp]
        L.orc_kinetic.restype = C.c_double
        L.orc_mt_sizeof.restype = C.c_size_t
        L.orc_mt_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_mt_u32.argtypes = [C.c_void_p]
        L.orc_mt_u32.restype = C.c_uint32
        L.orc_mt_rand53.argtypes = [C.c_void_p]
        L.orc_mt_rand53.restype = C.c_double
        L.orc_mt_fill53.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_philox_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, _dp]
        L.orc_md_init.argtypes = [C.c_void_p, C.c_longlong]
        L.orc_md_step.argtypes = [C.c_void_p, C.c_longlong]
        L.orc_md_resume.argtypes = [C.c_void_p]
        L.orc_md_step_begin.argtypes = [C.c_void_p, C.c_longlong]
        L.orc_md_step_end.argtypes = [C.c_void_p, C.c_longlong]
        L.orc_total_potential.argtypes = [C.c_void_p]
        L.orc_total_potential.restype = C.c_double
        L.orc_total_dpotential.argtypes = [C.c_void_p, d3]
        L.orc_total_dpotential.restype = C.c_double
        _lib = L
    return _lib


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


# ------------------------------------------------------------------ thin functional wrappers
def cell_ids(xyz, box, cutoff):
    n = len(xyz)
    key = np.zeros(n, np.int32)
    lib().orc_cell_ids(n, np.ascontiguousarray(xyz, np.float64), _d3(box), cutoff, key)
    return key


def build_lists(xyz, box, cutoff):
    n = len(xyz)
    nc = [int(box[d] / cutoff) for d in range(3)]
    nct = nc[0] * nc[1] * nc[2]
    head = np.zeros(nct, np.int32)
    nxt = np.zeros(max(n, 1), np.int32)
    full = np.zeros(nct, np.int32)
    nf = lib().orc_build(n, np.ascontiguousarray(xyz, np.float64), _d3(box), cutoff, head, nxt, full)
    return head, nxt[:n], full[:nf]


def pair_force(xyz, typ, nT, fC, box, cutoff):
    acc = np.zeros((len(xyz), 3))
    rc = lib().orc_pair_force(len(xyz), np.ascontiguousarray(xyz, np.float64), np.ascontiguousarray(typ, np.int32), nT,
                              np.ascontiguousarray(fC, np.float64), _d3(box), cutoff, acc)
    if rc:
        raise RuntimeError("oracle: too many particles in a cell")
    return acc


def pair_potential(xyz, typ, nT, uC, box, cutoff):
    return lib().orc_pair_potential(len(xyz), np.ascontiguousarray(xyz, np.float64), np.ascontiguousarray(typ, np.int32),
                                    nT, np.ascontiguousarray(uC, np.float64), _d3(box), cutoff)


def pair_dpotential(xyz, typ, nT, uC, box, cutoff, scale):
    return lib().orc_pair_dpotential(len(xyz), np.ascontiguousarray(xyz, np.float64), np.ascontiguousarray(typ, np.int32),
                                     nT, np.ascontiguousarray(uC, np.float64), _d3(box), cutoff, _d3(scale))


def pair_count(xyz, typ, nT, box, cutoff, per_particle=False):
    n = len(xyz)
    nc = np.zeros(n, np.int32) if per_particle else None
    tot = lib().orc_pair_count(n, np.ascontiguousarray(xyz, np.float64), np.ascontiguousarray(typ, np.int32), nT,
                               _d3(box), cutoff, nc.ctypes.data if per_particle else None)
    return (tot, nc) if per_particle else tot


def mt_rand53(seed, count):
    st = C.create_string_buffer(lib().orc_mt_sizeof())
    lib().orc_mt_seed(st, seed)
    out = np.zeros(count)
    lib().orc_mt_fill53(st, count, out)
    return out


def mt_u32(seed, count):
    st = C.create_string_buffer(lib().orc_mt_sizeof())
    lib().orc_mt_seed(st, seed)
    return np.array([lib().orc_mt_u32(st) for _ in range(count)], np.uint32)


def philox_uniforms(seed, step, n, ids=None):
    u = np.zeros(3 * n)
    if ids is not None:
        ids = np.ascontiguousarray(ids, np.int32)
    lib().orc_philox_uniforms(seed, step, n, None if ids is None else ids.ctypes.data, u)
    return u.reshape(n, 3)


def kinetic(vel):
    return lib().orc_kinetic(len(vel), np.ascontiguousarray(vel, np.float64))


# ------------------------------------------------------------------ whole system (mirrors orc_sys / orc_mol)
class _Mol(C.Structure):
    _fields_ = [("type", C.c_int), ("nbond", C.c_int), ("bonds", C.c_void_p), ("c", C.c_void_p)]


class _MT(C.Structure):
    _fields_ = [("s", C.c_uint32 * 624), ("pos", C.c_int)]


class _Sys(C.Structure):
    _fields_ = [("n", C.c_int), ("nT", C.c_int), ("xyz", C.c_void_p), ("vel", C.c_void_p), ("acc", C.c_void_p),
                ("type", C.c_void_p), ("box", C.c_double * 3), ("cutoff", C.c_double), ("dt", C.c_double),
                ("gamma", C.c_double), ("temperature", C.c_double), ("deltaLXY", C.c_double), ("tension", C.c_double),
                ("fC", C.c_void_p), ("uC", C.c_void_p), ("nmol", C.c_int), ("mol", C.c_void_p), ("seed", C.c_uint32),
                ("noise", C.c_int), ("lang_rng", _MT), ("mc_rng", _MT), ("trials", C.c_longlong),
                ("accepted", C.c_longlong), ("unwrapped", C.c_void_p), ("last_dU", C.c_double)]


BOND, BEND, CHAIN, BEAD, BALL = 6, 7, 8, 9, 19
SOLID, BOUNDARY, RIGIDBEND, PULLBEAD, OFFSET_BOUNDARY, FLOATING_BASE, ZTORQUE, ZPOWER, NANOCORE = 10, 11, 12, 13, 14, 15, 16, 17, 18


class System:
    """The oracle's MD loop over a parsed .mpd (see read_mpd).  noise: 'mt' = the reference's single MT19937
    stream (OMP_NUM_THREADS=1), 'philox' = the product's counter-based noise."""

    def __init__(self, mpd, noise="mt"):
        self.m = mpd
        self.n = mpd["nParticles"]
        self.xyz = np.ascontiguousarray(mpd["xyz"], np.float64).copy()
        self.vel = np.ascontiguousarray(mpd["vel"], np.float64).copy()
        self.acc = np.zeros_like(self.xyz)
        self.type = np.ascontiguousarray(mpd["type"], np.int32).copy()
        self.fC = np.ascontiguousarray(mpd["twoBodyFconst"], np.float64)
        self.uC = np.ascontiguousarray(mpd["twoBodyUconst"], np.float64)
        self._keep = []
        mols = (_Mol * max(1, len(mpd["molecules"])))()
        for k, mol in enumerate(mpd["molecules"]):
            b = np.ascontiguousarray(mol["bonds"], np.int32)
            c = np.ascontiguousarray(mol["constants"], np.float64)
            self._keep += [b, c]
            mols[k].type = mol["type"]
            mols[k].nbond = len(b)
            mols[k].bonds = b.ctypes.data
            mols[k].c = c.ctypes.data
        self._mols = mols
        s = _Sys()
        assert C.sizeof(_Sys) == lib().orc_sys_sizeof() and C.sizeof(_Mol) == lib().orc_mol_sizeof()
        s.n, s.nT = self.n, mpd["nTypes"]
        s.xyz, s.vel, s.acc, s.type = self.xyz.ctypes.data, self.vel.ctypes.data, self.acc.ctypes.data, self.type.ctypes.data
        s.box = _d3(mpd["size"])
        s.cutoff, s.dt, s.gamma = mpd["cutoff"], mpd["deltaT"], mpd["gamma"]
        s.temperature = mpd["initialTemp"]
        s.deltaLXY, s.tension = mpd.get("deltaLXY", 0.0), mpd.get("tension", 0.0)
        s.fC, s.uC = self.fC.ctypes.data, self.uC.ctypes.data
        s.nmol, s.mol = len(mpd["molecules"]), C.addressof(mols)
        s.seed = mpd["seed"]
        s.noise = {"mt": 0, "philox": 1}[noise]
        s.unwrapped = None
        self.s = s

    @property
    def box(self):
        return np.array(list(self.s.box))

    def init(self, step=0):
        lib().orc_md_init(C.byref(self.s), step)

    def step(self, i):
        lib().orc_md_step(C.byref(self.s), i)

    def resume(self):
        lib().orc_md_resume(C.byref(self.s))

    def run_like_reference(self, nsteps):
        """Replays what `MD name` does for nsteps iterations and returns the state it would have stored at its
        last iteration (after Verlet::first, before the new forces; MD.cpp:373-381)."""
        start = int(self.m["initialTime"] / self.m["deltaT"] + 1e-7)
        self.init(start)
        if self.m["initialTime"] != 0:
            self.resume()
        for i in range(start, start + nsteps):
            self.step(i)
        self.step_begin(start + nsteps)
        return self.xyz, self.vel, self.box

    def step_begin(self, i):
        lib().orc_md_step_begin(C.byref(self.s), i)

    def step_end(self, i):
        lib().orc_md_step_end(C.byref(self.s), i)

    def potential(self):
        return lib().orc_total_potential(C.byref(self.s))

    def dpotential(self, scale):
        return lib().orc_total_dpotential(C.byref(self.s), _d3(scale))


# ------------------------------------------------------------------ file readers (test side)
def read_dump(path):
    """the SMDG1 container written by oracle/ref_harness.cpp"""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8)[:5] == b"SMDG1"
        while True:
            h = f.read(41)
            if len(h) < 41:
                break
            name = h[:32].split(b"\0")[0].decode()
            dtype = chr(h[32])
            (count,) = struct.unpack("<q", h[33:41])
            if dtype == "d":
                out[name] = np.frombuffer(f.read(8 * count), np.float64).copy()
            else:
                out[name] = np.frombuffer(f.read(4 * count), np.int32).copy()
    return out


# type -> (nConstants, ints per bond record), system.h:1036-1120; BEAD / FLOATING_BASE / NANOCORE depend on nTypes / nBonds
_MOL_SHAPE = {BOND: (2, 2), BEND: (2, 3), CHAIN: (4, 3), BALL: (2, 2), SOLID: (1, 1), BOUNDARY: (4, 1), OFFSET_BOUNDARY: (4, 1),
              RIGIDBEND: (5, 2), PULLBEAD: (4, 1), ZTORQUE: (4, 3), ZPOWER: (2, 2)}


def mol_shape(t, nTypes, nBonds):
    if t == BEAD:
        return 22 * nTypes ** 2, 1
    if t == FLOATING_BASE:
        return 6 * nTypes, 1
    if t == NANOCORE:
        return 22 * nBonds, 1
    return _MOL_SHAPE[t]


def read_mpd(path):
    """minimal whitespace-token .mpd reader (format contract: system.h:589-1313 of the reference)."""
    tok = open(path).read().split()
    i = 0
    m = {"molecules": []}
    scal = {"gamma": float, "initialTemp": float, "finalTemp": float, "seed": int, "nTypes": int, "nMolecules": int,
            "nParticles": int, "periodic": int, "cutoff": float, "initialTime": float, "finalTime": float,
            "deltaT": float, "storeInterval": float, "measureInterval": float, "deltaLXY": float,
            "removeSolvent": float, "tempStepInterval": float, "tension": float}
    while i < len(tok):
        w = tok[i]
        i += 1
        if w in scal:
            m[w] = scal[w](tok[i])
            i += 1
        elif w == "size":
            m[w] = [float(t) for t in tok[i:i + 3]]
            i += 3
        elif w in ("twoBodyFconst", "twoBodyUconst"):
            k = 6 * m["nTypes"] ** 2
            m[w] = np.array(tok[i:i + k], np.float64)
            i += k
        elif w == "positions":
            k = 4 * m["nParticles"]
            a = np.array(tok[i:i + k], np.float64).reshape(-1, 4)
            m["type"] = a[:, 0].astype(np.int32)
            m["xyz"] = np.ascontiguousarray(a[:, 1:])
            i += k
        elif w == "velocities":
            k = 3 * m["nParticles"]
            m["vel"] = np.array(tok[i:i + k], np.float64).reshape(-1, 3)
            i += k
        elif w == "molecule":
            for _ in range(m["nMolecules"]):
                t, nb = int(tok[i]), int(tok[i + 1])
                i += 2
                ncst, width = mol_shape(t, m["nTypes"], nb)
                c = np.array(tok[i:i + ncst], np.float64)
                i += ncst
                b = np.array(tok[i:i + nb * width], np.int32).reshape(nb, width)
                i += nb * width
                m["molecules"].append({"type": t, "constants": c, "bonds": b})
        elif w == "gammaType":
            m[w] = [float(t) for t in tok[i:i + m["nTypes"]]]
            i += m["nTypes"]
        elif w == "banana":
            pass
        else:
            raise ValueError(f"unrecognised .mpd command {w!r}")
    return m


def write_mpd(path, m):
    """write a dict as produced by read_mpd back to .mpd text (17 significant digits: loss-free)."""
    r = lambda x: repr(float(x))
    with open(path, "w") as f:
        for k in ("gamma", "initialTemp", "finalTemp", "seed", "nTypes", "nMolecules", "nParticles", "periodic", "cutoff"):
            f.write(f"{k} {m[k]}\n")
        f.write("size " + " ".join(r(x) for x in m["size"]) + "\n")
        for k in ("initialTime", "finalTime", "deltaT", "storeInterval", "measureInterval"):
            f.write(f"{k} {r(m[k])}\n")
        for k in ("twoBodyFconst", "twoBodyUconst"):
            f.write(k + "\n")
            for row in np.asarray(m[k]).reshape(-1, 6):
                f.write(" " + " ".join(r(x) for x in row) + "\n")
        f.write("positions\n")
        for t, p in zip(m["type"], m["xyz"]):
            f.write(f" {int(t)} {r(p[0])} {r(p[1])} {r(p[2])}\n")
        f.write("velocities\n")
        for v in m["vel"]:
            f.write(f" {r(v[0])} {r(v[1])} {r(v[2])}\n")
        if m["molecules"]:
            f.write("molecule\n")
            for mol in m["molecules"]:
                f.write(f"{mol['type']}\t{len(mol['bonds'])}\n")
                f.write(" " + " ".join(r(x) for x in mol["constants"]) + "\n")
                for b in mol["bonds"]:
                    f.write(" " + " ".join(str(int(x)) for x in b) + "\n")
        for k in ("deltaLXY", "removeSolvent", "tempStepInterval", "tension"):
            if k in m:
                f.write(f"{k} {r(m[k])}\n")
        if "gammaType" in m:
            f.write("gammaType " + " ".join(r(x) for x in m["gammaType"]) + "\n")


def load_golden(path):
    """tests/golden/*.npz -> (mpd-like dict usable by System / the product API, dict of reference outputs)."""
    z = np.load(path)
    m = {"nTypes": int(z["nTypes"]), "size": [float(x) for x in z["box"]], "cutoff": float(z["cutoff"]),
         "deltaT": float(z["deltaT"]), "gamma": float(z["gamma"]), "initialTemp": float(z["temperature"]),
         "finalTemp": float(z["temperature"]), "seed": int(z["seed"]), "initialTime": float(z["initialTime"]),
         "nParticles": len(z["xyz"]), "xyz": z["xyz"], "vel": z["vel"], "type": z["type"].astype(np.int32),
         "twoBodyFconst": z["fC"], "twoBodyUconst": z["uC"], "molecules": [], "periodic": 1,
         "finalTime": 0.0, "storeInterval": 1e9, "measureInterval": 1e9}
    if float(z["deltaLXY"]) != 0:
        m["deltaLXY"] = float(z["deltaLXY"])
    if float(z["tension"]) != 0:
        m["tension"] = float(z["tension"])
    for k in range(int(z["nmol"])):
        m["molecules"].append({"type": int(z[f"mol{k}_type"]), "bonds": z[f"mol{k}_bonds"].astype(np.int32),
                               "constants": z[f"mol{k}_const"]})
    m["nMolecules"] = len(m["molecules"])
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    ref.update({k: z[k] for k in z.files if k.startswith("traj_")})
    return m, ref
