"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes binding of oracle/libamr_oracle.so (the C restatement of the reference hot path) plus the
script runner shared by the golden generator and the parity tests.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EQ_ADVECTION, EQ_EULER = 0, 1
STABLE, REFINE, COARSEN = 0, 1, 2


def build():
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libamr_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_tree_create.restype = C.c_void_p
        L.orc_tree_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                      C.c_size_t]
        L.orc_tree_destroy.argtypes = [C.c_void_p]
        L.orc_tree_size.restype = C.c_size_t
        L.orc_tree_size.argtypes = [C.c_void_p]
        L.orc_tree_flat_size.restype = C.c_size_t
        L.orc_tree_flat_size.argtypes = [C.c_void_p]
        L.orc_tree_ids.restype = C.POINTER(C.c_uint64)
        L.orc_tree_ids.argtypes = [C.c_void_p]
        L.orc_tree_field.restype = C.POINTER(C.c_double)
        L.orc_tree_field.argtypes = [C.c_void_p, C.c_int]
        L.orc_tree_reconstruct.restype = C.c_int
        L.orc_tree_reconstruct.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tree_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_halo_exchange.argtypes = [C.c_void_p]
        L.orc_compute_dt.restype = C.c_double
        L.orc_compute_dt.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double,
                                     C.POINTER(C.c_double)]
        L.orc_time_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double,
                                    C.POINTER(C.c_double)]
        L.orc_advance_batch.restype = C.c_double
        L.orc_advance_batch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double,
                                        C.POINTER(C.c_double), C.c_size_t, C.c_double,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


class Config:
    """Compile-time configuration of a reference instantiation, as runtime values."""

    def __init__(self, rank, size, halo, depth, eq, length=None, gamma=1.4, cfl=0.3):
        self.rank, self.size, self.halo, self.depth, self.eq = rank, size, halo, depth, eq
        self.length = (1000.0 if eq == EQ_EULER else 1.0) if length is None else length
        self.gamma, self.cfl = gamma, cfl
        self.nvar = 1 if eq == EQ_ADVECTION else rank + 2
        self.psize = size + 2 * halo
        self.flat = self.psize ** rank
        self.ndir = 2 * rank
        self.kf = 1 << (rank - 1)

    @property
    def name(self):
        return "r%d_s%d_h%d_d%d_%s" % (self.rank, self.size, self.halo, self.depth,
                                        "euler" if self.eq == EQ_EULER else "adv")

    @staticmethod
    def from_name(name, **kw):
        r, s, h, d, e = name.split("_")
        return Config(int(r[1:]), int(s[1:]), int(h[1:]), int(d[1:]),
                      EQ_EULER if e == "euler" else EQ_ADVECTION, **kw)


# ---------------------------------------------------------------------------- id helpers
def morton_decode(ids, rank):
    """ids (uint64) -> (coords[n, rank] uint32 in finest-level units (x first), level[n])."""
    ids = np.asarray(ids, dtype=np.uint64)
    level = (ids & np.uint64(63)).astype(np.int64)
    m = ids >> np.uint64(6)
    coords = np.zeros((len(ids), rank), dtype=np.uint64)
    for b in range(20):
        for a in range(rank):
            coords[:, a] |= ((m >> np.uint64(rank * b + a)) & np.uint64(1)) << np.uint64(b)
    return coords.astype(np.int64), level


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def flags_all(ids):
    return np.full(len(ids), REFINE, dtype=np.int8)


def flags_hash(ids, seed, pr, pc, minl, maxl):
    ids = np.asarray(ids, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = ids ^ (np.uint64(seed) * np.uint64(0x100000001B3))
    u = (splitmix64(key) % np.uint64(1000)).astype(np.int64)
    lvl = (ids & np.uint64(63)).astype(np.int64)
    f = np.zeros(len(ids), dtype=np.int8)
    f[(u < pr) & (lvl < maxl)] = REFINE
    f[(u >= pr) & (u < pr + pc) & (lvl > minl) & (f == 0)] = COARSEN
    return f


def flags_ball(ids, cfg, r, minl, maxl, centre):
    coords, lvl = morton_decode(ids, cfg.rank)
    L = cfg.length
    r2 = np.zeros(len(ids))
    for d in range(cfg.rank):
        org = L * coords[:, d].astype(np.float64) / float(1 << cfg.depth)
        sz = L * (2.0 ** (cfg.depth - lvl)) / float(1 << cfg.depth)
        x = (org + 0.5 * sz) / L - centre[d]
        r2 = r2 + x * x
    f = np.zeros(len(ids), dtype=np.int8)
    f[(r2 < r * r) & (lvl < maxl)] = REFINE
    f[(r2 > 4 * r * r) & (lvl > minl) & (f == 0)] = COARSEN
    return f


def cell_centres(ids, cfg):
    """Physical cell-centre coordinates of every interior cell:
    [P, S.., rank] following amr_solver::initialize (amr_solver.hpp:105-145, physics_system.hpp:107-138)."""
    coords, lvl = morton_decode(ids, cfg.rank)
    R, S, L = cfg.rank, cfg.size, cfg.length
    out = np.zeros((len(ids),) + (S,) * R + (R,))
    for d in range(R):  # physical dim d <-> layout dim R-1-d
        org = L * coords[:, d].astype(np.float64) / float(1 << cfg.depth)
        dx = (L * (2.0 ** (cfg.depth - lvl)) / float(1 << cfg.depth)) / float(S)
        shape = [1] * (R + 1)
        shape[0] = len(ids)
        k = np.arange(S, dtype=np.float64)
        kshape = [1] * (R + 1)
        kshape[1 + (R - 1 - d)] = S
        cell_org = org.reshape(shape) + k.reshape(kshape) * dx.reshape(shape)
        out[..., d] = cell_org + 0.5 * dx.reshape(shape)
    return out


def initial_condition(ids, cfg):
    """Conservative interior state [nvar, P, S..] of the benchmark ICs (see oracle/ref_dump.cpp ic())."""
    x = cell_centres(ids, cfg)
    L, R = cfg.length, cfg.rank
    if cfg.eq == EQ_ADVECTION:
        r2 = sum((x[..., d] - 0.2 * L) * (x[..., d] - 0.2 * L) for d in range(R))
        return np.exp(-r2 / (0.005 * L * L))[None]
    r2 = sum((x[..., d] - 0.5 * L) * (x[..., d] - 0.5 * L) for d in range(R))
    pert = 10.0 * np.exp(-r2 / (0.01 * L * L))
    rho = 0.5 + pert * 0.2
    p = 1.0 + pert
    z = np.zeros_like(rho)
    E = p / (cfg.gamma - 1.0) + 0.5 * rho * 0.0
    return np.stack([rho] + [z] * R + [E])


def interior_slices(cfg):
    h = cfg.halo
    return (slice(h, h + cfg.size),) * cfg.rank


def face_halo_mask(cfg):
    """Boolean mask over the padded patch: interior + face-halo cells (everything the
    reference ever writes; corner/edge ghosts are excluded)."""
    h, S, R = cfg.halo, cfg.size, cfg.rank
    idx = np.indices((cfg.psize,) * R)
    outside = [(idx[k] < h) | (idx[k] >= h + S) for k in range(R)]
    return sum(o.astype(int) for o in outside) <= 1


# ---------------------------------------------------------------------------- oracle tree
class OracleTree:
    def __init__(self, cfg, capacity=20000):
        self.cfg = cfg
        self.L = lib()
        size = (C.c_int * 3)(*([cfg.size] * cfg.rank + [1] * (3 - cfg.rank)))
        self.h = self.L.orc_tree_create(cfg.rank, cfg.depth, size, cfg.halo, cfg.nvar, capacity)
        if not self.h:
            raise RuntimeError("orc_tree_create failed")
        self._len = (C.c_double * 3)(*([cfg.length] * 3))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_tree_destroy(self.h)
            self.h = None

    @property
    def size(self):
        return self.L.orc_tree_size(self.h)

    def ids(self):
        n = self.size
        return np.ctypeslib.as_array(self.L.orc_tree_ids(self.h), shape=(n,)).copy()

    def reconstruct(self, flags):
        flags = np.ascontiguousarray(flags, dtype=np.int8)
        assert len(flags) == self.size
        rc = self.L.orc_tree_reconstruct(self.h, flags.ctypes.data)
        if rc < 0:
            raise RuntimeError("oracle tree capacity exceeded")
        return rc

    def tables(self):
        cfg, n = self.cfg, self.size
        rel = np.zeros((n, cfg.ndir), dtype=np.int8)
        nbr = np.zeros((n, cfg.ndir, cfg.kf), dtype=np.int32)
        quad = np.zeros((n, cfg.ndir, cfg.rank), dtype=np.int8)
        self.L.orc_tree_tables(self.h, rel.ctypes.data, nbr.ctypes.data, quad.ctypes.data)
        return rel, nbr, quad

    def field_view(self, f):
        """Writable view of field f's current padded buffer: [P, psize..]."""
        cfg, n = self.cfg, self.size
        a = np.ctypeslib.as_array(self.L.orc_tree_field(self.h, f), shape=(n * cfg.flat,))
        return a.reshape((n,) + (cfg.psize,) * cfg.rank)

    def get_padded(self):
        return np.stack([self.field_view(f).copy() for f in range(self.cfg.nvar)])

    def set_padded(self, data):
        for f in range(self.cfg.nvar):
            self.field_view(f)[...] = data[f].reshape(self.field_view(f).shape)

    def set_interior(self, data):
        sl = (slice(None),) + interior_slices(self.cfg)
        for f in range(self.cfg.nvar):
            self.field_view(f)[sl] = data[f]

    def get_interior(self):
        sl = (slice(None),) + interior_slices(self.cfg)
        return np.stack([self.field_view(f)[sl].copy() for f in range(self.cfg.nvar)])

    def halo_exchange(self):
        self.L.orc_halo_exchange(self.h)

    def compute_dt(self):
        c = self.cfg
        return self.L.orc_compute_dt(self.h, c.eq, c.gamma, c.cfl, self._len)

    def time_step(self, dt):
        c = self.cfg
        self.L.orc_time_step(self.h, c.eq, c.gamma, dt, self._len)

    def advance_batch(self, steps, remaining=1.7976931348623157e308):
        c = self.cfg
        ex = C.c_size_t(0)
        dts = (C.c_double * max(steps, 1))()
        acc = self.L.orc_advance_batch(self.h, c.eq, c.gamma, c.cfl, self._len, steps, remaining,
                                       C.byref(ex), dts)
        return acc, ex.value, np.array(dts[:ex.value])

    def advance(self):
        acc, _, _ = self.advance_batch(1)
        return acc


# ---------------------------------------------------------------------------- script runner
def probe_pattern(cfg, P):
    """The exact-integer index probe of ref_dump.cpp op P: (f*2^20 + patch)*2^12 + cell."""
    f = np.arange(cfg.nvar, dtype=np.float64).reshape(-1, 1, 1)
    p = np.arange(P, dtype=np.float64).reshape(1, -1, 1)
    c = np.arange(cfg.flat, dtype=np.float64).reshape(1, 1, -1)
    return ((f * 2.0 ** 20 + p) * 4096.0 + c).reshape((cfg.nvar, P) + (cfg.psize,) * cfg.rank)


def run_script(tree, script, ic_override=None):
    """Run a ref_dump.cpp script against any backend exposing the OracleTree interface.
    Returns {tag/ids, tag/rel, tag/nbr, tag/quad, tag/data, tag/dts}."""
    cfg = tree.cfg
    out, dts = {}, []
    for line in script.strip().splitlines():
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        op = tok[0]
        if op == "A":
            tree.reconstruct(flags_all(tree.ids()))
        elif op == "H":
            seed, pr, pc, minl, maxl = (int(t) for t in tok[1:6])
            tree.reconstruct(flags_hash(tree.ids(), seed, pr, pc, minl, maxl))
        elif op == "B":
            r, minl, maxl = float(tok[1]), int(tok[2]), int(tok[3])
            centre = [float(t) for t in tok[4:4 + cfg.rank]]
            tree.reconstruct(flags_ball(tree.ids(), cfg, r, minl, maxl, centre))
        elif op in ("R", "K"):
            want = np.array([int(t) for t in tok[1:]], dtype=np.uint64)
            f = np.zeros(tree.size, dtype=np.int8)
            f[np.isin(tree.ids(), want)] = REFINE if op == "R" else COARSEN
            tree.reconstruct(f)
        elif op == "X":
            tree.halo_exchange()
        elif op == "P":
            tree.set_padded(probe_pattern(cfg, tree.size))
        elif op == "I":
            if ic_override is not None:
                tree.set_interior(ic_override(tree))
            else:
                tree.set_interior(initial_condition(tree.ids(), cfg))
        elif op == "S":
            for _ in range(int(tok[1])):
                dts.append(tree.advance())
        elif op == "D":
            tag = tok[1]
            rel, nbr, quad = tree.tables()
            out[tag + "/ids"] = tree.ids()
            out[tag + "/rel"], out[tag + "/nbr"], out[tag + "/quad"] = rel, nbr, quad
            out[tag + "/data"] = tree.get_padded().reshape(cfg.nvar, tree.size, cfg.flat)
            out[tag + "/dts"] = np.array(dts)
        else:
            raise ValueError("unknown op %r" % op)
    return out
