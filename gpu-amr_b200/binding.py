"""ctypes binding of the C ABI in include/gpuamr_b200.h (lib/libgpuamr_b200.so).

This is the Python-side harness used by tests/, bench.py and __graft_entry__.py; the product
host layer is the C++ headers under include/ (ndtree / solver mirrors) which call the very same
entry points.  Nothing here computes: every operation is a call into the CUDA library, and
loading fails loudly when the library has not been built (no CPU fallback).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libgpuamr_b200.so")
HEADER = os.path.join(ROOT, "include", "gpuamr_b200.h")

EQ_ADVECTION, EQ_EULER = 0, 1
STORAGE_PADDED, STORAGE_INTERIOR = 0, 1
STABLE, REFINE, COARSEN = 0, 1, 2
DBL_MAX = 1.7976931348623157e308

_LIB = None


class AmrbError(RuntimeError):
    pass


class Layout(C.Structure):
    _fields_ = [("rank", C.c_int32), ("size", C.c_int32 * 3), ("halo", C.c_int32),
                ("nvar", C.c_int32), ("equation", C.c_int32), ("depth", C.c_int32),
                ("storage", C.c_int32)]


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "-j4"], stdout=out)


def declared_symbols():
    """Every function name declared in include/gpuamr_b200.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(amrb_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise AmrbError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, i32p, i8p, dp = C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p
    L.amrb_last_error.restype = C.c_char_p
    L.amrb_version.restype = C.c_char_p
    L.amrb_device_count.restype = C.c_int
    sig = {
        "amrb_device_malloc": [C.POINTER(vp), sz],
        "amrb_device_free": [vp],
        "amrb_host_pinned_malloc": [C.POINTER(vp), sz],
        "amrb_host_pinned_free": [vp],
        "amrb_copy_host_to_device": [vp, vp, sz],
        "amrb_copy_host_to_device_async": [vp, vp, sz, vp],
        "amrb_copy_device_to_host": [vp, vp, sz],
        "amrb_copy_device_to_host_async": [vp, vp, sz, vp],
        "amrb_copy_device_to_device": [vp, vp, sz],
        "amrb_stream_create": [C.POINTER(vp)],
        "amrb_stream_destroy": [vp],
        "amrb_stream_synchronize": [vp],
        "amrb_stream_wait_fence": [vp, vp],
        "amrb_fence_create": [C.POINTER(vp)],
        "amrb_fence_destroy": [vp],
        "amrb_fence_record": [vp, vp],
        "amrb_fence_wait": [vp],
        "amrb_device_synchronize": [],
        "amrb_pool_create": [C.POINTER(Layout), sz, C.c_int, C.POINTER(vp)],
        "amrb_pool_create_external": [C.POINTER(Layout), sz, C.c_int, C.POINTER(vp),
                                      C.POINTER(vp), vp, C.POINTER(vp)],
        "amrb_pool_destroy": [vp],
        "amrb_pool_set_topology": [vp, sz, sz, i32p, i8p, i32p, i8p],
        "amrb_pool_set_topology_from_ids": [vp, C.POINTER(C.c_uint64), sz],
        "amrb_pool_get_tables": [vp, i32p, C.POINTER(C.c_uint8), i32p],
        "amrb_pool_set_physics": [vp, C.POINTER(C.c_double), C.c_double, C.c_double],
        "amrb_pool_upload": [vp, C.c_int, sz, sz, dp],
        "amrb_pool_download": [vp, C.c_int, sz, sz, dp],
        "amrb_pool_upload_next": [vp, C.c_int, sz, sz, dp],
        "amrb_pool_upload_interior": [vp, C.c_int, sz, sz, dp],
        "amrb_pool_download_interior": [vp, C.c_int, sz, sz, dp],
        "amrb_pool_halo_exchange": [vp],
        "amrb_pool_swap_buffers": [vp],
        "amrb_pool_set_lazy_halos": [vp, C.c_int],
        "amrb_pool_ensure_halos": [vp],
        "amrb_pool_compute_dt": [vp, C.POINTER(C.c_double)],
        "amrb_pool_step": [vp, C.c_double],
        "amrb_pool_advance_batch_async": [vp, sz, C.c_double],
        "amrb_pool_finish_advance_batch": [vp, C.POINTER(C.c_double), C.POINTER(sz), dp, sz],
        "amrb_pool_set_mode": [vp, C.c_int],
        "amrb_pool_set_variant": [vp, C.c_int],
        "amrb_pool_mark_dirty": [vp],
        "amrb_pool_batch_begin": [vp, sz, C.c_double],
        "amrb_pool_step_partial": [vp, i32p, sz],
        "amrb_pool_step_commit": [vp],
        "amrb_pool_batch_end": [vp, C.c_int],
        "amrb_pool_pack_faces": [vp, i32p, sz, dp],
        "amrb_pool_unpack_faces": [vp, i32p, sz, dp],
        "amrb_exchange_create": [vp, C.c_int, C.c_int, i32p, vp, vp, i32p, sz, C.POINTER(vp)],
        "amrb_exchange_destroy": [vp],
        "amrb_ipc_export": [vp, vp],
        "amrb_ipc_open": [vp, C.POINTER(vp)],
        "amrb_ipc_close": [vp],
        "amrb_exchange_connect": [vp, C.c_int, vp, vp, vp],
        "amrb_exchange_halo": [vp],
        "amrb_exchange_advance_batch_async": [vp, sz, C.c_double, C.c_int],
        "amrb_exchange_set_lists": [vp, i32p, sz, i32p, sz],
        "amrb_exchange_set_timing": [vp, C.c_int],
        "amrb_exchange_get_timing": [vp, C.POINTER(C.c_double)],
        "amrb_exchange_push": [vp, C.c_int, sz],
        "amrb_exchange_wait": [vp, C.c_int, sz],
        "amrb_pool_apply_plan": [vp, sz, i8p, i32p, i8p],
        "amrb_pool_flag_patches": [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int],
        "amrb_pool_reconstruct_device": [vp, vp, C.POINTER(C.c_int), C.POINTER(sz)],
        "amrb_pool_get_ids": [vp, C.POINTER(C.c_uint64), sz],
        "amrb_pool_get_plan": [vp, i8p, i32p, i8p],
        "amrb_tree_assign": [vp, C.POINTER(C.c_uint64), sz],
        "amrb_pool_patch_max_flags": [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, i8p],
        "amrb_patch_max_flags_device": [vp, vp, sz, sz, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp],
        "amrb_profile_capture_start": [],
        "amrb_profile_capture_stop": [],
        "amrb_profile_range_push": [C.c_char_p],
        "amrb_profile_range_pop": [],
        "amrb_tree_create": [C.c_int, C.c_int, C.POINTER(vp)],
        "amrb_tree_destroy": [vp],
        "amrb_tree_reconstruct": [vp, i8p, sz, C.POINTER(C.c_int)],
        "amrb_tree_plan": [vp, i8p, i32p, i8p],
        "amrb_tree_tables": [vp, i32p, i8p, i32p, i8p],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.amrb_layout_supported.argtypes = [C.POINTER(Layout)]
    L.amrb_layout_supported.restype = C.c_int
    for name in ("amrb_layout_flat_size", "amrb_layout_data_size", "amrb_layout_storage_size"):
        getattr(L, name).argtypes = [C.POINTER(Layout)]
        getattr(L, name).restype = sz
    for name in ("amrb_pool_capacity", "amrb_pool_size", "amrb_tree_size", "amrb_tree_plan_size"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = sz
    L.amrb_pool_stream.argtypes = [vp]
    L.amrb_pool_stream.restype = vp
    for name in ("amrb_pool_field", "amrb_pool_next_field"):
        getattr(L, name).argtypes = [vp, C.c_int]
        getattr(L, name).restype = vp
    L.amrb_pool_levels.argtypes = [vp]
    L.amrb_pool_levels.restype = vp
    L.amrb_pool_dtmin_slot.argtypes = [vp, sz]
    L.amrb_pool_dtmin_slot.restype = vp
    L.amrb_exchange_buffer.argtypes = [vp, C.c_int]
    L.amrb_exchange_buffer.restype = vp
    L.amrb_exchange_timed_out.argtypes = [vp]
    L.amrb_exchange_timed_out.restype = C.c_int
    L.amrb_exchange_launch_count.argtypes = [vp]
    L.amrb_exchange_launch_count.restype = C.c_uint64
    L.amrb_pool_launch_count.argtypes = [vp]
    L.amrb_pool_launch_count.restype = C.c_uint64
    L.amrb_pool_face_slab_doubles.argtypes = [vp, C.c_int]
    L.amrb_pool_face_slab_doubles.restype = sz
    L.amrb_tree_ids.argtypes = [vp]
    L.amrb_tree_ids.restype = C.POINTER(C.c_uint64)
    L.amrb_morton_encode.argtypes = [C.c_int, C.POINTER(C.c_uint32), C.c_int]
    L.amrb_morton_encode.restype = C.c_uint64
    L.amrb_morton_decode.argtypes = [C.c_int, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.amrb_morton_decode.restype = None
    _LIB = L
    return L


def check(status):
    if status != 0:
        raise AmrbError("amrb status %d: %s" % (status, lib().amrb_last_error().decode()))


def make_layout(rank, size, halo, eq, depth, storage=STORAGE_PADDED):
    lay = Layout()
    lay.rank, lay.halo, lay.equation, lay.depth, lay.storage = rank, halo, eq, depth, storage
    for k in range(3):
        lay.size[k] = size if k < rank else 1
    lay.nvar = 1 if eq == EQ_ADVECTION else rank + 2
    return lay


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostTree:
    """Morton-ordered leaf set + neighbor tables (C-ABI section 7); pure host, no GPU needed."""

    def __init__(self, rank, depth):
        self.L = lib()
        self.rank, self.depth = rank, depth
        self.ndir, self.kf = 2 * rank, 1 << (rank - 1)
        h = C.c_void_p()
        check(self.L.amrb_tree_create(rank, depth, C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.amrb_tree_destroy(self.h)
            self.h = None

    @property
    def size(self):
        return self.L.amrb_tree_size(self.h)

    def ids(self):
        return np.ctypeslib.as_array(self.L.amrb_tree_ids(self.h), shape=(self.size,)).copy()

    def reconstruct(self, flags, capacity=0):
        flags = np.ascontiguousarray(flags, dtype=np.int8)
        assert len(flags) == self.size
        ch = C.c_int(0)
        check(self.L.amrb_tree_reconstruct(self.h, _ptr(flags), capacity, C.byref(ch)))
        return ch.value

    def assign(self, ids):
        """adopt a leaf set computed on the device (DevicePool.reconstruct_device)"""
        ids = np.ascontiguousarray(ids, np.uint64)
        check(self.L.amrb_tree_assign(self.h, ids.ctypes.data_as(C.POINTER(C.c_uint64)), len(ids)))

    def plan(self):
        n = self.L.amrb_tree_plan_size(self.h)
        kind, src, child = np.zeros(n, np.int8), np.zeros(n, np.int32), np.zeros(n, np.int8)
        check(self.L.amrb_tree_plan(self.h, _ptr(kind), _ptr(src), _ptr(child)))
        return kind, src, child

    def tables(self):
        n = self.size
        levels = np.zeros(n, np.int32)
        rel = np.zeros((n, self.ndir), np.int8)
        nbr = np.zeros((n, self.ndir, self.kf), np.int32)
        quad = np.zeros((n, self.ndir, self.rank), np.int8)
        check(self.L.amrb_tree_tables(self.h, _ptr(levels), _ptr(rel), _ptr(nbr), _ptr(quad)))
        return levels, rel, nbr, quad


class DevicePool:
    """Device-resident SoA patch pool (C-ABI sections 3-6)."""

    def __init__(self, lay, capacity, device=0, external=None, stream=None):
        self.L = lib()
        self.lay = lay
        self.flat = self.L.amrb_layout_flat_size(C.byref(lay))      # padded patch (host exchange format)
        self.data = self.L.amrb_layout_data_size(C.byref(lay))
        self.stored = self.L.amrb_layout_storage_size(C.byref(lay))  # doubles per field-patch in the pool
        h = C.c_void_p()
        if external is None:
            check(self.L.amrb_pool_create(C.byref(lay), capacity, device, C.byref(h)))
        else:
            cur, nxt = external
            a = (C.c_void_p * lay.nvar)(*cur)
            b = (C.c_void_p * lay.nvar)(*nxt)
            check(self.L.amrb_pool_create_external(C.byref(lay), capacity, device, a, b,
                                                   C.c_void_p(stream or 0), C.byref(h)))
        self.h = h
        self.capacity = capacity

    def close(self):
        if getattr(self, "h", None):
            self.L.amrb_pool_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def size(self):
        return self.L.amrb_pool_size(self.h)

    def set_physics(self, lengths, gamma, cfl):
        a = (C.c_double * 3)(*(list(lengths) + [1.0] * 3)[:3])
        check(self.L.amrb_pool_set_physics(self.h, a, gamma, cfl))

    def set_topology(self, levels, rel, nbr, quad, n_total=None):
        levels = np.ascontiguousarray(levels, np.int32)
        rel = np.ascontiguousarray(rel, np.int8)
        nbr = np.ascontiguousarray(nbr, np.int32)
        quad = np.ascontiguousarray(quad, np.int8)
        n = len(levels)
        check(self.L.amrb_pool_set_topology(self.h, n, n if n_total is None else n_total,
                                            _ptr(levels), _ptr(rel), _ptr(nbr), _ptr(quad)))

    def set_topology_from_ids(self, ids):
        """tables built on the device from the ascending leaf ids (single GPU, no ghost slots)"""
        ids = np.ascontiguousarray(ids, np.uint64)
        check(self.L.amrb_pool_set_topology_from_ids(self.h, ids.ctypes.data_as(C.POINTER(C.c_uint64)), len(ids)))

    def get_tables(self, n, rank):
        """device tables in the compact device form: levels[n], meta[n][2R], nbr[n][2R][2^(R-1)]"""
        levels = np.zeros(n, np.int32)
        meta = np.zeros((n, 2 * rank), np.uint8)
        nbr = np.zeros((n, 2 * rank, 1 << (rank - 1)), np.int32)
        check(self.L.amrb_pool_get_tables(self.h, _ptr(levels), meta.ctypes.data_as(C.POINTER(C.c_uint8)), _ptr(nbr)))
        return levels, meta, nbr

    def upload(self, field, data, first=0):
        data = np.ascontiguousarray(data, np.float64).reshape(-1, self.flat)
        check(self.L.amrb_pool_upload(self.h, field, first, data.shape[0], _ptr(data)))

    def download(self, field, n, first=0):
        out = np.empty((n, self.flat), np.float64)
        check(self.L.amrb_pool_download(self.h, field, first, n, _ptr(out)))
        return out

    def upload_interior(self, field, data, first=0):
        data = np.ascontiguousarray(data, np.float64).reshape(-1, self.data)
        check(self.L.amrb_pool_upload_interior(self.h, field, first, data.shape[0], _ptr(data)))

    def download_interior(self, field, n, first=0):
        out = np.empty((n, self.data), np.float64)
        check(self.L.amrb_pool_download_interior(self.h, field, first, n, _ptr(out)))
        return out

    def halo_exchange(self):
        check(self.L.amrb_pool_halo_exchange(self.h))

    def compute_dt(self):
        v = C.c_double(0.0)
        check(self.L.amrb_pool_compute_dt(self.h, C.byref(v)))
        return v.value

    def step(self, dt):
        check(self.L.amrb_pool_step(self.h, dt))

    def advance_batch_async(self, steps, remaining=DBL_MAX):
        check(self.L.amrb_pool_advance_batch_async(self.h, steps, remaining))

    def finish_advance_batch(self, max_steps=0):
        s, n = C.c_double(0.0), C.c_size_t(0)
        dts = np.zeros(max(max_steps, 1), np.float64)
        check(self.L.amrb_pool_finish_advance_batch(self.h, C.byref(s), C.byref(n), _ptr(dts),
                                                    max_steps))
        return s.value, n.value, dts[:min(n.value, max_steps)]

    def set_mode(self, mode):
        check(self.L.amrb_pool_set_mode(self.h, mode))

    def set_variant(self, variant):
        check(self.L.amrb_pool_set_variant(self.h, variant))

    def mark_dirty(self):
        check(self.L.amrb_pool_mark_dirty(self.h))

    def apply_plan(self, kind, src, child):
        check(self.L.amrb_pool_apply_plan(self.h, len(kind), _ptr(kind), _ptr(src), _ptr(child)))

    def patch_max_flags(self, field, refine_thr, coarsen_thr, min_level, max_level):
        out = np.zeros(self.size, np.int8)
        check(self.L.amrb_pool_patch_max_flags(self.h, field, refine_thr, coarsen_thr, min_level,
                                               max_level, _ptr(out)))
        return out

    def flag_patches(self, field, refine_thr, coarsen_thr, min_level, max_level):
        """the criterion with the flags left on the device (no read-back)"""
        check(self.L.amrb_pool_flag_patches(self.h, field, refine_thr, coarsen_thr, min_level, max_level))

    def reconstruct_device(self, dev_flags=None):
        """reconstruct_tree on the device: returns (changed, new size)"""
        ch, n = C.c_int(0), C.c_size_t(0)
        check(self.L.amrb_pool_reconstruct_device(self.h, C.c_void_p(dev_flags or 0), C.byref(ch), C.byref(n)))
        return ch.value, n.value

    def get_ids(self):
        out = np.zeros(self.size, np.uint64)
        check(self.L.amrb_pool_get_ids(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), len(out)))
        return out

    def get_plan(self):
        n = self.size
        kind, src, child = np.zeros(n, np.int8), np.zeros(n, np.int32), np.zeros(n, np.int8)
        check(self.L.amrb_pool_get_plan(self.h, _ptr(kind), _ptr(src), _ptr(child)))
        return kind, src, child

    def launch_count(self):
        return int(self.L.amrb_pool_launch_count(self.h))

    def synchronize(self):
        check(self.L.amrb_stream_synchronize(self.L.amrb_pool_stream(self.h)))


class DeviceTree:
    """HostTree + DevicePool behind the interface the oracle's script runner drives
    (same method names as oracle.OracleTree, which mirrors ndtree + amr_solver)."""

    def __init__(self, cfg, capacity=20000, device=0, mode=0, storage=STORAGE_PADDED):
        self.cfg = cfg
        self.lay = make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage)
        self.tree = HostTree(cfg.rank, cfg.depth)
        self.pool = DevicePool(self.lay, capacity, device)
        self.pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
        self.pool.set_mode(mode)
        self.capacity = capacity
        self._push_topology()

    def _push_topology(self):
        # tables are built on the device from the leaf ids (AMRB_DEVICE_TOPOLOGY=0: host build + upload)
        if os.environ.get("AMRB_DEVICE_TOPOLOGY", "1") != "0":
            self.pool.set_topology_from_ids(self.tree.ids())
        else:
            levels, rel, nbr, quad = self.tree.tables()
            self.pool.set_topology(levels, rel, nbr, quad)

    @property
    def size(self):
        return self.tree.size

    def ids(self):
        return self.tree.ids()

    def reconstruct(self, flags):
        changed = self.tree.reconstruct(flags, self.capacity)
        if changed:
            self.pool.apply_plan(*self.tree.plan())
            self._push_topology()
        return changed

    def tables(self):
        _, rel, nbr, quad = self.tree.tables()
        return rel, nbr, quad

    def get_padded(self):
        cfg, n = self.cfg, self.size
        return np.stack([self.pool.download(f, n).reshape((n,) + (cfg.psize,) * cfg.rank)
                         for f in range(cfg.nvar)])

    def set_padded(self, data):
        for f in range(self.cfg.nvar):
            self.pool.upload(f, data[f])

    def set_interior(self, data):
        for f in range(self.cfg.nvar):
            self.pool.upload_interior(f, data[f])

    def get_interior(self):
        cfg, n = self.cfg, self.size
        return np.stack([self.pool.download_interior(f, n).reshape((n,) + (cfg.size,) * cfg.rank)
                         for f in range(cfg.nvar)])

    def halo_exchange(self):
        self.pool.halo_exchange()

    def compute_dt(self):
        return self.pool.compute_dt()

    def time_step(self, dt):
        self.pool.step(dt)
        self.pool.halo_exchange()

    def advance_batch(self, steps, remaining=DBL_MAX):
        self.pool.advance_batch_async(steps, remaining)
        return self.pool.finish_advance_batch(steps)

    def advance(self):
        return self.advance_batch(1)[0]
