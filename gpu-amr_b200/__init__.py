"""gpu-amr_b200: B200-native (sm_100a) implementation of gpu-amr's per-timestep finite-volume hot
path (halo fill -> face-flux stencil -> conservative update over the Morton-ordered patch store).

The product is the CUDA library lib/libgpuamr_b200.so behind the C ABI in
include/gpuamr_b200.h plus the C++ host headers under include/.  This Python package is the thin
ctypes harness used by tests, bench.py and __graft_entry__.py.  Import it with
importlib.import_module("gpu-amr_b200") (the hyphen comes from the reference's repository name).
"""
from .binding import (AmrbError, DevicePool, DeviceTree, HostTree, Layout, build, check,  # noqa: F401
                      declared_symbols, lib, make_layout, EQ_ADVECTION, EQ_EULER, STABLE, REFINE,
                      COARSEN, DBL_MAX, LIB_PATH, STORAGE_PADDED, STORAGE_INTERIOR)
