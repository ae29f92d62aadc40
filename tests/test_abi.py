"""The drop-in boundary: the C-ABI shared library loads and exports every symbol declared in
include/gpuamr_b200.h; compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes as C

import pytest


def test_library_exports_every_declared_symbol(amrb):
    L = amrb.lib()
    names = amrb.declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_layout_queries(amrb):
    L = amrb.lib()
    lay = amrb.make_layout(2, 64, 1, amrb.EQ_EULER, 7)
    assert L.amrb_layout_flat_size(C.byref(lay)) == 66 * 66
    assert L.amrb_layout_data_size(C.byref(lay)) == 64 * 64
    assert L.amrb_layout_supported(C.byref(lay)) == 1
    lay = amrb.make_layout(3, 8, 1, amrb.EQ_EULER, 5)
    assert L.amrb_layout_flat_size(C.byref(lay)) == 1000
    assert L.amrb_layout_supported(C.byref(lay)) == 1
    odd = amrb.make_layout(2, 12, 1, amrb.EQ_EULER, 7)
    assert L.amrb_layout_supported(C.byref(odd)) == 0
    # interior-only device storage: rank 3, halo 1, 8^3 and 16^3 patches, both equations
    for size in (8, 16):
        for eq in (amrb.EQ_EULER, amrb.EQ_ADVECTION):
            d = amrb.make_layout(3, size, 1, eq, 8, amrb.STORAGE_INTERIOR)
            assert L.amrb_layout_supported(C.byref(d)) == 1
            assert L.amrb_layout_storage_size(C.byref(d)) == size ** 3
            assert L.amrb_layout_flat_size(C.byref(d)) == (size + 2) ** 3
    assert L.amrb_layout_storage_size(C.byref(lay)) == 1000
    for bad in (amrb.make_layout(2, 64, 1, amrb.EQ_EULER, 7, amrb.STORAGE_INTERIOR),
                amrb.make_layout(3, 4, 1, amrb.EQ_EULER, 5, amrb.STORAGE_INTERIOR),
                amrb.make_layout(3, 8, 1, amrb.EQ_EULER, 5, 7)):
        assert L.amrb_layout_supported(C.byref(bad)) == 0


def test_morton_roundtrip(amrb):
    L = amrb.lib()
    for rank in (2, 3):
        for coords, level in (((0, 0, 0), 0), ((64, 0, 32), 2), ((5, 9, 3), 7), ((127, 127, 127), 7)):
            c = (C.c_uint32 * 3)(*coords[:rank], *([0] * (3 - rank)))
            mid = L.amrb_morton_encode(rank, c, level)
            assert mid & 63 == level
            out = (C.c_uint32 * 3)()
            lv = C.c_int()
            L.amrb_morton_decode(rank, mid, out, C.byref(lv))
            assert lv.value == level and tuple(out[:rank]) == tuple(coords[:rank])
    # x is bit 0 of each interleaved group (morton_id.hpp), id = morton << 6 | level
    c = (C.c_uint32 * 3)(1, 0, 0)
    assert L.amrb_morton_encode(2, c, 7) == (1 << 6) | 7
    c = (C.c_uint32 * 3)(0, 1, 0)
    assert L.amrb_morton_encode(2, c, 7) == (2 << 6) | 7


def test_no_cpu_fallback(amrb):
    L = amrb.lib()
    if L.amrb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    lay = amrb.make_layout(2, 8, 1, amrb.EQ_EULER, 7)
    h = C.c_void_p()
    st = L.amrb_pool_create(C.byref(lay), 16, 0, C.byref(h))
    assert st != 0 and b"no CUDA device" in L.amrb_last_error()
    with pytest.raises(amrb.AmrbError):
        amrb.DevicePool(lay, 16)
