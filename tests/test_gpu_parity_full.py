"""GPU parity AT THE TIMED CONFIGURATIONS, against the unmodified reference itself.

`oracle/_ref/ref_bench_2d` / `ref_bench_3d` are the reference's own headers (amr_solver::advance on the
reference's CPU path) compiled by oracle/Makefile with the reference's Release flags; they travel to the
GPU box as prebuilt binaries (they do not read /root/reference at run time).  Each test runs the binary on
the full-size mesh of a bench configuration, reads its dump (leaf ids, neighbor tables, padded state, dt
sequence) and compares the CUDA path through the C ABI on the same mesh, initial state and step count:

  * C2 (bench_fvm_solver_integration: 2 272 patches of 64x64 Euler cells, levels 5-7, 9.3e6 cells) with
    the kernel instantiation the bench line times (variant 0 -> band chosen per launch = 64 rows at this
    size) and the 16- / 32-row bands, the block-cooperative pipeline and the thread-per-cell kernel;
  * C3 patch shape (8^3 Euler, halo 1) on a three-level mesh, every 3D kernel generation.

Bars: ids / relations / neighbor indices / quadrants bit-exact; dt sequence rtol 1e-12; state
field-max-normalised 1e-12 (BASELINE.json north_star)."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import refdump_io
from golden_util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _reference_dump(binary, script, tmp, capacity):
    path = os.path.join(REF, binary)
    if not os.path.exists(path):
        pytest.skip("%s was not built (oracle/Makefile needs the reference checkout)" % binary)
    sp, op = os.path.join(tmp, "script.txt"), os.path.join(tmp, "dump.bin")
    open(sp, "w").write(script + "\n")
    subprocess.run([path, sp, op, str(capacity)], check=True, capture_output=True, timeout=1200)
    d = refdump_io.load(op)
    os.remove(op)
    return d


def _device_run(amrb, cfg, script, g, capacity, mode, variant, storage=None):
    """the same script through the C ABI; the initial interior is taken from the reference's own t0 dump
    (solver.initialize evaluates exp() with the reference's libm under -ffast-math)"""
    kw = {} if storage is None else {"storage": storage}
    tree = amrb.DeviceTree(cfg, capacity=capacity, mode=mode, **kw)
    tree.pool.set_variant(variant)

    def ic(t):
        d = g["t0/data"].reshape((cfg.nvar, t.size) + (cfg.psize,) * cfg.rank)
        return d[(slice(None), slice(None)) + O.interior_slices(cfg)]

    out = O.run_script(tree, script, ic_override=ic)
    return tree, out


def _compare(cfg, out, g, tags, what):
    mask = O.face_halo_mask(cfg).ravel()
    for tag in tags:
        for k in ("ids", "rel", "nbr", "quad"):
            assert np.array_equal(out[tag + "/" + k], g[tag + "/" + k]), (what, tag, k)
        np.testing.assert_allclose(out[tag + "/dts"], g[tag + "/dts"], rtol=TOL, atol=0, err_msg=str(what))
        err = rel_err(out[tag + "/data"][..., mask], g[tag + "/data"][..., mask])
        assert err <= TOL, (what, tag, err)


# ------------------------------------------------------------------------------------------ C2
C2_STEPS = 12


@pytest.fixture(scope="module")
def c2_reference(tmp_path_factory):
    import importlib

    wl = importlib.import_module("gpu-amr_b200.workloads")
    script = "\n".join([wl.c2_script(), "I", "X", "D t0", "S %d" % C2_STEPS, "D t1"])
    g = _reference_dump("ref_bench_2d", script, str(tmp_path_factory.mktemp("c2ref")), 4096)
    assert len(g["t0/ids"]) == 2272, "the reference built a different C2 mesh"
    return script, g


# (mode, variant): 0/0 = what bench.py times (band picked per launch: 64 rows for 2272 patches),
# 0/1, 0/2 = 16- / 32-row bands, 0/10 = block-cooperative pipeline, 2/0 = thread per cell
@pytest.mark.parametrize("mode,variant", [(0, 0), (0, 1), (0, 2), (0, 3), (0, 10), (2, 0)],
                         ids=["bench_kernel", "band16", "band32", "band64", "blockcoop", "threadpercell"])
def test_c2_full_size_matches_reference(amrb, c2_reference, mode, variant):
    script, g = c2_reference
    cfg = O.Config.from_name("r2_s64_h1_d7_euler")
    tree, out = _device_run(amrb, cfg, script, g, 4096, mode, variant)
    assert tree.size == 2272
    _compare(cfg, out, g, ("t0", "t1"), ("C2", mode, variant))
    # refinement criterion of the active-AMR benchmark (max of rho over ALL flat cells,
    # bench_fvm_solver_integration_active_amr.b.cpp:71-100) against the REFERENCE's padded state
    rho = g["t1/data"][0].reshape(tree.size, -1)
    mx = np.maximum(rho.max(axis=1), 0.0)
    lv = (g["t1/ids"] & np.uint64(63)).astype(int)
    for thr_r, thr_c in ((0.512, 0.506), (0.53, 0.501)):
        assert np.abs(mx - thr_r).min() > 1e-9 and np.abs(mx - thr_c).min() > 1e-9
        want = np.zeros(tree.size, np.int8)
        want[(lv < 7) & (mx > thr_r)] = 1
        want[(want == 0) & (lv > 1) & (mx < thr_c)] = 2
        got = tree.pool.patch_max_flags(0, thr_r, thr_c, 1, 7)
        assert np.array_equal(got, want), ("criterion", thr_r, thr_c)
    assert want.any()
    tree.pool.close()


# ------------------------------------------------------------------------------------------ C3 shape
C3_STEPS = 8
C3_SCRIPT = "\n".join(["A\nX"] * 3 + ["B 0.3 99 8 0.5 0.5 0.5", "X", "B 0.15 99 8 0.5 0.5 0.5", "X",
                                      "I", "X", "D t0", "S %d" % C3_STEPS, "D t1"])


@pytest.fixture(scope="module")
def c3_reference(tmp_path_factory):
    g = _reference_dump("ref_bench_3d", C3_SCRIPT, str(tmp_path_factory.mktemp("c3ref")), 20000)
    levels = set((g["t0/ids"] & np.uint64(63)).astype(int).tolist())
    assert levels == {3, 4, 5}, levels
    return g


# storage 1 = interior-only device layout (the C3 bench line), 0 = the reference's padded device layout
@pytest.mark.parametrize("storage,mode,variant", [(1, 0, 0), (1, 0, 21), (1, 0, 28), (1, 0, 41), (1, 0, 44), (0, 0, 0), (0, 0, 11), (0, 0, 10), (0, 2, 0), (0, 1, 0)],
                         ids=["dense_bench_kernel", "dense_ring_2x3", "dense_pingpong", "dense_prefetch", "dense_prefetch_oneblock_earlyz",
                              "padded_march", "padded_march_1plane", "padded_blockcoop",
                              "padded_threadpercell", "padded_unfused"])
def test_c3_shape_three_levels_matches_reference(amrb, c3_reference, storage, mode, variant):
    g = c3_reference
    cfg = O.Config.from_name("r3_s8_h1_d8_euler")
    tree, out = _device_run(amrb, cfg, C3_SCRIPT, g, 20000, mode, variant, storage=storage)
    _compare(cfg, out, g, ("t0", "t1"), ("C3", storage, mode, variant))
    tree.pool.close()
