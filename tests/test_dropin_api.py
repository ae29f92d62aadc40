"""Drop-in check of the ndtree / neighbor API surface (get_neighbor_at, neighbor_linear_index,
direction<Dim>, neighbor_variant, get_node_index_at, get_patch, sync_current_{to,from}_device,
build_patch_levels_on_device, reconstruct_tree, halo_exchange_update, amr_solver::initialize / advance):
oracle/ref_dump.cpp — the scripted driver that produced the golden fixtures from the UNMODIFIED reference —
is compiled UNCHANGED against include/ of this repo (examples/Makefile: dropin_dump_*) and run on the GPU;
its dumps must reproduce the reference's: leaf ids, relations, neighbor indices and contact quadrants
bit-exact (every one of them travels through the neighbor variants), probe halos bit-exact, states and dt
sequences within 1e-12."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import refdump_io
from golden_util import load, rel_err, tags_in_order
from test_gpu_parity import _is_probe_tag

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


@pytest.mark.parametrize("name", ["c3_euler", "c3_amr", "ka2d", "c1_adv_h2"])
def test_reference_dump_driver_against_own_headers(name, tmp_path):
    cfg, script, g = load(name)
    binp = os.path.join(ROOT, "examples", "_build", "dropin_dump_" + cfg.name)
    if not os.path.exists(binp):
        pytest.skip("%s was not built (examples/Makefile)" % os.path.basename(binp))
    sp, op = str(tmp_path / "s.txt"), str(tmp_path / "o.bin")
    open(sp, "w").write(script + "\n")
    r = subprocess.run([binp, sp, op, "4096"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    out = refdump_io.load(op)
    mask = O.face_halo_mask(cfg).ravel()
    for tag in tags_in_order(script):
        for k in ("ids", "rel", "nbr", "quad"):
            assert np.array_equal(out[tag + "/" + k], g[tag + "/" + k]), (name, tag, k)
        np.testing.assert_allclose(out[tag + "/dts"], g[tag + "/dts"], rtol=TOL, atol=0)
        if tag + "/data" not in g:
            continue
        mine, ref = out[tag + "/data"][..., mask], g[tag + "/data"][..., mask]
        if _is_probe_tag(script, tag):
            assert np.array_equal(mine, ref), (name, tag, "halo indexing must be bit-exact")
        else:
            # the driver evaluates the initial condition itself (libm exp vs the fixture's): rounding-level
            assert rel_err(mine, ref) <= 1e-11, (name, tag, rel_err(mine, ref))
