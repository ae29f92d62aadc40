"""INTEGRATION.md route 2, built and run: the reference's OWN headers (ndtree.hpp, amr_solver.hpp with
-DAMR_ENABLE_CUDA_AMR=1) and the reference's CUDA library minus its two hot-path translation units
(src/cuda/halo_exchange.cu, src/cuda/fvm_time_step.cu), which integration/amrb_shim.cpp replaces on top of the
C ABI (include/gpuamr_b200.h section 9: amrb_raw_*).  oracle/Makefile links the scripted dump driver this way
(oracle/_ref/shim_dump_*; prebuilt, travels to the GPU box); its dumps must reproduce the golden fixtures of
the unmodified reference: tables and probe halos bit-exact (the reference's 36-byte halo metadata is converted
on the device), states and dt sequences within 1e-12."""
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


@pytest.mark.parametrize("name", ["c3_euler", "c3_amr", "ka2d", "c1_adv_h2"])
def test_reference_headers_over_the_shim(name, tmp_path):
    cfg, script, g = load(name)
    binp = os.path.join(ROOT, "oracle", "_ref", "shim_dump_" + cfg.name)
    if not os.path.exists(binp):
        pytest.skip("%s was not built (oracle/Makefile needs the reference checkout)" % os.path.basename(binp))
    sp, op = str(tmp_path / "s.txt"), str(tmp_path / "o.bin")
    open(sp, "w").write(script + "\n")
    r = subprocess.run([binp, sp, op, "4096"], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    out = refdump_io.load(op)
    mask = O.face_halo_mask(cfg).ravel()
    for tag in tags_in_order(script):
        for k in ("ids", "rel", "nbr", "quad"):
            assert np.array_equal(out[tag + "/" + k], g[tag + "/" + k]), (name, tag, k)
        np.testing.assert_allclose(out[tag + "/dts"], g[tag + "/dts"], rtol=1e-12, atol=0)
        if tag + "/data" not in g:
            continue
        mine, ref = out[tag + "/data"][..., mask], g[tag + "/data"][..., mask]
        if _is_probe_tag(script, tag):
            assert np.array_equal(mine, ref), (name, tag, "halo indexing must be bit-exact")
        else:
            assert rel_err(mine, ref) <= 1e-11, (name, tag, rel_err(mine, ref))
