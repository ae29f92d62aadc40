"""Drop-in check of BASELINE config C1: the reference's UNMODIFIED examples/fvm_solver_advection.e.cpp
(2D advection, 10x10 patches, halo 2, refine/coarsen every 5 steps, a VTK dump every 20 steps) is
built twice — against the reference's own headers on its CPU path (oracle/_ref/, the checker) and
against include/ of this repo + libgpuamr_b200.so (examples/_build/, the product) — and both are run.
Every VTK file must exist in both runs with identical headers, geometry (POINTS, CELLS, CELL_TYPES),
cell_index and is_halo arrays bit for bit (so every refinement decision along 2115 steps agreed),
and the field within 1e-12 of the field maximum."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "examples", "_build", "ref_example_fvm_solver_advection")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_example_fvm_solver_advection")


def parse_vtk(path):
    """legacy VTK BINARY unstructured grid -> dict of header text and big-endian arrays"""
    raw = open(path, "rb").read()
    pos, out = 0, {"text": []}

    def line():
        nonlocal pos
        end = raw.index(b"\n", pos)
        s = raw[pos:end].decode()
        pos = end + 1
        return s

    def block(dtype, count):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += a.nbytes
        return a

    for _ in range(4):
        out["text"].append(line())
    m = re.match(r"POINTS (\d+) (\w+)", line())
    npts, real = int(m.group(1)), {"double": ">f8", "float": ">f4"}[m.group(2)]
    out["points"] = block(real, 3 * npts)
    m = re.match(r"CELLS (\d+) (\d+)", line())
    ncell = int(m.group(1))
    out["cells"] = block(">i4", int(m.group(2)))
    assert line() == "CELL_TYPES %d" % ncell
    out["cell_types"] = block(">i4", ncell)
    assert line() == "CELL_DATA %d" % ncell
    while pos < len(raw):
        m = re.match(r"SCALARS (\S+) (\w+) 1", line())
        assert line() == "LOOKUP_TABLE default"
        dt = {"double": ">f8", "float": ">f4", "int": ">i4"}[m.group(2)]
        out["scalar:" + m.group(1)] = block(dt, ncell)
    return out


def test_vtk_parser_roundtrip(tmp_path):
    # tiny hand-made file: the parser itself is test infrastructure and is checked first
    p = tmp_path / "t.vtk"
    with open(p, "wb") as f:
        f.write(b"# vtk DataFile Version 3.0\nAMR Tree Structure\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(b"POINTS 4 double\n" + np.arange(12, dtype=">f8").tobytes())
        f.write(b"CELLS 1 5\n" + np.array([4, 0, 1, 2, 3], dtype=">i4").tobytes())
        f.write(b"CELL_TYPES 1\n" + np.array([9], dtype=">i4").tobytes())
        f.write(b"CELL_DATA 1\nSCALARS u double 1\nLOOKUP_TABLE default\n" + np.array([2.5], dtype=">f8").tobytes())
    v = parse_vtk(str(p))
    assert v["points"][11] == 11.0 and v["cells"][0] == 4 and v["scalar:u"][0] == 2.5


@pytest.mark.gpu
def test_reference_advection_example_is_a_drop_in(tmp_path):
    for b in (OURS, REF):
        if not os.path.exists(b):
            pytest.skip("%s was not built (needs the reference checkout at build time)" % b)
    runs = {}
    for name, binary in (("ours", OURS), ("ref", REF)):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([binary], cwd=str(d), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        runs[name] = (d / "vtk_output", r.stdout)
    files = {k: sorted(os.listdir(v[0])) for k, v in runs.items()}
    assert files["ours"] == files["ref"] and len(files["ref"]) == 105
    # the printed step sizes agree (6 significant digits are printed)
    steps = {k: re.findall(r"Step (\d+), t=(\S+), dt=(\S+)", v[1]) for k, v in runs.items()}
    assert steps["ours"] == steps["ref"] and len(steps["ref"]) == 2115
    worst = 0.0
    for fn in files["ref"]:
        a, b = parse_vtk(str(runs["ours"][0] / fn)), parse_vtk(str(runs["ref"][0] / fn))
        assert a["text"] == b["text"] and a.keys() == b.keys(), fn
        for k in ("points", "cells", "cell_types", "scalar:cell_index", "scalar:is_halo"):
            assert np.array_equal(a[k], b[k]), (fn, k)
        fields = [k for k in b if k.startswith("scalar:") and k not in ("scalar:cell_index", "scalar:is_halo")]
        assert len(fields) == 1
        for k in fields:
            err = np.abs(a[k] - b[k]).max() / np.abs(b[k]).max()
            worst = max(worst, err)
            assert err <= 1e-12, (fn, k, err)
    print("105 VTK files identical in structure; worst field error %.2e" % worst)
