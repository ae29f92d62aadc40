"""Host members of include/ndtree/intergrid_operator.hpp (reference include/ndtree/intergrid_operator.hpp:41-106):
injection coarse -> fine, block injection, mean of the children fine -> coarse.  Host only."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_linear_interpolator_host_members(tmp_path):
    cxx = next((p for p in map(shutil.which, ("/usr/bin/g++-13", "g++-13", "g++")) if p), None)
    if cxx is None:
        pytest.skip("no C++ compiler")
    exe = tmp_path / "intergrid_host_check"
    r = subprocess.run([cxx, "-std=c++23", "-O1", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(HERE, "intergrid_host_check.cpp"), "-o", str(exe)], capture_output=True, text=True)
    if r.returncode != 0 and "c++23" in r.stderr:
        pytest.skip("compiler without C++23")
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    fine, back = out.split("|")
    assert [float(v) for v in fine.split()] == [3.25, 3.25, 3.25, 3.25, 0.0, -2.0, 0.0, 0.0]
    assert [float(v) for v in back.split()] == [(7.0 + 8.0 + 9.0 + 10.0) / 4, (3.25 * 3 - 2.0) / 4]
