"""Drop-in check of the C++ boundary: the reference's UNMODIFIED benchmark drivers
(benchmark/bench_fvm_solver_integration{,3D}.b.cpp, compiled against include/ of this repo with
-DAMR_ENABLE_CUDA_AMR and linked to libgpuamr_b200.so by examples/Makefile) must reproduce the
counts of the reference's own CPU run recorded in BASELINE.md section 2 (cell updates, solver steps,
patch counts): every CFL step size, refinement decision, balancing ripple and restriction /
prolongation along 2114 (2D) and 362 (3D) steps has to agree for these integers to match.
Also runs our port of the reference's gtest property checks (tests/advection_equation_amr.t.cpp)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "examples", "_build")


def _run(name, timeout=900):
    path = os.path.join(BUILD, name)
    if not os.path.exists(path):
        pytest.skip("%s was not built (examples/Makefile needs the reference checkout)" % name)
    r = subprocess.run([path], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


def _field(out, label):
    return int(re.search(re.escape(label) + r":\s*(\d+)", out).group(1))


def test_reference_2d_benchmark_driver_reproduces_cpu_run():
    out = _run("ref_bench_fvm_solver_integration")
    assert "CUDA ENABLED" in out
    assert _field(out, "Updated cells") == 9293201408          # reference CPU SEQ run
    assert _field(out, "Solver timesteps") == 2114
    assert (_field(out, "Min patch count"), _field(out, "Max patch count")) == (64, 1372)


def test_reference_3d_benchmark_driver_reproduces_cpu_run():
    out = _run("ref_bench_fvm_solver_integration3D")
    assert _field(out, "Updated cells") == 2349748224
    assert _field(out, "Solver timesteps") == 362
    assert (_field(out, "Min patch count"), _field(out, "Max patch count")) == (64, 23136)


def test_reference_active_amr_benchmark_driver_reproduces_cpu_run():
    """BASELINE config C4 (bench_fvm_solver_integration_active_amr.b.cpp: reconstruct_tree every 2 steps,
    thresholds 0.512 / 0.506, levels 1-7, two pulses): 1 313 reconstructs, 351 of them changing the topology,
    every refinement decision taken by the device criterion.  Golden = the reference's own CPU run
    (tests/golden/counts/, 65 minutes on one core)."""
    gold = open(os.path.join(ROOT, "tests", "golden", "counts",
                             "ref_bench_fvm_solver_integration_active_amr.cpu.txt")).read()
    out = _run("ref_bench_fvm_solver_integration_active_amr")
    assert "CUDA ENABLED" in out and "CUDA DISABLED" in gold
    for label in ("Updated cells", "Solver timesteps", "Initial reconstructions", "Timed reconstructions",
                  "Identity reconstructions", "Topology-changing reconstructions", "Final patch count",
                  "Min patch count", "Max patch count"):
        assert _field(out, label) == _field(gold, label), label
    assert _field(gold, "Updated cells") == 39271899136 and _field(gold, "Topology-changing reconstructions") == 351


def test_advection_pulse_property_checks():
    out = _run("advection_pulse_check")
    assert "ALL OK" in out and out.count(": OK") == 2
