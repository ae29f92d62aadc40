"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys the
driver reads (impl, metric, unit, value, cpu_baseline, e2e with zero copy bytes, config.workload), under a
multi-rank launch only rank 0 prints, and our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True,
                          text=True, timeout=timeout, cwd=ROOT, env=e)


import pytest


@pytest.mark.parametrize("workload", ["c3", "c2"])
def test_reference_arm_prints_the_contract_line(workload):
    # --steps 50 selects the smallest bounded sample of the C3 family (2.1e6 cells): seconds, not minutes
    steps = 50 if workload == "c3" else 1
    r = _run(["--impl", "reference", "--steps", str(steps), "--warmup", "0", "--workload", workload])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_sec" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == steps
    assert d["value"] > 1e6 and d["dtype"] == "f64" and d["data"] == "synthetic"
    if workload == "c2":
        assert d["config"]["workload"].startswith("C2") and d["config"]["cells"] == 9306112 and d["scaling"] == "weak"
    else:
        assert d["config"]["workload"].startswith("C3") and d["scaling"] == "strong"
        assert d["config"]["cells"] == 1043062784 and d["config"]["sample_cells"] >= 1e6
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    try:
        import torch

        if torch.cuda.is_available():
            return
    except Exception:
        pass
    r = _run(["--steps", "2", "--warmup", "1", "--no-cpu-baseline"], timeout=300)
    assert r.returncode != 0 and "{" not in r.stdout
