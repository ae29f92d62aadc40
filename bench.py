#!/usr/bin/env python
"""bench.py — cell-updates/sec of the fused halo + flux + update step on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...          (one rank per GPU, N > 1)

A "step" is one explicit time step of the finite-volume solver over the whole mesh (ghost-cell
gather, Rusanov face fluxes, conservative update, CFL reduction for the next step).  Workload at
N = 1: BASELINE.json configs[1], the 2D static multi-level tree of 64x64 Euler patches
(levels 5-7, 2272 patches, 9.3e6 cells; SURVEY 8d "C2").  Prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cell_updates_per_sec"
UNIT = "cell-updates/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.15 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path, timed on this box's host cores:
    oracle/_ref/ref_bench_2d = the UNMODIFIED reference headers (amr_solver::advance over the same
    C2 mesh and IC) built with the reference's Release flags and EXECUTION=PAR.  Without TBB
    libstdc++'s parallel policies run serially, so the reference can use 1 core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    amrb = importlib.import_module("gpu-amr_b200")
    from importlib import import_module
    wl = import_module("gpu-amr_b200.workloads")
    binp = os.path.join(ROOT, "oracle", "_ref", "ref_bench_2d")
    # one bench step = ONE solver step over a bounded sample of the workload: the full C2 mesh
    # (9.3e6 cells, ~0.5 s per step on one core) when the run stays within a few minutes, else the
    # geometrically similar mesh one or two base levels coarser (4x / 16x fewer cells; per-core
    # throughput is size-independent once the state is out of cache)
    n_total = args.steps + args.warmup
    base = 5 if n_total <= 300 else (4 if n_total <= 1200 else 3)
    if os.path.exists(binp):
        kind, cores = "reference", 1
        with tempfile.TemporaryDirectory() as td:
            sp = os.path.join(td, "s.txt")
            lines = [wl.c2_script(base), "I", "X"]
            lines += ["T %d" % args.warmup] if args.warmup else []
            lines += ["T %d" % args.steps]
            open(sp, "w").write("\n".join(lines) + "\n")
            out = subprocess.run([binp, sp, os.path.join(td, "o.bin"), "4096"], check=True,
                                 capture_output=True, text=True).stdout
        rec = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
        value, secs, cells = rec["updates_per_s"], rec["seconds"], rec["cells"]
    else:
        # the oracle port (OpenMP) — only when the reference binary did not travel
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        cfg = O.Config.from_name("r2_s64_h1_d7_euler")
        tree = O.OracleTree(cfg, capacity=4096)
        O.run_script(tree, wl.c2_script(base) + "\nI\nX")
        kind, cores = "port", O.lib().orc_num_threads()
        tree.advance_batch(args.warmup)
        t0 = time.time()
        tree.advance_batch(args.steps)
        secs = time.time() - t0
        cells = tree.size * cfg.size ** cfg.rank
        value = cells * args.steps / secs
    sample = ("1 amr_solver::advance() per bench step over the C2 mesh at base level %d (%d cells); "
              "unmodified reference headers, Release flags, EXECUTION=PAR = serial PSTL (no TBB)"
              % (base, cells))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cells, {"note": "bounded sample of the N-GPU arm's C2-family workload: the "
                                          "N = 1 C2 mesh (CPU throughput per core is size-independent)"}
                                  if args.gpus > 1 else None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(cells, extra=None, name=None):
    c = {"workload": ("development workload " + name) if name else
                     "C2: bench_fvm_solver_integration 2D static multi-level tree, Euler fp64, "
                     "64x64 patches halo 1, levels 5-7, acoustic pulse",
         "cells": int(cells), "l2_policy": "inputs larger than L2 (state 2 x 0.4 GB vs 126 MB)"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------ our arm
def cpu_baseline_leg(wl, seconds_budget=20.0):
    """bounded sample of the same workload on the host: reference binary if it travelled
    (kind 'reference', 1 core: serial PSTL fallback), else the OpenMP oracle port."""
    binp = os.path.join(ROOT, "oracle", "_ref", "ref_bench_2d")
    try:
        if os.path.exists(binp):
            n = 30
            with tempfile.TemporaryDirectory() as td:
                sp = os.path.join(td, "s.txt")
                open(sp, "w").write(wl.c2_script() + "\nI\nX\nT 2\nT %d\n" % n)
                out = subprocess.run([binp, sp, os.path.join(td, "o.bin"), "4096"], check=True,
                                     capture_output=True, text=True, timeout=300).stdout
            rec = [json.loads(l) for l in out.splitlines() if l.startswith("{")][-1]
            out = {"value": rec["updates_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": "%d amr_solver::advance() steps of the full C2 mesh (%d cells), unmodified "
                             "reference headers, Release flags, EXECUTION=PAR (serial PSTL: no TBB)"
                             % (n, rec["cells"]), "seconds": rec["seconds"]}
            try:
                # context only: the OpenMP C restatement (oracle/) on all host cores, same mesh
                out["port_openmp_all_cores"] = oracle_port_leg(wl, 8.0)
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("oracle port leg failed: %r\n" % (e,))
            return out
    except Exception as e:  # fall through to the port
        sys.stderr.write("reference baseline failed: %r\n" % (e,))
    return oracle_port_leg(wl, seconds_budget / 2)


def oracle_port_leg(wl, seconds):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cfg = O.Config.from_name("r2_s64_h1_d7_euler")
    tree = O.OracleTree(cfg, capacity=4096)
    O.run_script(tree, wl.c2_script() + "\nI\nX")
    tree.advance_batch(2)
    n, t0 = 0, time.time()
    while time.time() - t0 < seconds and n < 400:
        tree.advance_batch(4)
        n += 4
    secs = time.time() - t0
    cells = tree.size * cfg.size ** cfg.rank
    return {"value": cells * n / secs, "unit": UNIT, "cores": O.lib().orc_num_threads(), "kind": "port",
            "sample": "%d steps of the full C2 mesh (%d cells), OpenMP C oracle" % (n, cells),
            "seconds": secs}


def run_ours(args):
    import numpy as np
    import torch

    amrb = importlib.import_module("gpu-amr_b200")
    wl = importlib.import_module("gpu-amr_b200.workloads")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        mg = importlib.import_module("gpu-amr_b200.multigpu")
        return mg.run_bench(args, METRIC, UNIT)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    peaks, peak_src = measured_peaks()

    if args.workload == "c2":
        cfg = wl.c2_config()
        host = wl.build_static_tree(cfg, wl.C2["base_level"], wl.C2["ball_radii"])
    else:
        # development workloads (not the driver's bench line): name = r<rank>_s<size>_h<halo>_<eq>_L<base>[m]
        # e.g. r3_s8_h1_euler_L5m = 3D Euler 8^3 patches, uniform level 5 + two refinement rings
        t = args.workload.split("_")
        rank, size, halo = int(t[0][1:]), int(t[1][1:]), int(t[2][1:])
        eq = amrb.EQ_EULER if t[3] == "euler" else amrb.EQ_ADVECTION
        base, multi = int(t[4][1:].rstrip("m")), t[4].endswith("m")
        cfg = wl.Config(rank, size, halo, 7 if rank == 2 else 7, eq)
        host = wl.build_static_tree(cfg, base, (0.25, 0.125) if multi else ())
    ids = host.ids()
    P = len(ids)
    cells = P * cfg.data
    lay = amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth)
    pool = amrb.DevicePool(lay, P, local)
    pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
    pool.set_topology(*host.tables())
    if os.environ.get("AMRB_MODE"):
        pool.set_mode(int(os.environ["AMRB_MODE"]))
    ic = wl.initial_condition(ids, cfg)                       # [nvar, P, S, S] on the host
    L = amrb.lib()
    stream = torch.cuda.ExternalStream(int(L.amrb_pool_stream(pool.h) or 0), device=local)

    # pinned host staging of the whole padded state (what ndtree::sync_current_to_device moves)
    pinned = [torch.zeros(P * pool.flat, dtype=torch.float64).pin_memory() for _ in range(cfg.nvar)]
    for f in range(cfg.nvar):
        v = pinned[f].numpy().reshape((P,) + (cfg.psize,) * cfg.rank)
        v[(slice(None),) + (slice(cfg.halo, cfg.halo + cfg.size),) * cfg.rank] = ic[f]

    def upload():
        for f in range(cfg.nvar):
            amrb.check(L.amrb_copy_host_to_device_async(L.amrb_pool_field(pool.h, f),
                                                        pinned[f].data_ptr(), P * pool.flat * 8,
                                                        L.amrb_pool_stream(pool.h)))

    def download():
        for f in range(cfg.nvar):
            amrb.check(L.amrb_copy_device_to_host_async(pinned[f].data_ptr(),
                                                        L.amrb_pool_field(pool.h, f), P * pool.flat * 8,
                                                        L.amrb_pool_stream(pool.h)))

    upload()
    pool.halo_exchange()
    pool.synchronize()

    K, W = args.steps, args.warmup
    # ---- warm-up (also leaves the carried dt-min so the timed batch starts without a dt pass)
    pool.advance_batch_async(max(W, 3))
    pool.finish_advance_batch()
    pool.advance_batch_async(K)          # one untimed batch of the timed shape (first-use effects)
    pool.finish_advance_batch()

    # ---- (1) device-resident throughput: EXACTLY K steps in one batch, CUDA events on the pool
    # stream; repeated REPS times back to back (each repetition is K steps), median reported
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    REPS = 5
    reps_ms = []
    launches0 = pool.launch_count()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    for _ in range(REPS):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        pool.advance_batch_async(K)
        ev1.record(stream)
        torch.cuda.synchronize()
        dt_sum, executed, _ = pool.finish_advance_batch()
        reps_ms.append(ev0.elapsed_time(ev1))
    t_wall1 = time.time()
    launches = (pool.launch_count() - launches0) // REPS
    ms_total = sorted(reps_ms)[REPS // 2]
    value = cells * K / (ms_total * 1e-3)

    # ---- (2) per-launch duration of the dominant kernel (fused step), events around each launch
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    amrb.check(L.amrb_pool_batch_begin(pool.h, K, amrb.DBL_MAX))
    torch.cuda.synchronize()
    evs[0].record(stream)
    for k in range(K):
        amrb.check(L.amrb_pool_step_partial(pool.h, None, 0))
        amrb.check(L.amrb_pool_step_commit(pool.h))
        evs[k + 1].record(stream)
    amrb.check(L.amrb_pool_batch_end(pool.h, 1))
    torch.cuda.synchronize()
    pool.finish_advance_batch()
    per = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(K))
    kern_ms_isolated = sum(per) / len(per)       # an event pair around every launch (adds gaps)
    # average launch duration over the timed region: the K-step batch is K back-to-back launches of
    # the fused step kernel (+ one init_scalars and one halo launch, < 0.3 % of the region)
    kern_ms = ms_total / K
    clk = clocks.stop(t_wall0, time.time())

    b_alg = 2 * cfg.nvar * 8                                  # read state once + write once, fp64
    achieved = cells * b_alg / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic() if args.workload == "c2" else None,
                "kernel": ("fused step kernel euler2d_march_kernel<64,1,band> (C2)" if args.workload == "c2"
                           else "fused step kernel of " + args.workload), "kernel_ms": kern_ms, "kernel_ms_event_pair_per_launch": kern_ms_isolated,
                "kernel_ms_median": per[len(per) // 2],
                "algorithmic_bytes_per_cell": b_alg, "peak_source": peak_src}

    # ---- (3) end to end through the C ABI with HOST buffers: sync_current_to_device (pinned H2D of
    # the padded state) -> advance_batch(K) -> finish (scalar read-back) -> sync_current_from_device
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    upload()
    pool.halo_exchange()
    pool.advance_batch_async(K)
    download()
    e1.record(stream)
    torch.cuda.synchronize()
    pool.finish_advance_batch()
    e2e_ms = e0.elapsed_time(e1)
    state_bytes = cfg.nvar * P * pool.flat * 8
    e2e = {"value": cells * K / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": state_bytes / K, "d2h_bytes_per_step": state_bytes / K + 8,
           "ms_total": e2e_ms,
           "what": "pinned-host H2D of the padded state + halo fill + %d steps + D2H of the state, "
                   "one job (the reference benchmark copies state once per run, b.cpp:194-197)" % K}

    cpu = cpu_baseline_leg(wl) if (not args.no_cpu_baseline and args.workload == "c2") else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(cells, {"patches": P, "executed_steps": int(executed), "sum_dt": dt_sum},
                                  None if args.workload == "c2" else args.workload),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk, "wall_s_timed": t_wall1 - t_wall0, "repetitions_ms": reps_ms,
    }
    print(json.dumps(line))
    pool.close()


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture"""
    p = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", help="c2 (bench line) or a development workload name")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 20 if args.steps is None else args.steps
        args.warmup = 2 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 200 if args.steps is None else args.steps
    args.warmup = 10 if args.warmup is None else args.warmup
    run_ours(args)


if __name__ == "__main__":
    main()
