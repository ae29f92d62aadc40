#!/usr/bin/env python
"""bench.py — cell-updates/sec of the fused halo + flux + update step on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|<dev>]
  torchrun ... bench.py --gpus N ...          (one rank per GPU, N > 1)

A "step" is one explicit time step of the finite-volume solver over the whole mesh (ghost-cell gather,
Rusanov face fluxes, conservative update, CFL reduction for the next step).

Workload (every N, default): BASELINE.json configs[2] "C3" — the 3D static multi-level tree of 8^3 Euler
patches, ~1.07e9 cells (levels 6-8, ~2.1e6 patches; gpu-amr_b200/workloads.py: C3), the SAME global mesh at
N = 1, 2, 4, 8 (strong scaling, Morton-range partition).  On one GPU it fits because rank-3 pools keep
interiors only (86 GB).  At N = 1 the line also carries "c2": the 2D configs[1] mesh (2 272 patches of 64x64
Euler cells, the round-1 bench line) measured in the same run.  --workload c2 makes C2 the line itself
(weak-scaled C2 family at N > 1).  Prints ONE JSON line.
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
C3_NAME = ("C3: bench_fvm_solver_integration3D patch shape (8^3 Euler fp64, halo 1) on a 3D static multi-level "
           "tree, levels %d-%d, acoustic pulse")
C2_NAME = ("C2: bench_fvm_solver_integration 2D static multi-level tree, Euler fp64, 64x64 patches halo 1, "
           "levels 5-7, acoustic pulse")


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
        sm, mx, power, reasons = [], None, [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ baselines
def _run_ref_binary(name, script, capacity, timeout=900):
    """run a prebuilt reference binary (oracle/_ref/*, built by oracle/Makefile from the unmodified reference
    sources) on a script; returns the JSON records of its T ops"""
    binp = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(binp):
        return None
    with tempfile.TemporaryDirectory() as td:
        sp = os.path.join(td, "s.txt")
        open(sp, "w").write(script + "\n")
        out = subprocess.run([binp, sp, os.path.join(td, "o.bin"), str(capacity)], check=True,
                             capture_output=True, text=True, timeout=timeout).stdout
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _workload_scripts(wl, workload, base):
    if workload == "c2":
        return wl.c2_script(base), "ref_bench_2d", "ref_cuda_bench_2d", "r2_s64_h1_d7_euler"
    return wl.c3_script(base), "ref_bench_3d", "ref_cuda_bench_3d", "r3_s8_h1_d8_euler"


def _ref_capacity(workload, base):
    """patch slots of the reference tree (it allocates capacity x padded patch x fields x 2 up front, and
    appends children before it compacts: headroom 2-3x the final leaf count)"""
    if workload == "c2":
        return 4096
    return {4: 100000, 3: 16000}.get(base, 100000 * 8 ** max(base - 4, 0))


def cpu_reference_leg(wl, workload, base, steps, warmup=1):
    """the reference's own CPU implementation (unmodified headers, Release flags, EXECUTION=PAR) on a
    bounded sample; libstdc++'s parallel policies run serially without TBB -> 1 core"""
    script, cpu_bin, _, cfgname = _workload_scripts(wl, workload, base)
    cap = _ref_capacity(workload, base)
    lines = [script, "I", "X"] + (["T %d" % warmup] if warmup else []) + ["T %d" % steps]
    recs = _run_ref_binary(cpu_bin, "\n".join(lines), cap)
    if recs:
        rec = recs[-1]
        return {"value": rec["updates_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
                "seconds": rec["seconds"], "cells": rec["cells"], "patches": rec["patches"],
                "steps": rec["steps"]}
    # the oracle port (OpenMP) — only when the reference binary did not travel
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cfg = O.Config.from_name(cfgname)
    tree = O.OracleTree(cfg, capacity=cap)
    O.run_script(tree, script + "\nI\nX")
    tree.advance_batch(warmup)
    t0 = time.time()
    tree.advance_batch(steps)
    secs = time.time() - t0
    cells = tree.size * cfg.size ** cfg.rank
    return {"value": cells * steps / secs, "unit": UNIT, "cores": O.lib().orc_num_threads(), "kind": "port",
            "seconds": secs, "cells": cells, "patches": tree.size, "steps": steps}


def reference_cuda_leg(wl, workload, base, steps):
    """the reference's OWN CUDA backend (src/cuda/*.cu built for sm_100, oracle/Makefile) on the same box:
    the GPU-vs-GPU baseline of SURVEY 8d / BASELINE.md 4.3.  One batch of `steps` steps, wall clock."""
    script, _, cuda_bin, _ = _workload_scripts(wl, workload, base)
    try:
        recs = _run_ref_binary(cuda_bin, "\n".join([script, "I", "X", "T 3", "T %d" % steps]),
                               _ref_capacity(workload, base))
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}
    if not recs:
        return None
    rec = recs[-1]
    return {"value": rec["updates_per_s"], "unit": UNIT, "ms_per_step": 1e3 * rec["seconds"] / max(rec["steps"], 1),
            "cells": rec["cells"], "patches": rec["patches"], "steps": rec["steps"],
            "what": "reference src/cuda kernels (compute_dt + finalize + time_step + per-field halo launches), "
                    "sm_100 build, one advance_batch_async(%d) timed by wall clock" % steps}


def sample_base(workload, n_steps):
    """base level of the bounded CPU sample: geometrically similar mesh, fewer cells (per-core CPU throughput
    is size-independent once the state is out of cache)"""
    if workload == "c2":
        return 5 if n_steps <= 300 else (4 if n_steps <= 1200 else 3)   # 9.3e6 / 2.3e6 / 5.8e5 cells
    return 4 if n_steps <= 40 else 3                                     # 1.7e7 / 2.1e6 cells


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on this box's host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = importlib.import_module("gpu-amr_b200.workloads")
    workload = "c2" if args.workload == "c2" else "c3"
    base = sample_base(workload, args.steps + args.warmup)
    r = cpu_reference_leg(wl, workload, base, args.steps, args.warmup)
    full = workload_config(wl, workload, args.gpus)
    sample = ("1 amr_solver::advance() per bench step over the geometrically similar %s mesh at base level %d "
              "(%d cells, %d patches); unmodified reference headers, Release flags, EXECUTION=PAR = serial PSTL "
              "(no TBB on the box) -> %d core" % (workload.upper(), base, r["cells"], r["patches"], r["cores"]))
    cfg = dict(full)
    cfg["sample_cells"] = int(r["cells"])
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
        "higher_is_better": True, "scaling": full_scaling(workload), "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def full_scaling(workload):
    return "weak" if workload == "c2" else "strong"


def workload_config(wl, workload, n_gpus, cells=None):
    """the `config` object both arms print (same keys, same values)"""
    if workload == "c2":
        return {"workload": C2_NAME, "cells": int(cells) if cells else 9306112,
                "l2_policy": "inputs larger than L2 (state 2 x 0.3 GB vs 126 MB)"}
    lv0 = wl.C3["base_level"]
    c = {"workload": C3_NAME % (lv0, lv0 + 2), "base_level": lv0, "ball_radii": list(wl.C3["ball_radii"]),
         "cells": int(wl.C3["cells"]), "l2_policy": "inputs larger than L2 (state 2 x 42 GB vs 126 MB)"}
    if cells:
        assert int(cells) == c["cells"], "the C3 mesh changed: update workloads.C3"
    return c


# ------------------------------------------------------------------------------------ our arm
def timed_batches(torch, pool, stream, K, min_seconds=0.5, min_reps=3, max_reps=400):
    """EXACTLY K steps per batch, CUDA events on the pool's stream; the batch is repeated back to back until
    the timed region is >= min_seconds; returns the per-batch times [ms]"""
    reps, total = [], 0.0
    while len(reps) < min_reps or (total < min_seconds * 1e3 and len(reps) < max_reps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        pool.advance_batch_async(K)
        ev1.record(stream)
        torch.cuda.synchronize()
        out = pool.finish_advance_batch()
        reps.append(ev0.elapsed_time(ev1))
        total += reps[-1]
    return reps, out


def per_launch_events(torch, amrb, pool, stream, K):
    """an event pair around every launch of the fused step kernel (adds inter-launch gaps)"""
    L = amrb.lib()
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
    return sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(K))


def spread(xs):
    s = sorted(xs)
    return {"n": len(s), "min": s[0], "median": s[len(s) // 2], "max": s[-1]}


def load_traffic(key):
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture of the same
    workload (profiles/step_kernel_traffic.json: not measured inside this run)"""
    p = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        rec = json.load(open(p)).get(key)
        return (rec.get("dram_bytes_per_launch"), rec.get("source"), rec.get("patches")) if rec else (None, None, None)
    except Exception:
        return None, None, None


def pool_field_views(torch, amrb, pool, n_doubles, nvar, which="cur"):
    mg = importlib.import_module("gpu-amr_b200.multigpu")
    L = amrb.lib()
    get = L.amrb_pool_field if which == "cur" else L.amrb_pool_next_field
    return [mg.raw_tensor(get(pool.h, f), n_doubles, torch) for f in range(nvar)]


def build_pool(torch, amrb, wl, cfg, host, storage, device, args):
    ids = host.ids()
    P = len(ids)
    lay = amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage)
    pool = amrb.DevicePool(lay, P, device)
    pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
    pool.set_topology_from_ids(ids)
    if args.variant is not None:
        pool.set_variant(args.variant)
    if os.environ.get("AMRB_MODE"):
        pool.set_mode(int(os.environ["AMRB_MODE"]))
    return pool, ids, P


def fill_ic(torch, amrb, wl, pool, ids, cfg, device):
    """initial condition evaluated on the device, written into the pool's current buffers"""
    stored, S, R = pool.stored, cfg.size, cfg.rank
    views = pool_field_views(torch, amrb, pool, len(ids) * stored, cfg.nvar)
    dev = torch.device("cuda", device)
    h = cfg.halo
    for s, fields in wl.device_initial_condition(torch, ids, cfg, dev):
        n = fields[0].shape[0]
        for f, t in enumerate(fields):
            dst = views[f][s * stored:(s + n) * stored]
            if stored == cfg.data:
                dst.copy_(t.reshape(-1))
            else:   # padded pool: interior of the padded patches
                v = dst.view((n,) + (cfg.psize,) * R)
                v[(slice(None),) + (slice(h, h + S),) * R] = t
    torch.cuda.synchronize()
    pool.mark_dirty()


def host_state_buffers(torch, nbytes_per_field, nvar):
    """pinned host image of the state, one buffer per field; falls back to ONE pinned window shared by all
    fields when the box's free memory does not allow the full image (the bytes moved are the same)"""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    n = nbytes_per_field // 8
    if avail > 2.5 * nbytes_per_field * nvar:
        try:
            return [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(nvar)], "pinned, full state image"
        except Exception:
            pass
    w = torch.empty(n, dtype=torch.float64).pin_memory()
    return [w] * nvar, "pinned, one field-sized window shared by all fields (host memory bound)"


def measure_single(torch, amrb, wl, args, workload, device, light=False):
    """device-resident throughput, per-launch timing and the end-to-end leg of one workload on one GPU"""
    L = amrb.lib()
    peaks, peak_src = measured_peaks()
    if workload == "c2":
        cfg = wl.c2_config()
        host = wl.build_static_tree(cfg, wl.C2["base_level"], wl.C2["ball_radii"])
        storage, name, kernel, tkey = amrb.STORAGE_PADDED, C2_NAME, "euler2d_march_kernel<64,1,band> (C2)", "c2"
    elif workload == "c3":
        cfg = wl.c3_config()
        host = wl.build_static_tree(cfg, wl.C3["base_level"], wl.C3["ball_radii"])
        storage, kernel, tkey = amrb.STORAGE_INTERIOR, "euler3d_dense_kernel<8,...> (C3, interior-only layout)", "c3"
        name = C3_NAME % (wl.C3["base_level"], wl.C3["base_level"] + 2)
    else:
        # development workloads: r<rank>_s<size>_h<halo>_<eq>_L<base>[m][_d<depth>]  e.g. r3_s8_h1_euler_L5m
        t = workload.split("_")
        rank, size, halo = int(t[0][1:]), int(t[1][1:]), int(t[2][1:])
        eq = amrb.EQ_EULER if t[3] == "euler" else amrb.EQ_ADVECTION
        base, multi = int(t[4][1:].rstrip("m")), t[4].endswith("m")
        depth = int(t[5][1:]) if len(t) > 5 else 7
        cfg = wl.Config(rank, size, halo, depth, eq)
        host = wl.build_static_tree(cfg, base, (0.25, 0.125) if multi else ())
        storage = args.storage
        name, kernel, tkey = "development workload " + workload, "fused step kernel of " + workload, workload
    pool, ids, P = build_pool(torch, amrb, wl, cfg, host, storage, device, args)
    cells = P * cfg.data
    stream = torch.cuda.ExternalStream(int(L.amrb_pool_stream(pool.h) or 0), device=device)
    fill_ic(torch, amrb, wl, pool, ids, cfg, device)
    pool.halo_exchange()
    pool.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    pool.advance_batch_async(W)
    pool.finish_advance_batch()
    pool.advance_batch_async(K)          # one untimed batch of the timed shape (first-use effects)
    pool.finish_advance_batch()

    clocks = ClockSampler(device)
    clocks.start()
    time.sleep(0.25)
    launches0 = pool.launch_count()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    reps_ms, (dt_sum, executed, _) = timed_batches(torch, pool, stream, K, 0.25 if light else 0.5)
    t_wall1 = time.time()
    launches = (pool.launch_count() - launches0) // len(reps_ms)
    ms_total = sorted(reps_ms)[len(reps_ms) // 2]
    value = cells * K / (ms_total * 1e-3)
    per = per_launch_events(torch, amrb, pool, stream, K)
    clk = clocks.stop(t_wall0, time.time())

    b_alg = 2 * cfg.nvar * 8                                  # read state once + write once, fp64
    kern_ms = ms_total / K   # K back-to-back launches of the fused step (+ init_scalars [+ halo]: < 0.3 %)
    achieved = cells * b_alg / (kern_ms * 1e-3) / 1e9
    traffic, tsrc, tpatches = load_traffic(tkey)
    if traffic and tpatches and tpatches != P:
        # the capture is of the same kernel on a mesh of another size (development runs): scale by the patch count
        traffic = traffic * P / float(tpatches)
        tsrc = (tsrc or "") + "; scaled by the patch count (%d / %d)" % (P, tpatches)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": tsrc,
                "kernel": kernel, "kernel_ms": kern_ms,
                "kernel_ms_event_pair_per_launch": sum(per) / len(per), "kernel_ms_median": per[len(per) // 2],
                "algorithmic_bytes_per_cell": b_alg, "peak_source": peak_src}
    out = {"value": value, "ms_per_step": kern_ms, "cells": cells, "patches": P, "executed_steps": int(executed),
           "sum_dt": dt_sum, "roofline": roofline, "launches": int(launches), "clocks": clk,
           "batch_ms": spread(reps_ms), "timed_region_s": sum(reps_ms) * 1e-3, "name": name,
           "wall_s_timed": t_wall1 - t_wall0, "cfg": cfg, "storage": storage}
    if light:
        pool.close()
        return out

    # ---- end to end through the C ABI with HOST buffers: pinned H2D of the state (sync_current_to_device),
    # K steps, scalar read-back, D2H of the state (sync_current_from_device); measured for K and for 10 K
    nvar, stored = cfg.nvar, pool.stored
    fbytes = P * stored * 8
    bufs, how = host_state_buffers(torch, fbytes, nvar)
    st = L.amrb_pool_stream(pool.h)
    for f in range(nvar):
        amrb.check(L.amrb_copy_device_to_host(bufs[f].data_ptr(), L.amrb_pool_field(pool.h, f), fbytes))

    def e2e_run(k):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for f in range(nvar):
            amrb.check(L.amrb_copy_host_to_device_async(L.amrb_pool_field(pool.h, f), bufs[f].data_ptr(), fbytes, st))
        pool.mark_dirty()                       # raw writes through amrb_pool_field: drop the carried dt-min
        pool.halo_exchange()
        pool.advance_batch_async(k)
        for f in range(nvar):
            amrb.check(L.amrb_copy_device_to_host_async(bufs[f].data_ptr(), L.amrb_pool_field(pool.h, f), fbytes, st))
        e1.record(stream)
        torch.cuda.synchronize()
        pool.finish_advance_batch()
        return e0.elapsed_time(e1)

    ms_k = e2e_run(K)
    k10 = 10 * K
    ms_k10 = e2e_run(k10) if (kern_ms * k10 < 20e3) else None
    state_bytes = nvar * fbytes
    out["e2e"] = {"value": cells * K / (ms_k * 1e-3), "unit": UNIT,
                  "h2d_bytes_per_step": state_bytes / K, "d2h_bytes_per_step": state_bytes / K + 8,
                  "ms_total": ms_k, "steps": K, "host_buffers": how,
                  "what": "pinned-host H2D of the state + %d steps + D2H of the state, one job (the reference "
                          "benchmark copies the state once per run, b.cpp:194-197); the copies are amortised over "
                          "the job's steps, so the figure depends on --steps: see e2e_10x" % K,
                  "e2e_10x": None if ms_k10 is None else
                  {"steps": k10, "value": cells * k10 / (ms_k10 * 1e-3), "ms_total": ms_k10}}
    pool.close()
    return out


def run_ours(args):
    import torch

    amrb = importlib.import_module("gpu-amr_b200")
    wl = importlib.import_module("gpu-amr_b200.workloads")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "c5":
        # BASELINE configs[4]: 3D active-AMR advection (development line, not the driver's bench line)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(local)
        aa = importlib.import_module("gpu-amr_b200.active_amr")
        dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        aa.bench_c5(args, torch, METRIC, UNIT, ClockSampler, measured_peaks, dist)
        if dist is not None:
            dist.destroy_process_group()
        return
    if world > 1:
        mg = importlib.import_module("gpu-amr_b200.multigpu")
        return mg.run_bench(args, METRIC, UNIT)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    workload = args.workload
    m = measure_single(torch, amrb, wl, args, workload, local)
    torch.cuda.empty_cache()
    K = args.steps
    named = workload in ("c2", "c3")
    cfgobj = workload_config(wl, workload, 1, m["cells"]) if named else \
        {"workload": m["name"], "cells": int(m["cells"]), "l2_policy": "inputs larger than L2"}
    cfgobj.update({"patches": m["patches"], "executed_steps": m["executed_steps"], "sum_dt": m["sum_dt"],
                   "device_layout": "interior-only [P][S^3] per field" if m["storage"] else "padded (reference device layout)"})
    cpu = None
    if named and not args.no_cpu_baseline:
        base = sample_base(workload, 8)
        try:
            r = cpu_reference_leg(wl, workload, base, 3 if workload == "c3" else 30)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "seconds": r["seconds"],
                   "sample": "%d amr_solver::advance() steps of the geometrically similar %s mesh at base level %d "
                             "(%d cells), unmodified reference headers, Release flags, EXECUTION=PAR (serial PSTL: "
                             "no TBB)" % (r["steps"], workload.upper(), base, r["cells"])}
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("cpu baseline failed: %r\n" % (e,))
        try:
            # GPU-vs-GPU: the reference's own CUDA backend on this GPU (C3: the base-level-4 sample, its padded
            # pool and host-side tree do not reach 2e6 patches; C2: the full mesh)
            cb = 4 if workload == "c3" else 5
            rc = reference_cuda_leg(wl, workload, cb, 20)
            if cpu is not None:
                cpu["reference_cuda"] = rc
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference CUDA leg failed: %r\n" % (e,))
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": full_scaling(workload) if named else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfgobj,
        "roofline": m["roofline"], "cpu_baseline": cpu, "e2e": m.get("e2e"), "gpu_launches": m["launches"],
        "clocks": m["clocks"], "batch_ms": m["batch_ms"], "timed_region_s": m["timed_region_s"],
    }
    if workload == "c3" and not args.no_secondary:
        # the 2D configs[1] mesh in the same run (device-resident throughput + roofline only)
        try:
            c2 = measure_single(torch, amrb, wl, args, "c2", local, light=True)
            sec = {"workload": C2_NAME, "value": c2["value"], "unit": UNIT, "ms_per_step": c2["ms_per_step"],
                   "cells": c2["cells"], "patches": c2["patches"], "roofline": c2["roofline"],
                   "batch_ms": c2["batch_ms"], "timed_region_s": c2["timed_region_s"], "clocks": c2["clocks"]}
            if not args.no_cpu_baseline:
                sec["reference_cuda"] = reference_cuda_leg(wl, "c2", 5, 20)
            line["c2"] = sec
        except Exception as e:  # noqa: BLE001
            line["c2"] = {"error": repr(e)[:300]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 measurement attached to the C3 line")
    ap.add_argument("--workload", default="c3", help="c3 (bench line), c2, or a development workload name")
    ap.add_argument("--storage", type=int, default=0, help="development workloads: 0 padded, 1 interior-only")
    ap.add_argument("--variant", type=int, default=None, help="kernel variant (amrb_pool_set_variant)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 20 if args.steps is None else args.steps
        args.warmup = 2 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 20 if args.steps is None else args.steps
    args.warmup = 5 if args.warmup is None else args.warmup
    run_ours(args)


if __name__ == "__main__":
    main()
