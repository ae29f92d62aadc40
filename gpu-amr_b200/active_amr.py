"""Active-AMR driving loops over the C ABI (BASELINE configs C4 / C5): `interval` fused steps, the refinement
criterion on the device, one host reconstruct (refine + coarsen + 2:1 ripple), the refine / coarsen data
motion fused with the Morton re-sort on the device, new halo tables — and, on a sharded mesh, the
re-slicing of the Morton ranges.  The same loop as the reference's benchmark drivers
(benchmark/bench_fvm_solver_integration_active_amr.b.cpp:218-247) and its advection example
(examples/fvm_solver_advection.e.cpp: reconstruct every 5 steps, refine > 0.1, coarsen < 0.05 on the
patch maximum).

C5 = 3D advection of a Gaussian pulse with >= 4 refinement levels: 8^3 patches, interior-only pools,
levels `min_level` .. `max_level`, pulse at 0.2 L, width 0.005 L^2.
"""
import numpy as np

from . import binding as B
from . import workloads as wl

C5 = dict(rank=3, size=8, halo=1, depth=8, min_level=4, max_level=8, refine=0.1, coarsen=0.05, interval=5,
          capacity=400000)


def c5_config():
    return wl.Config(C5["rank"], C5["size"], C5["halo"], C5["depth"], B.EQ_ADVECTION)


class ActiveAmr:
    """single-GPU loop (DeviceTree) or sharded loop (ShardedSolver + global host tree), same interface"""

    def __init__(self, cfg, torch, device=0, dist=None, rank=0, world=1, p=C5, storage=B.STORAGE_INTERIOR,
                 transport="p2p", host_ic=False, device_regrid=True):
        self.cfg, self.torch, self.device, self.p = cfg, torch, device, p
        self.host_ic = host_ic            # evaluate the pulse with numpy on the host (bit-comparable with the oracle)
        self.dist, self.rank, self.world = dist, rank, world
        self.storage, self.transport = storage, transport
        self.sharded = world > 1
        # single GPU: the whole reconstruct (selection included) runs on the device, the host tree only mirrors
        # the leaf ids; sharded meshes select on the host (every rank the same global flags)
        self.device_regrid = device_regrid and not self.sharded
        self.updates = 0
        self.regrids = self.changed = 0
        host = B.HostTree(cfg.rank, cfg.depth)
        for _ in range(p["min_level"]):
            host.reconstruct(wl.flags_refine_all(host.ids()))
        self.host = host
        self._build()
        # initial adaptation: evaluate the pulse, flag, refine -- until the mesh stops changing
        for _ in range(p["max_level"] - p["min_level"] + 1):
            self.fill_ic()
            if not self.regrid():
                break
        self.fill_ic()

    # ---- construction
    def _build(self):
        cfg, p = self.cfg, self.p
        if self.sharded:
            from . import multigpu as mg
            self.sol = mg.ShardedSolver(cfg, self.host, self.rank, self.world, self.device, self.dist, self.torch,
                                        capacity=p["capacity"], storage=self.storage, transport=self.transport)
            self.pool = self.sol.pool
        else:
            lay = B.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, self.storage)
            self.pool = B.DevicePool(lay, p["capacity"], self.device)
            self.pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
            self.pool.set_topology_from_ids(self.host.ids())

    def my_ids(self):
        return self.sol.ids if self.sharded else self.host.ids()

    def fill_ic(self):
        import importlib
        mg = importlib.import_module("gpu-amr_b200.multigpu")
        ids, cfg, torch = self.my_ids(), self.cfg, self.torch
        if self.host_ic:
            self.pool.upload_interior(0, wl.initial_condition(ids, cfg)[0])
            self.halo_exchange()
            return
        stored = self.pool.stored
        view = mg.raw_tensor(self.pool.L.amrb_pool_field(self.pool.h, 0), len(ids) * stored, torch)
        for s, fields in wl.device_initial_condition(torch, ids, cfg, torch.device("cuda", self.device)):
            n = fields[0].shape[0]
            dst = view[s * stored:(s + n) * stored]
            if stored == cfg.data:
                dst.copy_(fields[0].reshape(-1))
            else:
                h, S, R = cfg.halo, cfg.size, cfg.rank
                dst.view((n,) + (cfg.psize,) * R)[(slice(None),) + (slice(h, h + S),) * R] = fields[0]
        torch.cuda.synchronize()
        self.pool.mark_dirty()
        self.halo_exchange()

    def halo_exchange(self):
        if self.sharded:
            self.sol.halo_exchange()
        else:
            self.pool.halo_exchange()

    # ---- one reconstruct: criterion on the device, selection on the host, data motion on the device
    def flags(self):
        p = self.p
        if self.sharded:
            mine = self.sol.pool.patch_max_flags(0, p["refine"], p["coarsen"], p["min_level"], p["max_level"])
            mine = np.ascontiguousarray(mine[:self.sol.plan.n_owned])
            parts = [None] * self.world
            self.dist.all_gather_object(parts, mine)
            return np.concatenate(parts)
        return self.pool.patch_max_flags(0, p["refine"], p["coarsen"], p["min_level"], p["max_level"])

    def regrid(self):
        if self.device_regrid:
            p = self.p
            self.pool.flag_patches(0, p["refine"], p["coarsen"], p["min_level"], p["max_level"])
            changed, _ = self.pool.reconstruct_device(None)
            self.regrids += 1
            if changed:
                self.changed += 1
                self.host.assign(self.pool.get_ids())
                self.pool.halo_exchange()
            return changed
        flags = self.flags()
        old_size = self.host.size
        changed = self.host.reconstruct(flags, self.p["capacity"])
        self.regrids += 1
        if not changed:
            return 0
        self.changed += 1
        if self.sharded:
            self.sol.reshard(self.host, old_size, self.host.plan())
            self.sol.halo_exchange()
        else:
            self.pool.apply_plan(*self.host.plan())
            self.pool.set_topology_from_ids(self.host.ids())
            self.pool.halo_exchange()
        return 1

    # ---- the driver loop: `cycles` x (interval steps, reconstruct)
    def run(self, cycles):
        steps = self.p["interval"]
        dts = []
        for _ in range(cycles):
            patches = self.host.size
            if self.sharded:
                self.sol.advance_batch_async(steps)
                _, n, d = self.sol.finish_advance_batch(steps)
            else:
                self.pool.advance_batch_async(steps)
                _, n, d = self.pool.finish_advance_batch(steps)
            self.updates += n * patches * self.cfg.data
            dts += list(d)
            self.regrid()
        return np.asarray(dts)

    def close(self):
        if self.sharded:
            self.sol.close()
        else:
            self.pool.close()


def bench_c5(args, torch, metric, unit, clock_sampler, measured_peaks, dist=None):
    """`bench.py --workload c5`: K solver steps of the C5 loop (reconstruct every 5 steps), everything between
    the first and the last step inside the timed region (steps, criterion, host reconstruct, data motion, new
    tables, on N > 1 the re-slicing); wall clock between device synchronisations, max over ranks."""
    import json
    import os
    import time

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = c5_config()
    run = ActiveAmr(cfg, torch, device=local, dist=dist, rank=rank, world=world,
                    transport=os.environ.get("AMRB_TRANSPORT", "p2p"))
    p0 = run.host.size
    levels = sorted(set((run.host.ids() & np.uint64(63)).astype(int).tolist()))
    K = max(args.steps, C5["interval"])
    cycles = K // C5["interval"]
    run.run(max(1, args.warmup // C5["interval"]))
    clocks = clock_sampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    run.updates, r0, c0 = 0, run.regrids, run.changed
    t0 = time.time()
    run.run(cycles)
    torch.cuda.synchronize()
    secs = time.time() - t0
    if dist is not None:
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    t1 = time.time()
    if rank == 0:
        clk = clocks.stop(t0, t1)
        peaks, src = measured_peaks()
        value = run.updates / secs
        achieved = value * 16 / 1e9 / world
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": cycles * C5["interval"],
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / (cycles * C5["interval"]), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5: 3D active-AMR advection of a Gaussian pulse, 8^3 patches halo 1, levels "
                                   "%d-%d, criterion refine > %g / coarsen < %g on the patch maximum, reconstruct "
                                   "every %d steps (everything inside the timed region)"
                                   % (C5["min_level"], C5["max_level"], C5["refine"], C5["coarsen"], C5["interval"]),
                       "patches_start": int(p0), "patches_end": int(run.host.size), "levels": levels,
                       "cell_updates": int(run.updates), "reconstructs": run.regrids - r0,
                       "topology_changing_reconstructs": run.changed - c0,
                       "device_layout": "interior-only [P][S^3]", "l2_policy": "state smaller than L2 on this mesh"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                         "kernel": "whole AMR loop (advect3d_dense_kernel + criterion + plan + topology kernels + host reconstruct)",
                         "algorithmic_bytes_per_cell": 16, "peak_source": src},
            "cpu_baseline": None, "e2e": None, "gpu_launches": None, "clocks": clk,
        }
        print(json.dumps(line))
    run.close()
