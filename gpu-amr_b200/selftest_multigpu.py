"""torchrun entry: sharded run of a small multi-level mesh on N GPUs compared, on rank 0, with the
single-GPU pool on the same mesh (same kernels, so the result must agree to rounding of nothing:
bit-exact) and the step-size sequence.  Exit code 0 = parity."""
import importlib
import os
import sys

import numpy as np


def main():
    import torch
    import torch.distributed as dist

    amrb = importlib.import_module("gpu-amr_b200")
    mg = importlib.import_module("gpu-amr_b200.multigpu")
    wl = importlib.import_module("gpu-amr_b200.workloads")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = sys.argv[1] if len(sys.argv) > 1 else "2d"
    if name == "2d":
        cfg = wl.Config(2, 16, 1, 7, amrb.EQ_EULER)
        host = wl.build_static_tree(cfg, 3, (0.3, 0.15))
    else:
        cfg = wl.Config(3, 8, 1, 5, amrb.EQ_EULER)
        host = wl.build_static_tree(cfg, 2, (0.3,))
    storage = amrb.STORAGE_INTERIOR if name.endswith("interior") else amrb.STORAGE_PADDED
    steps = 7
    ids = host.ids()
    transport = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    sol = mg.ShardedSolver(cfg, host, rank, world, local, dist, torch, storage=storage, transport=transport)
    if transport == "p2p" and sol.transport != "p2p":
        raise SystemExit("the peer-memory exchange could not be set up on this box")
    sol.upload_interior(wl.initial_condition(sol.ids, cfg))
    sol.halo_exchange()
    for overlap in (True, False):
        sol.advance_batch_async(steps, overlap=overlap)
        acc, n, dts = sol.finish_advance_batch(steps)
    # optionally: the same batch recorded into a CUDA graph and replayed (even step count)
    use_graph = os.environ.get("AMRB_SELFTEST_GRAPH", "0") == "1"
    sol.advance_batch_async(6)
    sol.finish_advance_batch(6)
    if use_graph:
        sol.advance_batch_graph(6)
        torch.cuda.synchronize()
    else:
        sol.advance_batch_async(6, overlap=False)
    acc, n, dts = sol.finish_advance_batch(6)
    # a batch longer than any before: the library re-allocates its per-step scalar arrays, the views of
    # the dt-min slots the all-reduce writes through must follow (stale views = every rank steps with its
    # LOCAL CFL minimum and the shards drift apart)
    sol.advance_batch_async(70)
    acc, n, dts = sol.finish_advance_batch(70)
    mine = sol.download_interior()
    halo = np.stack([sol.pool.download(f, sol.plan.n_owned) for f in range(cfg.nvar)])
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, halo, acc, n))
    ok = True
    if rank == 0:
        lay = amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage)
        pool = amrb.DevicePool(lay, len(ids), local)
        pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
        pool.set_topology(*host.tables())
        ic = wl.initial_condition(ids, cfg)
        for f in range(cfg.nvar):
            pool.upload_interior(f, ic[f])
        pool.halo_exchange()
        for st in (steps, steps, 6, 6, 70):
            pool.advance_batch_async(st)
            acc1, n1, _ = pool.finish_advance_batch(st)
        ref = np.stack([pool.download_interior(f, len(ids)) for f in range(cfg.nvar)])
        refh = np.stack([pool.download(f, len(ids)) for f in range(cfg.nvar)])
        got = np.concatenate([g[0].reshape(cfg.nvar, -1, cfg.data) for g in gathered], axis=1)
        goth = np.concatenate([g[1] for g in gathered], axis=1)
        # face halos (corners excluded) must match too: they were filled from ghost slots
        R, h, S = cfg.rank, cfg.halo, cfg.size
        idx = np.indices((cfg.psize,) * R)
        outside = sum(((idx[k] < h) | (idx[k] >= h + S)).astype(int) for k in range(R))
        mask = (outside <= 1).ravel()
        ok = np.array_equal(got, ref) and np.array_equal(goth[..., mask], refh[..., mask])
        ok = ok and all(g[2] == acc1 and g[3] == n1 for g in gathered)
        print("multigpu selftest %s [%s] world=%d patches=%d: %s (sum dt %.17g, steps %d)"
              % (name, sol.transport, world, len(ids), "PARITY" if ok else "MISMATCH", acc1, n1))
        pool.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    rc = 0 if int(flag.item()) == 1 else 1
    if use_graph:
        # a process group whose NCCL kernels live in a recorded graph does not tear down cleanly
        sol.graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(rc)
    sol.close()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
