"""torchrun entry: active AMR (3D advection pulse, criterion -> global reconstruct -> re-slicing of the Morton
ranges over NCCL point-to-point copies -> new halo tables, ghost slots and exchange lists) on N GPUs compared,
on rank 0, with the same loop on ONE pool: leaf ids after every cycle, dt sequence and final state must be
identical.  Exit code 0 = parity."""
import importlib
import os
import sys

import numpy as np


def main():
    import torch
    import torch.distributed as dist

    amrb = importlib.import_module("gpu-amr_b200")
    aa = importlib.import_module("gpu-amr_b200.active_amr")
    wl = importlib.import_module("gpu-amr_b200.workloads")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    storage = amrb.STORAGE_INTERIOR if (len(sys.argv) < 2 or sys.argv[1] == "interior") else amrb.STORAGE_PADDED
    transport = sys.argv[2] if len(sys.argv) > 2 else "p2p"
    p = dict(aa.C5, depth=6, min_level=2, max_level=5, capacity=20000)
    cfg = wl.Config(3, 8, 1, p["depth"], amrb.EQ_ADVECTION)
    cycles = 6
    run = aa.ActiveAmr(cfg, torch, device=local, dist=dist, rank=rank, world=world, p=p, storage=storage,
                       transport=transport, host_ic=True)
    ids_log, dts = [run.host.ids().copy()], []
    for _ in range(cycles):
        dts.append(run.run(1))
        ids_log.append(run.host.ids().copy())
    mine = np.stack([run.sol.pool.download_interior(0, run.sol.plan.n_owned)])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        one = aa.ActiveAmr(cfg, torch, device=local, p=p, storage=storage, host_ic=True)
        ok = np.array_equal(one.host.ids(), ids_log[0])
        for c in range(cycles):
            d1 = one.run(1)
            ok = ok and np.array_equal(d1, dts[c]) and np.array_equal(one.host.ids(), ids_log[c + 1])
        ref = one.pool.download_interior(0, one.host.size)
        got = np.concatenate([g[0] for g in gathered], axis=0)
        ok = ok and got.shape == ref.shape and np.array_equal(got, ref)
        levels = sorted(set((one.host.ids() & np.uint64(63)).astype(int).tolist()))
        print("active AMR selftest [%s, %s] world=%d: %s (patches %d -> %d, levels %s, %d of %d reconstructs changed the mesh)"
              % ("interior" if storage else "padded", run.sol.transport, world, "PARITY" if ok else "MISMATCH",
                 len(ids_log[0]), one.host.size, levels, one.changed, one.regrids))
        ok = ok and one.changed >= 3
        one.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    run.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
