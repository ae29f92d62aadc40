"""one-GPU reproduction of the N = 8 weak-scaled C2 mesh: 8 shards in one process (LocalCluster), peer-memory
and copy transports, against the single pool; reports where the states differ"""
import importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amrb = importlib.import_module("gpu-amr_b200")
mg = importlib.import_module("gpu-amr_b200.multigpu")
wl = importlib.import_module("gpu-amr_b200.workloads")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg, host, base, radius = mg.weak_scaled_tree(wl, world)
ids = host.ids()
print("mesh", host.size, "patches, base", base, flush=True)
steps = 3
ic = wl.initial_condition(ids, cfg)
pool = amrb.DevicePool(amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth), len(ids))
pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
pool.set_topology_from_ids(ids)
for f in range(cfg.nvar):
    pool.upload_interior(f, ic[f])
pool.halo_exchange()
pool.advance_batch_async(steps)
ref_out = pool.finish_advance_batch(steps)
ref = np.stack([pool.download_interior(f, len(ids)) for f in range(cfg.nvar)])
pool.close()
for transport in ("copy", "p2p"):
    cl = mg.LocalCluster(cfg, host, world, 0, torch, transport=transport)
    for s in cl.sols:
        s.upload_interior(ic[:, s.plan.lo:s.plan.hi])
    cl.halo_exchange()
    for ov in (False, True):
        if ov:
            for s in cl.sols:
                s.upload_interior(ic[:, s.plan.lo:s.plan.hi])
            cl.halo_exchange()
        outs = cl.advance_batch(steps, overlap=ov)
        got = np.concatenate([s.download_interior().reshape(cfg.nvar, -1, cfg.data) for s in cl.sols], axis=1)
        diff = np.abs(got - ref.reshape(got.shape))
        bad = np.argwhere(diff.max(axis=2) > 0)
        print(transport, "overlap" if ov else "single launch", "identical" if len(bad) == 0 else
              "DIFFERENT: %d field-patch pairs, max |diff| %.3e (field max %.3e), first %s, dt equal %s"
              % (len(bad), diff.max(), np.abs(ref).max(), bad[:5].tolist(),
                 all(np.array_equal(np.asarray(o[2]), np.asarray(ref_out[2])) for o in outs)), flush=True)
        if len(bad):
            owners = np.searchsorted(cl.sols[0].plan.bounds[1:], bad[:, 1], side="right")
            print("   owners of differing patches:", np.bincount(owners, minlength=world).tolist(),
                  " boundary?", [int(b[1] - cl.sols[o].plan.lo) in set(cl.sols[o].plan.boundary.tolist()) for b, o in zip(bad[:8], owners[:8])])
    cl.close()
