"""N-GPU sharded path (Morton-range partition, ghost-slab exchange over NCCL, all-reduce(min) dt)
against the single-GPU pool: bit-exact state, halos and step sizes.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("case", ["2d", "3d"])
def test_sharded_matches_single_gpu(case):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "gpu-amr_b200", "selftest_multigpu.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PARITY" in r.stdout


@pytest.mark.parametrize("case,world", [("2d", 3), ("3d", 2), ("3d", 5)])
def test_sharding_logic_in_one_process(case, world):
    """The sharded path on ONE GPU: W shards of a multi-level mesh (Morton ranges, ghost slots, slab
    pack / unpack, interior / boundary launches, per-step CFL minimum) driven phase by phase in this
    process, slabs moved by device-to-device copies instead of NCCL.  State, materialised halos and
    the dt sequence must be bit-identical to the single pool — with and without the interior /
    boundary split."""
    import importlib

    import numpy as np
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    amrb = importlib.import_module("gpu-amr_b200")
    mg = importlib.import_module("gpu-amr_b200.multigpu")
    wl = importlib.import_module("gpu-amr_b200.workloads")
    if case == "2d":
        cfg = wl.Config(2, 16, 1, 7, amrb.EQ_EULER)
        host = wl.build_static_tree(cfg, 3, (0.3, 0.15))
    else:
        cfg = wl.Config(3, 8, 1, 5, amrb.EQ_EULER)
        host = wl.build_static_tree(cfg, 2, (0.3,))
    ids = host.ids()
    steps = 5
    cl = mg.LocalCluster(cfg, host, world, 0, torch)
    for s in cl.sols:
        s.upload_interior(wl.initial_condition(s.ids, cfg))
    cl.halo_exchange()
    outs = [cl.advance_batch(steps, overlap=ov) for ov in (True, False)]
    got = np.concatenate([s.download_interior().reshape(cfg.nvar, -1, cfg.data) for s in cl.sols], axis=1)
    goth = np.concatenate([np.stack([s.pool.download(f, s.plan.n_owned) for f in range(cfg.nvar)])
                           for s in cl.sols], axis=1)

    pool = amrb.DevicePool(amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth), len(ids))
    pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
    pool.set_topology(*host.tables())
    ic = wl.initial_condition(ids, cfg)
    for f in range(cfg.nvar):
        pool.upload_interior(f, ic[f])
    pool.halo_exchange()
    ref_out = []
    for _ in range(2):
        pool.advance_batch_async(steps)
        ref_out.append(pool.finish_advance_batch(steps))
    ref = np.stack([pool.download_interior(f, len(ids)) for f in range(cfg.nvar)]).reshape(cfg.nvar, -1, cfg.data)
    refh = np.stack([pool.download(f, len(ids)) for f in range(cfg.nvar)])
    R, h, S = cfg.rank, cfg.halo, cfg.size
    idx = np.indices((cfg.psize,) * R)
    outside = sum(((idx[k] < h) | (idx[k] >= h + S)).astype(int) for k in range(R))
    mask = (outside <= 1).ravel()
    assert np.array_equal(got, ref)
    assert np.array_equal(goth[..., mask], refh[..., mask])
    for per_rank, (acc1, n1, dts1) in zip(outs, ref_out):
        for acc, n, dts in per_rank:
            assert acc == acc1 and n == n1 and np.array_equal(np.asarray(dts), np.asarray(dts1))
    cl.close()
    pool.close()
