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


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("case", ["2d", "3d", "3d_interior"])
def test_sharded_matches_single_gpu(case, transport):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "gpu-amr_b200", "selftest_multigpu.py"), case, transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PARITY" in r.stdout


@pytest.mark.parametrize("storage,transport", [("interior", "p2p"), ("interior", "nccl"), ("padded", "p2p")])
def test_active_amr_sharded_matches_single_gpu(storage, transport):
    """BASELINE config C5 family on >= 2 GPUs: criterion on every shard, one global reconstruct, re-slicing of
    the Morton ranges (whole old patches travel between curve neighbours), new tables / ghost slots / exchange
    lists -- against the same loop on one pool (gpu-amr_b200/selftest_active_amr.py)."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29633",
           os.path.join(ROOT, "gpu-amr_b200", "selftest_active_amr.py"), storage, transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PARITY" in r.stdout


@pytest.mark.parametrize("transport", ["copy", "p2p"])
@pytest.mark.parametrize("case,world", [("2d", 3), ("3d", 2), ("3d", 5), ("3d_interior", 3)])
def test_sharding_logic_in_one_process(case, world, transport):
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
    storage = amrb.STORAGE_INTERIOR if case.endswith("interior") else amrb.STORAGE_PADDED
    ids = host.ids()
    steps = 5
    cl = mg.LocalCluster(cfg, host, world, 0, torch, storage=storage, transport=transport)
    for s in cl.sols:
        s.upload_interior(wl.initial_condition(s.ids, cfg))
    cl.halo_exchange()
    outs = [cl.advance_batch(steps, overlap=ov) for ov in (True, False)]
    got = np.concatenate([s.download_interior().reshape(cfg.nvar, -1, cfg.data) for s in cl.sols], axis=1)
    goth = np.concatenate([np.stack([s.pool.download(f, s.plan.n_owned) for f in range(cfg.nvar)])
                           for s in cl.sols], axis=1)

    pool = amrb.DevicePool(amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage), len(ids))
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


@pytest.mark.parametrize("case,world", [("2d", 3), ("3d", 3), ("3d_interior", 3)])
def test_reslicing_after_reconstruct_in_one_process(case, world):
    """Active AMR on a sharded mesh, on ONE GPU through LocalCluster: steps, the criterion
    on every shard, one global reconstruct (refine + coarsen + 2:1 ripple), re-slicing — old patches move
    to the ranks that need them, every shard applies its slice of the transfer plan and installs its new
    tables and ghost slots — and more steps.  Leaf ids, flags, state and dt sequence must be identical to
    the single-pool DeviceTree doing the same."""
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
        base, radii, thr = 3, (0.3, 0.15), (0.62, 0.52)
    else:
        cfg = wl.Config(3, 8, 1, 5, amrb.EQ_EULER)
        base, radii, thr = 2, (0.3,), (0.62, 0.52)
    steps, cap = 4, 4096
    storage = amrb.STORAGE_INTERIOR if case.endswith("interior") else amrb.STORAGE_PADDED

    # single pool (the checker of this test)
    one = amrb.DeviceTree(cfg, capacity=cap, storage=storage)
    host0 = wl.build_static_tree(cfg, base, radii)
    # replay the same refinement history on the DeviceTree's own host tree
    for _ in range(base):
        one.reconstruct(wl.flags_refine_all(one.ids()))
    for i, r in enumerate(radii):
        one.reconstruct(wl.flags_ball(one.ids(), cfg, r, base + i + 1, (0.5, 0.5, 0.5)))
    assert np.array_equal(one.ids(), host0.ids())
    ic = wl.initial_condition(one.ids(), cfg)
    one.set_interior(ic)
    one.halo_exchange()
    ref_a = one.advance_batch(steps)
    flags1 = one.pool.patch_max_flags(0, thr[0], thr[1], 1, cfg.depth if case == "2d" else 4)

    # W shards in this process
    host = wl.build_static_tree(cfg, base, radii)
    cl = mg.LocalCluster(cfg, host, world, 0, torch, capacity=cap, storage=storage)
    for s in cl.sols:
        s.upload_interior(wl.initial_condition(s.ids, cfg))
    cl.halo_exchange()
    out_a = cl.advance_batch(steps)
    flags = cl.patch_max_flags(0, thr[0], thr[1], 1, cfg.depth if case == "2d" else 4)
    assert np.array_equal(flags, flags1) and (flags == 1).any() and (flags == 2).any()

    old_size = host.size
    assert host.reconstruct(flags, cap) == 1
    assert one.reconstruct(flags1) == 1
    assert np.array_equal(host.ids(), one.ids()) and host.size != old_size
    rp = cl.reshard(host, old_size, host.plan())
    if case == "2d":
        assert any(r != q for (r, q) in rp.moves), "this case moves patches between ranks"
    cl.halo_exchange()
    one.halo_exchange()
    out_b = cl.advance_batch(steps)
    ref_b = one.advance_batch(steps)

    got = np.concatenate([s.download_interior().reshape(cfg.nvar, -1, cfg.data) for s in cl.sols], axis=1)
    ref = one.get_interior().reshape(cfg.nvar, -1, cfg.data)
    assert np.array_equal(got, ref)
    for per_rank, (acc1, n1, dts1) in ((out_a, ref_a), (out_b, ref_b)):
        for acc, n, dts in per_rank:
            assert acc == acc1 and n == n1 and np.array_equal(np.asarray(dts), np.asarray(dts1))
    cl.close()
