"""reconstruct_tree on the device (amrb_pool_reconstruct_device: eligibility, 2:1 ripple, coarsening veto, new
leaf ids, transfer plan, data motion, new tables) against the host selection (amrb_tree_reconstruct, itself
pinned bit-exact against dumps of the unmodified reference: tests/test_host_topology.py,
test_oracle_golden.py): identical leaf ids, identical plans, identical tables and identical data after every
pass, over multi-level 2D / 3D trees with hash-driven refine + coarsen flags (ripples and vetoes included)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfgname,storage", [("r2_s8_h1_d7_euler", 0), ("r3_s4_h1_d5_euler", 0),
                                             ("r3_s8_h1_d5_adv", 1), ("r2_s10_h2_d7_adv", 0)])
def test_device_reconstruct_matches_host_selection(amrb, cfgname, storage):
    cfg = O.Config.from_name(cfgname)
    cap = 30000
    host = amrb.DeviceTree(cfg, capacity=cap, storage=storage)       # host selection + device data motion
    lay = amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage)
    pool = amrb.DevicePool(lay, cap)                                  # everything on the device
    pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
    pool.set_topology_from_ids(host.ids())
    rng = np.random.default_rng(7)
    import torch
    passes = [("all", None), ("all", None)] + [("hash", s) for s in range(31, 43)]
    n_changed = n_merge = n_ripple = 0
    for kind, seed in passes:
        ids = host.ids()
        n = len(ids)
        # same data on both sides before the pass (the plan moves it)
        data = rng.standard_normal((cfg.nvar, n) + (cfg.size,) * cfg.rank)
        host.set_interior(data)
        for f in range(cfg.nvar):
            pool.upload_interior(f, data[f])
        maxl = min(cfg.depth, 5 if cfg.rank == 2 else 4)
        # 3D: a family of 8 siblings recombines only if all are flagged -> coarsen-heavy flags on odd passes
        pr, pc = (300, 350) if (cfg.rank == 2 or seed is None or seed % 2 == 0) else (60, 900)
        flags = O.flags_all(ids) if kind == "all" else O.flags_hash(ids, seed, pr, pc, 1, maxl)
        if kind == "all" and n * (1 << cfg.rank) > 3000:
            flags = O.flags_hash(ids, 5, 200, 0, 1, maxl)
        d_flags = torch.from_numpy(flags.copy()).cuda()
        changed_h = host.reconstruct(flags)
        changed_d, new_n = pool.reconstruct_device(d_flags.data_ptr())
        assert bool(changed_h) == bool(changed_d), (kind, seed)
        assert new_n == host.size == pool.size
        assert np.array_equal(pool.get_ids(), host.ids()), (kind, seed, "leaf ids")
        if changed_h:
            n_changed += 1
            hk, hs, hc = host.tree.plan()
            dk, ds, dc = pool.get_plan()
            assert np.array_equal(hk, dk) and np.array_equal(hs, ds) and np.array_equal(hc, dc), (kind, seed, "plan")
            n_merge += int((hk == 2).any())
            n_ripple += int((np.bincount(hs[hk == 1], minlength=n) > 0).sum() > int((flags == 1).sum()))
        lv_h, meta_h, nbr_h = host.pool.get_tables(host.size, cfg.rank)
        lv_d, meta_d, nbr_d = pool.get_tables(pool.size, cfg.rank)
        assert np.array_equal(lv_h, lv_d) and np.array_equal(meta_h, meta_d) and np.array_equal(nbr_h, nbr_d)
        a = host.get_interior()
        b = np.stack([pool.download_interior(f, pool.size).reshape((pool.size,) + (cfg.size,) * cfg.rank)
                      for f in range(cfg.nvar)])
        assert np.array_equal(a, b), (kind, seed, "data after the plan")
    assert n_changed >= 8 and n_merge >= 2 and n_ripple >= 1, (n_changed, n_merge, n_ripple)
    pool.close()
    host.pool.close()


def test_device_reconstruct_needs_device_ids(amrb):
    cfg = O.Config.from_name("r2_s8_h1_d7_euler")
    t = amrb.HostTree(cfg.rank, cfg.depth)
    t.reconstruct(O.flags_all(t.ids()))
    pool = amrb.DevicePool(amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth), 64)
    pool.set_topology(*t.tables())                       # host tables: no leaf ids on the device
    with pytest.raises(amrb.AmrbError):
        pool.reconstruct_device(None)
    pool.close()
