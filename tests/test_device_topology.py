"""Row 8f.4: neighbor / halo tables built on the device from the ascending leaf ids
(amrb_pool_set_topology_from_ids, topology_kernel) against the host builder (amrb_tree_tables, itself
pinned bit for bit against the reference's tables by the golden fixtures): relation, neighbor indices,
contact quadrants and levels must be identical for single roots, the level-1 periodic wrap, and
multi-level 2D / 3D trees produced by hash refinement + coarsening."""
import importlib

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    ("r2_s8_h1_d7_euler", ""),                                       # periodic root: its own neighbor
    ("r2_s8_h1_d7_adv", "A\nX"),                                     # level 1: +/- neighbor identical
    ("r2_s8_h1_d7_euler", "A\nX\nR 1\nX"),                           # finer across the wrap
    ("r2_s16_h1_d7_euler", "A\nX\nA\nX\nH 1 300 0 1 4\nX\nH 2 250 300 1 5\nX\nH 3 300 200 1 6\nX"),
    ("r3_s4_h1_d5_euler", "A\nX\nH 21 400 0 1 3\nX\nH 22 150 300 1 4\nX"),
    ("r3_s8_h1_d5_adv", "A\nX\nA\nX\nH 7 300 0 1 3\nX\nH 8 250 300 1 4\nX"),
]


def _compact(rank, rel, quad):
    meta = rel.astype(np.int32).copy()
    for k in range(rank):
        meta |= np.where(rel == 3, (quad[..., k].astype(np.int32) & 1) << (2 + k), 0)
    return meta.astype(np.uint8)


@pytest.mark.parametrize("cfgname,script", CASES, ids=[c[0] + "#%d" % i for i, c in enumerate(CASES)])
def test_device_tables_equal_host_tables(amrb, cfgname, script):
    cfg = O.Config.from_name(cfgname)
    dev = amrb.DeviceTree(cfg, capacity=8192)
    if script:
        O.run_script(dev, script)
    levels, rel, nbr, quad = dev.tree.tables()
    n = dev.size
    dev.pool.set_topology_from_ids(dev.ids())
    dl, dm, dn = dev.pool.get_tables(n, cfg.rank)
    assert np.array_equal(dl, levels)
    assert np.array_equal(dm, _compact(cfg.rank, rel, quad))
    kf = 1 << (cfg.rank - 1)
    need = np.where(rel == 2, kf, np.where(rel == 0, 0, 1))          # entries that carry an index
    mask = np.arange(kf)[None, None, :] < need[..., None]
    assert np.array_equal(np.where(mask, dn, -1), np.where(mask, nbr, -1))
    assert (dn[~mask] == -1).all()
    assert len(np.unique(levels)) >= (1 if not script else 1)


def test_device_tables_at_scale(amrb):
    """2.6e5 leaves (uniform level 6 in 3D): every relation is `same` and the neighbor index is the
    Morton rank of the shifted anchor; spot-checked against the host builder on a sample."""
    wl = importlib.import_module("gpu-amr_b200.workloads")
    cfg = wl.Config(3, 8, 1, 7, amrb.EQ_ADVECTION)
    host = wl.build_static_tree(cfg, 6, ())
    ids = host.ids()
    pool = amrb.DevicePool(amrb.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth), len(ids))
    pool.set_topology_from_ids(ids)
    dl, dm, dn = pool.get_tables(len(ids), 3)
    levels, rel, nbr, quad = host.tables()
    assert (dm == 1).all() and np.array_equal(dl, levels)
    assert np.array_equal(dn[:, :, 0], nbr[:, :, 0])
    pool.close()
