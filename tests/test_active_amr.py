"""Active AMR in 3D (BASELINE config C5 family: advection of a Gaussian pulse, 8^3 patches, >= 3 levels,
criterion on the patch maximum, reconstruct every 5 steps) on interior-only pools, against the pinned C
oracle driven through the same loop: refinement flags (device criterion over interior + gathered face
ghosts vs the oracle's padded patches), leaf ids after every reconstruct, dt sequence and final state."""
import importlib

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _oracle_flags(tree, cfg, p):
    data = tree.get_padded()[0].reshape(tree.size, -1)
    mx = np.maximum(data.max(axis=1), 0.0)
    lv = (tree.ids() & np.uint64(63)).astype(int)
    f = np.zeros(tree.size, np.int8)
    f[(lv < p["max_level"]) & (mx > p["refine"])] = 1
    f[(f == 0) & (lv > p["min_level"]) & (mx < p["coarsen"])] = 2
    return f


@pytest.mark.parametrize("device_regrid", [True, False], ids=["device_selection", "host_selection"])
@pytest.mark.parametrize("storage", [1, 0], ids=["interior", "padded"])
def test_active_amr_3d_advection_matches_oracle(amrb, storage, device_regrid):
    import torch

    aa = importlib.import_module("gpu-amr_b200.active_amr")
    wl = importlib.import_module("gpu-amr_b200.workloads")
    p = dict(aa.C5, depth=5, min_level=2, max_level=4, capacity=6000)
    cfg = wl.Config(3, 8, 1, p["depth"], amrb.EQ_ADVECTION)
    run = aa.ActiveAmr(cfg, torch, p=p, storage=storage, host_ic=True, device_regrid=device_regrid)

    ocfg = O.Config.from_name("r3_s8_h1_d5_adv")
    orc = O.OracleTree(ocfg, capacity=6000)
    for _ in range(p["min_level"]):
        orc.reconstruct(O.flags_all(orc.ids()))
    for _ in range(p["max_level"] - p["min_level"] + 1):
        orc.set_interior(wl.initial_condition(orc.ids(), cfg))
        orc.halo_exchange()
        if not orc.reconstruct(_oracle_flags(orc, cfg, p)):
            break
        orc.halo_exchange()
    orc.set_interior(wl.initial_condition(orc.ids(), cfg))
    orc.halo_exchange()
    assert np.array_equal(run.host.ids(), orc.ids())
    levels = set((orc.ids() & np.uint64(63)).astype(int).tolist())
    assert levels == {2, 3, 4}, levels

    changed = 0
    for cycle in range(8):
        dts = run.run(1)
        _, n, odts = orc.advance_batch(p["interval"])
        np.testing.assert_allclose(dts, odts, rtol=TOL, atol=0)
        # run.run() has already reconstructed with the device criterion; the oracle gets its own flags
        before = orc.size
        if orc.reconstruct(_oracle_flags(orc, cfg, p)):
            changed += 1
        orc.halo_exchange()
        assert np.array_equal(run.host.ids(), orc.ids()), ("cycle", cycle, before, orc.size)
    assert changed >= 2 and run.changed >= 2, "the pulse must move the refined region"
    got = np.stack([run.pool.download_interior(0, run.host.size)]).reshape(1, -1)
    ref = orc.get_interior().reshape(1, -1)
    assert np.abs(got - ref).max() / np.abs(ref).max() <= TOL
    run.close()
