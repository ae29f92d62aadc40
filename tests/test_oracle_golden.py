"""Pins the oracle (oracle/amr_oracle.c) against dumps of the unmodified reference.

Every fixture was produced by oracle/gen_golden.py running oracle/_ref/ref_dump_<cfg> (reference
headers compiled with -O2 -ffp-contract=off).  Topology, neighbor tables and halo indexing must be
bit-exact; the oracle is compiled with the same strict fp flags, so states and dt are bit-exact too.
"""
import numpy as np
import pytest

import oracle as O
from golden_util import fixtures, ic_from_golden, load, tags_in_order


@pytest.mark.parametrize("name", fixtures())
def test_oracle_matches_reference_dump(name):
    cfg, script, g = load(name)
    tree = O.OracleTree(cfg)
    out = O.run_script(tree, script, ic_override=ic_from_golden(cfg, script, g))
    mask = O.face_halo_mask(cfg).ravel()
    for tag in tags_in_order(script):
        assert np.array_equal(out[tag + "/ids"], g[tag + "/ids"]), (name, tag, "leaf ids")
        assert np.array_equal(out[tag + "/rel"], g[tag + "/rel"]), (name, tag, "relation")
        assert np.array_equal(out[tag + "/nbr"], g[tag + "/nbr"]), (name, tag, "neighbor ids")
        assert np.array_equal(out[tag + "/quad"], g[tag + "/quad"]), (name, tag, "quadrant")
        assert np.array_equal(out[tag + "/dts"], g[tag + "/dts"]), (name, tag, "dt sequence")
        mine = out[tag + "/data"][..., mask]
        if tag + "/data" in g:
            assert np.array_equal(mine, g[tag + "/data"][..., mask]), (name, tag, "state")
        else:  # stats-only tag (full state too large to commit)
            full = out[tag + "/data"]
            assert np.array_equal(full[:, 0, :][..., mask], g[tag + "/first_patch"][..., mask])
            np.testing.assert_allclose(full[..., mask].sum(axis=(1, 2)), g[tag + "/sum"],
                                       rtol=1e-13, atol=1e-9)
            assert np.array_equal(full[..., mask].max(axis=(1, 2)), g[tag + "/max"])


def test_known_answer_ka2d():
    """BASELINE.md 'KA-2D' values, produced by the survey's own reference build."""
    cfg, script, g = load("ka2d")
    assert g["t0/ids"].shape[0] == 280
    lv = (g["t0/ids"] & np.uint64(63)).astype(int)
    assert lv.min() == 3 and lv.max() == 5
    dts = g["t200/dts"]
    assert abs(dts[0] - 0.2360834027907068) < 1e-15
    assert abs(dts[:10].sum() - 2.3461516365724315) < 1e-13
    assert abs(dts.sum() - 37.38921417430192) < 1e-11
    assert abs(g["t200/max"][0] - 1.1701661374019001) < 1e-14


def test_probe_halo_sources_are_interior_cells():
    """Index oracle: after a halo exchange over the probe pattern every face-halo
    cell holds the exact code of ONE interior source cell (same / coarser) or the mean of 2^R."""
    cfg, script, g = load("amr2d_euler")
    d = g["probe/data"].reshape((cfg.nvar, -1) + (cfg.psize,) * cfg.rank)
    rel = g["probe/rel"]
    h, S = cfg.halo, cfg.size
    P = d.shape[1]
    for p in range(P):
        for di in range(cfg.ndir):
            if rel[p, di] not in (1, 3):
                continue
            dim, pos = di // 2, di & 1
            sl = [slice(h, h + S)] * cfg.rank
            sl[dim] = slice(h + S, h + S + h) if pos else slice(0, h)
            v = d[0, p][tuple(sl)].astype(np.int64)
            src_patch, src_cell = v // 4096, v % 4096
            assert (src_patch == g["probe/nbr"][p, di, 0]).all()
            idx = np.stack(np.unravel_index(src_cell, (cfg.psize,) * cfg.rank))
            assert ((idx >= h) & (idx < h + S)).all()
