"""Product host topology (amrb_tree_*: set-based leaf store, 2:1 ripple, coarsening veto, neighbor
tables derived by key lookup) against the dumps of the unmodified reference: leaf ids, relations,
neighbor linear indices and contact quadrants must be bit-exact after every scripted reconstruct."""
import numpy as np
import pytest

import oracle as O
from golden_util import fixtures, load, tags_in_order


class TopologyOnly:
    """Adapter so oracle.run_script can drive the product HostTree without any device data."""

    def __init__(self, amrb, cfg):
        self.cfg = cfg
        self.t = amrb.HostTree(cfg.rank, cfg.depth)

    size = property(lambda self: self.t.size)

    def ids(self):
        return self.t.ids()

    def reconstruct(self, flags):
        return self.t.reconstruct(flags)

    def tables(self):
        return self.t.tables()[1:]

    def get_padded(self):
        return np.zeros((self.cfg.nvar, self.size, self.cfg.flat))

    def set_padded(self, d):
        pass

    def set_interior(self, d):
        pass

    def halo_exchange(self):
        pass

    def advance(self):
        return 0.0


@pytest.mark.parametrize("name", fixtures())
def test_tables_match_reference(amrb, name):
    cfg, script, g = load(name)
    # state-dependent refinement never occurs in the scripts, so data ops can be no-ops
    out = O.run_script(TopologyOnly(amrb, cfg), script, ic_override=lambda t: None)
    for tag in tags_in_order(script):
        assert np.array_equal(out[tag + "/ids"], g[tag + "/ids"]), (name, tag, "leaf ids")
        assert np.array_equal(out[tag + "/rel"], g[tag + "/rel"]), (name, tag, "relation")
        assert np.array_equal(out[tag + "/nbr"], g[tag + "/nbr"]), (name, tag, "neighbor ids")
        assert np.array_equal(out[tag + "/quad"], g[tag + "/quad"]), (name, tag, "quadrant")


def test_plan_is_consistent(amrb):
    """every new leaf has exactly one source; restriction sources are 2^R consecutive old leaves"""
    cfg = O.Config.from_name("r3_s4_h1_d5_euler")
    t = amrb.HostTree(3, 5)
    t.reconstruct(O.flags_all(t.ids()))
    t.reconstruct(O.flags_all(t.ids()))
    old = t.ids()
    f = O.flags_hash(old, 5, 300, 500, 1, 4)
    assert t.reconstruct(f) == 1
    kind, src, child = t.plan()
    new = t.ids()
    assert len(kind) == len(new) and np.all(np.diff(new.astype(np.int64)) > 0)
    lv_old, lv_new = (old & np.uint64(63)).astype(int), (new & np.uint64(63)).astype(int)
    assert np.array_equal(lv_new[kind == 0], lv_old[src[kind == 0]])
    assert np.array_equal(lv_new[kind == 1], lv_old[src[kind == 1]] + 1)
    assert np.array_equal(lv_new[kind == 2], lv_old[src[kind == 2]] - 1)
    assert np.array_equal(new[kind == 0], old[src[kind == 0]])
    used = np.zeros(len(old), int)
    np.add.at(used, src[kind == 0], 1)
    for s in src[kind == 2]:
        used[s:s + 8] += 1
    for s in np.unique(src[kind == 1]):
        used[s] += 1
        assert sorted(child[(kind == 1) & (src == s)]) == list(range(8))
    assert np.all(used == 1)
    assert cfg.rank == 3
