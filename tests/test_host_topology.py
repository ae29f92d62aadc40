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


# ---- edge cases of the leaf store (the reference leaves most of these undefined: `append` never
# checks m_capacity, ndtree.hpp:1304-1316; the product must refuse loudly instead)
def test_capacity_overflow_is_refused_and_leaves_the_tree_unchanged(amrb):
    t = amrb.HostTree(2, 7)
    assert t.reconstruct(np.full(t.size, amrb.REFINE, np.int8), capacity=4) == 1      # 1 -> 4 fits
    before = t.ids()
    with pytest.raises(amrb.AmrbError, match="capacity"):
        t.reconstruct(np.full(t.size, amrb.REFINE, np.int8), capacity=15)             # 4 -> 16 does not
    assert np.array_equal(t.ids(), before)
    assert t.reconstruct(np.full(t.size, amrb.REFINE, np.int8), capacity=16) == 1


def test_level_limits_and_identity_reconstructs(amrb):
    t = amrb.HostTree(3, 2)                                      # depth 2: levels 0..2
    for _ in range(2):
        assert t.reconstruct(np.full(t.size, amrb.REFINE, np.int8)) == 1
    assert t.size == 64 and ((t.ids() & np.uint64(63)) == 2).all()
    # refining leaves at the maximum level and a pass with only Stable flags change nothing
    assert t.reconstruct(np.full(t.size, amrb.REFINE, np.int8)) == 0 and t.size == 64
    assert t.reconstruct(np.zeros(t.size, np.int8)) == 0
    # the root cannot be coarsened; coarsening needs ALL 2^R siblings flagged
    flags = np.full(t.size, amrb.COARSEN, np.int8)
    flags[5] = 0
    assert t.reconstruct(flags) == 1 and t.size == 64 - 7 * 7    # 7 of 8 families merge
    r = amrb.HostTree(2, 3)
    assert r.reconstruct(np.full(1, amrb.COARSEN, np.int8)) == 0 and r.size == 1


def test_ripple_keeps_the_tree_two_to_one_balanced(amrb):
    """refine one corner leaf repeatedly: the 2:1 ripple (ndtree.hpp:1127-1166) must split coarser
    neighbors, also across the periodic wrap, so that no face ever sees a level jump above one"""
    t = amrb.HostTree(2, 7)
    t.reconstruct(np.full(1, amrb.REFINE, np.int8))
    for _ in range(5):
        f = np.zeros(t.size, np.int8)
        f[0] = amrb.REFINE
        assert t.reconstruct(f) == 1
        levels, rel, nbr, quad = t.tables()
        assert (rel != 0).all()                                  # periodic: every face has a neighbor
        for i in range(t.size):
            for d in range(4):
                k = 2 if rel[i, d] == 2 else 1
                for j in nbr[i, d, :k]:
                    assert abs(int(levels[j]) - int(levels[i])) <= 1
    assert len(np.unique(levels)) >= 5


def test_argument_validation(amrb):
    import ctypes as C

    L = amrb.lib()
    h = C.c_void_p()
    assert L.amrb_tree_create(4, 3, C.byref(h)) != 0 and b"rank" in L.amrb_last_error()
    assert L.amrb_tree_create(2, 0, C.byref(h)) != 0
    assert L.amrb_tree_create(2, 3, None) != 0
    assert L.amrb_tree_size(None) == 0
