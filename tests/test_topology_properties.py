"""Property tests of the host topology (hypothesis): for random sequences of refine / coarsen flag
vectors the leaf store must stay a valid 2:1-balanced periodic partition and its tables symmetric —
the invariants SURVEY Appendix A lists for the reference's tree (A1: ascending ids, linear index ==
rank of the id; A2: eligibility / ripple / veto keep the level jump across every face at most one)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st


def _volume(ids, rank, depth):
    lv = (ids & np.uint64(63)).astype(np.int64)
    return float(np.sum(2.0 ** (-rank * lv)))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(rank=st.sampled_from([2, 3]), seeds=st.lists(st.integers(0, 2 ** 31 - 1), min_size=2, max_size=6),
       p_ref=st.floats(0.05, 0.6), p_coa=st.floats(0.0, 0.6))
def test_random_flag_sequences_keep_the_tree_valid(amrb, rank, seeds, p_ref, p_coa):
    depth = 5 if rank == 2 else 4
    t = amrb.HostTree(rank, depth)
    t.reconstruct(np.full(1, amrb.REFINE, np.int8))
    for seed in seeds:
        rng = np.random.default_rng(seed)
        u = rng.random(t.size)
        flags = np.where(u < p_ref, amrb.REFINE, np.where(u > 1.0 - p_coa, amrb.COARSEN, 0)).astype(np.int8)
        t.reconstruct(flags, capacity=20000)
        ids = t.ids()
        # A1: strictly ascending ids; the leaves tile the periodic unit cube exactly once
        assert (np.diff(ids.astype(np.uint64)) > 0).all()
        assert abs(_volume(ids, rank, depth) - 1.0) < 1e-12
        levels, rel, nbr, quad = t.tables()
        nd, kf = 2 * rank, 1 << (rank - 1)
        assert (rel != 0).all()                                   # periodic: no boundary faces
        for i in range(t.size):
            for d in range(nd):
                r = rel[i, d]
                js = nbr[i, d, :kf] if r == 2 else nbr[i, d, :1]
                assert (js >= 0).all() and (js < t.size).all()
                for j in js:
                    dl = int(levels[j]) - int(levels[i])
                    assert dl == (0 if r == 1 else 1 if r == 2 else -1)      # A2: 2:1 balance
                    # symmetry: seen from j across the opposite face, i is there
                    back = rel[j, d ^ 1]
                    assert back == (1 if r == 1 else 3 if r == 2 else 2)
                    assert i in (nbr[j, d ^ 1, :kf] if back == 2 else nbr[j, d ^ 1, :1])


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(rank=st.sampled_from([2, 3]), seed=st.integers(0, 2 ** 31 - 1))
def test_transfer_plan_accounts_for_every_leaf(amrb, rank, seed):
    """plan of a pass: every old leaf is consumed exactly once (copied, split into 2^R children, or
    merged with its siblings) and the plan is monotone in Morton order (what ReshardPlan relies on)"""
    depth = 5 if rank == 2 else 4
    t = amrb.HostTree(rank, depth)
    for _ in range(2):
        t.reconstruct(np.full(t.size, amrb.REFINE, np.int8))
    rng = np.random.default_rng(seed)
    n_old = t.size
    flags = rng.choice([0, 1, 2], size=n_old, p=[0.4, 0.2, 0.4]).astype(np.int8)
    fam = (np.arange(n_old) // (1 << rank))                       # sibling families are contiguous here
    flags[np.isin(fam, rng.choice(fam.max() + 1, size=max(1, (fam.max() + 1) // 3), replace=False))] = amrb.COARSEN
    if not t.reconstruct(flags):
        return
    kind, src, child = t.plan()
    fan = 1 << rank
    assert (np.diff(src) >= 0).all()
    used = np.zeros(n_old, np.int64)
    for k, s in zip(kind, src):
        used[s:s + (fan if k == 2 else 1)] += 1 if k != 1 else 0
    split = np.unique(src[kind == 1])
    assert all((src[kind == 1] == s).sum() == fan for s in split)
    used[split] += 1
    assert (used == 1).all()
    assert sorted(child[(kind == 1) & (src == split[0])].tolist()) == list(range(fan)) if len(split) else True
