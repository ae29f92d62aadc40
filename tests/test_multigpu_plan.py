"""Host-side logic of the Morton-range sharding (gpu-amr_b200/multigpu.py: ShardPlan): ownership,
ghost slots, pack/unpack entry lists.  In-process cross-checks for several world sizes plus a
world_size-2 run over torch.distributed/gloo that moves real slabs between two processes and checks
every ghost slot against the global state."""
import importlib
import os
import socket

import numpy as np
import pytest

import oracle as O


def _tree(amrb, cfgname="r2_s8_h1_d7_euler"):
    cfg = O.Config.from_name(cfgname)
    t = amrb.HostTree(cfg.rank, cfg.depth)
    for _ in range(3 if cfg.rank == 2 else 2):
        t.reconstruct(O.flags_all(t.ids()))
    t.reconstruct(O.flags_hash(t.ids(), 21, 300, 0, 1, 5))
    t.reconstruct(O.flags_hash(t.ids(), 22, 250, 300, 1, 6))
    return cfg, t


def _slab_index(cfg, face, layers):
    """flat padded indices of the `layers` interior layers next to `face` (dim-major like the kernel)"""
    h, S, R = cfg.halo, cfg.size, cfg.rank
    dim, pos = face // 2, face & 1
    idx = np.indices((layers,) + (S,) * (R - 1)).reshape(R, -1)
    coords, t = [None] * R, 1
    for k in range(R):
        if k == dim:
            coords[k] = (h + S - 1 - idx[0]) if pos else (h + idx[0])
        else:
            coords[k] = h + idx[t]
            t += 1
    return np.ravel_multi_index(coords, (cfg.psize,) * R)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("cfgname", ["r2_s8_h1_d7_euler", "r3_s4_h1_d5_euler"])
def test_plans_are_mutually_consistent(amrb, world, cfgname):
    mg = importlib.import_module("gpu-amr_b200.multigpu")
    cfg, t = _tree(amrb, cfgname)
    levels, rel, nbr, quad = t.tables()
    plans = [mg.ShardPlan(levels, rel, nbr, quad, r, world) for r in range(world)]
    assert sum(p.n_owned for p in plans) == len(levels)
    for r, p in enumerate(plans):
        # every neighbor index is a local slot; ghosts are exactly the remote neighbors
        used = p.nbr[p.nbr >= 0]
        assert used.max() < p.n_total and set(used[used >= p.n_owned]) == set(range(p.n_owned, p.n_total))
        assert len(p.interior) + len(p.boundary) == p.n_owned
        assert (p.nbr[p.interior] < p.n_owned).all()
        # what r sends to q is what q expects from r, in the same order
        so = np.concatenate([[0], np.cumsum(p.send_counts)])
        for q, pq in enumerate(plans):
            ro = np.concatenate([[0], np.cumsum(pq.recv_counts)])
            assert np.array_equal(p.send_global[so[q]:so[q + 1]], pq.recv_global[ro[r]:ro[r + 1]])
            # peer-memory exchange: r's segment inside q's receive buffer starts where q expects it
            assert p.send_offsets[q] == ro[r]
        # every remote (patch, face) a halo will read is received
        need = set()
        for i in p.boundary:
            for d in range(2 * cfg.rank):
                for j in p.nbr[i, d]:
                    if j >= p.n_owned:
                        need.add((int(j), d ^ 1))
        assert need == set(map(tuple, p.recv_entries.tolist()))


@pytest.mark.parametrize("cfgname", ["r2_s8_h1_d7_euler", "r3_s4_h1_d5_euler"])
def test_layers_on_the_wire(amrb, cfgname):
    """entry word = face | layers << 4: all min(2h, S) layers (0) exactly for the faces a COARSER patch reads,
    h layers otherwise; both sides of every pair compute the same word"""
    mg = importlib.import_module("gpu-amr_b200.multigpu")
    cfg, t = _tree(amrb, cfgname)
    levels, rel, nbr, quad = t.tables()
    world = 3
    plans = [mg.ShardPlan(levels, rel, nbr, quad, r, world, halo_layers=cfg.halo) for r in range(world)]
    full = part = 0
    for r, p in enumerate(plans):
        so = np.concatenate([[0], np.cumsum(p.send_counts)])
        for q, pq in enumerate(plans):
            ro = np.concatenate([[0], np.cumsum(pq.recv_counts)])
            assert np.array_equal(p.send_global[so[q]:so[q + 1]], pq.recv_global[ro[r]:ro[r + 1]])
        glob = np.concatenate([np.arange(p.lo, p.hi), p.ghost_global])       # local slot -> global patch
        for slot, word in p.recv_entries.tolist():
            j, face, layers = int(glob[slot]), word & 15, word >> 4
            readers = [i for i in range(p.lo, p.hi) if j in nbr[i, face ^ 1]]
            assert readers, (j, face)
            coarser = any(levels[i] < levels[j] for i in readers)
            assert layers == (0 if coarser else cfg.halo), (j, face, layers, coarser)
            full += coarser
            part += not coarser
    assert full > 0 and part > 0


def _worker(rank, world, port, cfgname, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        amrb = importlib.import_module("gpu-amr_b200")
        mg = importlib.import_module("gpu-amr_b200.multigpu")
        cfg, t = _tree(amrb, cfgname)
        levels, rel, nbr, quad = t.tables()
        p = mg.ShardPlan(levels, rel, nbr, quad, rank, world)
        rng = np.random.default_rng(7)
        glob = rng.random((len(levels), cfg.flat))                      # same on every rank
        pool = np.zeros((p.n_total, cfg.flat))
        pool[:p.n_owned] = glob[p.lo:p.hi]
        T = min(2 * cfg.halo, cfg.size)
        slab = [_slab_index(cfg, f, T) for f in range(2 * cfg.rank)]
        send = np.concatenate([pool[e[0], slab[e[1]]] for e in p.send_entries]) if len(p.send_entries) else np.zeros(0)
        n = T * cfg.size ** (cfg.rank - 1)
        so = np.concatenate([[0], np.cumsum(p.send_counts)]) * n
        ro = np.concatenate([[0], np.cumsum(p.recv_counts)]) * n
        recv = np.zeros(int(ro[-1]))
        reqs, keep = [], []
        for peer in range(world):
            if peer == rank:
                continue
            if so[peer + 1] > so[peer]:
                ts = torch.from_numpy(send[so[peer]:so[peer + 1]].copy())
                keep.append(ts)
                reqs.append(dist.isend(ts, peer))
            if ro[peer + 1] > ro[peer]:
                tr = torch.zeros(int(ro[peer + 1] - ro[peer]), dtype=torch.float64)
                keep.append((tr, int(ro[peer])))
                reqs.append(dist.irecv(tr, peer))
        for r in reqs:
            r.wait()
        for k in keep:
            if isinstance(k, tuple):
                recv[k[1]:k[1] + len(k[0])] = k[0].numpy()
        for e, chunk in zip(p.recv_entries, recv.reshape(-1, n)):
            pool[e[0], slab[e[1]]] = chunk
        ok = True
        for e in p.recv_entries:
            g = p.ghost_global[e[0] - p.n_owned]
            ok &= np.array_equal(pool[e[0], slab[e[1]]], glob[g, slab[e[1]]])
        # dt all-reduce(min) path
        tmin = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        ok &= float(tmin) == 1.0
        q.put((rank, bool(ok), len(p.recv_entries)))
    finally:
        dist.destroy_process_group()


def test_two_process_slab_exchange_gloo(amrb):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "r2_s8_h1_d7_euler", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res) and all(n > 0 for _, _, n in res), res


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("cfgname", ["r2_s8_h1_d7_euler", "r3_s4_h1_d5_euler"])
def test_reshard_plan_moves_every_needed_patch_once(amrb, world, cfgname):
    """Re-slicing after a reconstruct, host logic: with the old leaves spread over W
    Morton ranges and a refine + coarsen pass applied, (1) the old leaves a rank needs for its new range
    are one contiguous old range and arrive exactly once, (2) the rank's re-indexed slice of the transfer
    plan, evaluated on what arrived, reproduces the global plan."""
    mg = importlib.import_module("gpu-amr_b200.multigpu")
    cfg, t = _tree(amrb, cfgname)
    old_ids = t.ids()
    n_old = len(old_ids)
    # Morton-contiguous blocks keep sibling families together: merge in the first half of the curve,
    # split in the last quarter
    flags = np.zeros(n_old, np.int8)
    flags[:n_old // 2] = amrb.COARSEN
    flags[-(n_old // 4):] = amrb.REFINE
    assert t.reconstruct(flags) == 1
    kind, src, child = t.plan()
    n_new, fan = t.size, 1 << cfg.rank
    assert (kind == 2).any() and (kind == 1).any() and (kind == 0).any()
    old_bounds = [(r * n_old) // world for r in range(world + 1)]
    rp = mg.ReshardPlan(old_bounds, n_new, kind, src, child, fan, world)

    def evaluate(k, s, c, payload):                  # what a new patch is made of, symbolically
        return ("copy", payload[s]) if k == 0 else ("prolong", payload[s], int(c)) if k == 1 else \
               ("restrict", tuple(payload[s:s + fan]))

    want = [evaluate(kind[j], src[j], child[j], np.arange(n_old)) for j in range(n_new)]
    got = []
    for q in range(world):
        a, b = rp.need[q]
        staged = np.full(b - a, -1, np.int64)        # old global index held by every staging slot
        for (r, qq), (first, count) in rp.moves.items():
            if qq != q:
                continue
            assert old_bounds[r] <= first and first + count <= old_bounds[r + 1]
            assert (staged[first - a:first - a + count] == -1).all()
            staged[first - a:first - a + count] = np.arange(first, first + count)
        assert (staged >= 0).all()
        ks, ss, cs = rp.sub[q]
        assert len(ks) == rp.new_bounds[q + 1] - rp.new_bounds[q]
        got += [evaluate(ks[j], ss[j], cs[j], staged) for j in range(len(ks))]
    assert got == want


def _reshard_worker(rank, world, port, cfgname, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        amrb = importlib.import_module("gpu-amr_b200")
        mg = importlib.import_module("gpu-amr_b200.multigpu")
        cfg, t = _tree(amrb, cfgname)
        n_old = t.size
        flags = np.zeros(n_old, np.int8)
        flags[:n_old // 2] = amrb.COARSEN
        flags[-(n_old // 4):] = amrb.REFINE
        assert t.reconstruct(flags) == 1
        kind, src, child = t.plan()
        old_bounds = [(r * n_old) // world for r in range(world + 1)]
        rp = mg.ReshardPlan(old_bounds, t.size, kind, src, child, 1 << cfg.rank, world)
        flat, nv = 6, 2                                             # tiny stand-in patches
        glob = [np.arange(n_old * flat, dtype=np.float64) + 1e6 * f for f in range(nv)]
        lo, hi = old_bounds[rank], old_bounds[rank + 1]
        cap = max(hi - lo, rp.staging_slots(rank)) + 4
        cur = [torch.zeros(cap * flat, dtype=torch.float64) for _ in range(nv)]
        nxt = [torch.full((cap * flat,), -1.0, dtype=torch.float64) for _ in range(nv)]
        for f in range(nv):
            cur[f][:(hi - lo) * flat] = torch.from_numpy(glob[f][lo * flat:hi * flat])
        mg.migrate_old_patches(rp, rank, flat, cur, nxt, dist)
        a, b = rp.need[rank]
        ok = all(np.array_equal(nxt[f][:(b - a) * flat].numpy(), glob[f][a * flat:b * flat]) for f in range(nv))
        moved = sum(c for (r, qq), (_, c) in rp.moves.items() if qq == rank and r != rank)
        q.put((rank, bool(ok), int(moved)))
    finally:
        dist.destroy_process_group()


def test_patch_migration_over_gloo(amrb):
    """phase A of the re-slicing (contiguous old ranges between Morton neighbours) over
    torch.distributed point-to-point operations, 3 processes, gloo"""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reshard_worker, args=(r, 3, port, "r2_s8_h1_d7_euler", q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) > 0, "the case must move patches between ranks"
