"""Morton-range sharding of the patch store across the GPUs of one box (no reference
counterpart — the reference is single-GPU, default stream).

One process per GPU.  Every rank derives the same global topology (leaf ids + neighbor tables),
owns a contiguous range of the Morton-sorted leaves, and appends *ghost slots* for the remote
patches its halos read.  Per step: `face_pack_kernel` gathers the min(2h, S)-thick interior slabs
remote halos need into one send buffer, `all_to_all_single` (NCCL over NVLink/NVSwitch, or gloo in
the CPU tests) moves them on a side stream while the fused kernel advances the patches whose
neighbors are all local, the received slabs are unpacked into the ghost slots, the boundary patches
are advanced, and the step's CFL minimum is all-reduced (min) — the only collective on the data path.

ShardPlan is pure numpy (testable without a GPU); ShardedSolver drives the C ABI.
"""
import os
import time

import numpy as np

from . import binding as B


class ShardPlan:
    """Who owns what, which ghost slots exist, and the pack / unpack entry lists."""

    def __init__(self, levels, rel, nbr, quad, rank, world, slab_layers_ok=True, halo_layers=None):
        P = len(levels)
        self.rank, self.world, self.P = rank, world, P
        ndir, kf = nbr.shape[1], nbr.shape[2]
        self.bounds = np.array([(r * P) // world for r in range(world + 1)], dtype=np.int64)
        lo, hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.lo, self.hi, self.n_owned = lo, hi, hi - lo
        owner = np.searchsorted(self.bounds[1:], np.arange(P), side="right").astype(np.int32)
        self.owner = owner

        sub = nbr[lo:hi]
        remote = (sub >= 0) & ((sub < lo) | (sub >= hi))
        ghosts = np.unique(sub[remote])
        self.ghost_global = ghosts
        self.n_total = self.n_owned + len(ghosts)
        loc = np.full(P, -1, dtype=np.int64)
        loc[lo:hi] = np.arange(self.n_owned)
        loc[ghosts] = self.n_owned + np.arange(len(ghosts))
        self.loc = loc
        self.levels = np.ascontiguousarray(levels[lo:hi], np.int32)
        self.rel = np.ascontiguousarray(rel[lo:hi], np.int8)
        self.quad = np.ascontiguousarray(quad[lo:hi], np.int8)
        self.nbr = np.where(sub >= 0, loc[np.maximum(sub, 0)], -1).astype(np.int32)
        has_remote = remote.reshape(self.n_owned, -1).any(axis=1)
        self.interior = np.nonzero(~has_remote)[0].astype(np.int32)
        self.boundary = np.nonzero(has_remote)[0].astype(np.int32)

        # every (reader i, direction d, source j) pair that crosses a rank boundary needs the slab
        # of j next to j's face d^1.  Both sides sort by (peer, j, face) -> identical order.
        I, D, K = np.nonzero(nbr >= 0)
        J = nbr[I, D, K].astype(np.int64)
        cross = owner[I] != owner[J]
        I, D, J = I[cross], D[cross], J[cross]
        ent, inv = np.unique(np.stack([owner[J].astype(np.int64), owner[I].astype(np.int64), J, D ^ 1],
                                      axis=1), axis=0, return_inverse=True)   # src rank, dst rank, patch, face
        # layers on the wire (entry word = face | layers << 4): all min(2h, S) of them only when a COARSER patch
        # reads the face (restriction of 2h fine layers); same-level copies and injection into finer patches
        # look at the first h layers
        lv = np.asarray(levels)
        coarser_reader = np.zeros(len(ent), dtype=bool)
        np.logical_or.at(coarser_reader, inv.reshape(-1), lv[I] < lv[J])
        if slab_layers_ok and halo_layers is not None:
            ent[:, 3] |= np.where(coarser_reader, 0, halo_layers).astype(np.int64) << 4
        snd = ent[ent[:, 0] == rank]
        snd = snd[np.lexsort((snd[:, 3], snd[:, 2], snd[:, 1]))]
        rcv = ent[ent[:, 1] == rank]
        rcv = rcv[np.lexsort((rcv[:, 3], rcv[:, 2], rcv[:, 0]))]
        self.send_entries = np.ascontiguousarray(np.stack([loc[snd[:, 2]], snd[:, 3]], axis=1), np.int32)
        self.recv_entries = np.ascontiguousarray(np.stack([loc[rcv[:, 2]], rcv[:, 3]], axis=1), np.int32)
        self.send_counts = np.bincount(snd[:, 1], minlength=world).astype(np.int64)
        self.recv_counts = np.bincount(rcv[:, 0], minlength=world).astype(np.int64)
        self.send_global = snd[:, 2:4].copy()
        self.recv_global = rcv[:, 2:4].copy()
        # peer-memory exchange: where this rank's segment starts inside rank q's receive buffer
        # (= the entries ranks below this one send to q; receive buffers are ordered by source rank)
        pair = np.zeros((world, world), dtype=np.int64)
        np.add.at(pair, (ent[:, 0], ent[:, 1]), 1)
        self.send_offsets = np.array([pair[:rank, q].sum() for q in range(world)], dtype=np.int64)
        assert (self.send_entries[:, 0] >= 0).all() and (self.send_entries[:, 0] < self.n_owned).all()
        assert (self.recv_entries[:, 0] >= self.n_owned).all()


class ReshardPlan:
    """Re-slicing after a reconstruct.  `kind/src/child` is the transfer plan of
    amrb_tree_reconstruct over the NEW global leaf order (0 copy of old leaf src, 1 prolongation from
    old leaf src, 2 restriction of old leaves src .. src+fan-1); old and new order are both Morton
    order, so the old leaves a rank needs for its new range form ONE contiguous old range [a, b).
    Phase A moves whole old patches so that rank q holds [a_q, b_q) in slots 0.. of its next buffer
    (contiguous range copies between ranks that are neighbours on the curve); phase B applies the
    rank's slice of the plan, re-indexed to a_q, with the single-GPU plan kernel."""

    def __init__(self, old_bounds, new_size, kind, src, child, fan, world):
        kind, src, child = np.asarray(kind), np.asarray(src, np.int64), np.asarray(child)
        self.world = world
        self.old_bounds = np.asarray(old_bounds, np.int64)
        self.new_bounds = np.array([(r * new_size) // world for r in range(world + 1)], dtype=np.int64)
        end = src + np.where(kind == 2, fan, 1)
        assert (np.diff(src) >= 0).all(), "the transfer plan is monotone in Morton order"
        self.need = []          # [a_q, b_q) per rank
        self.sub = []           # (kind, src - a_q, child) per rank
        for q in range(world):
            lo, hi = int(self.new_bounds[q]), int(self.new_bounds[q + 1])
            if hi > lo:
                a, b = int(src[lo:hi].min()), int(end[lo:hi].max())
            else:
                a = b = 0
            self.need.append((a, b))
            self.sub.append((np.ascontiguousarray(kind[lo:hi], np.int8),
                             np.ascontiguousarray(src[lo:hi] - a, np.int32),
                             np.ascontiguousarray(child[lo:hi], np.int8)))
        # moves[(r, q)] = (first old global index, count): old owner r -> new holder q
        self.moves = {}
        for q, (a, b) in enumerate(self.need):
            for r in range(world):
                lo, hi = max(int(self.old_bounds[r]), a), min(int(self.old_bounds[r + 1]), b)
                if hi > lo:
                    self.moves[(r, q)] = (lo, hi - lo)

    def staging_slots(self, q):
        return self.need[q][1] - self.need[q][0]


def migrate_old_patches(rp, rank, flat, cur, nxt, dist):
    """Phase A of the re-slicing over torch.distributed point-to-point operations (NCCL between GPUs,
    gloo in the CPU test): cur[f] / nxt[f] are flat views of this rank's current / next field arrays
    (device or host tensors); the old patches [a, b) this rank needs end up in nxt[f][0 : (b-a)*flat].
    Both sides walk the moves in the same (sender, receiver, field) order."""
    a, lo = rp.need[rank][0], int(rp.old_bounds[rank])
    ops = []
    for (r, q), (first, count) in sorted(rp.moves.items()):
        n = count * flat
        so, do = (first - lo) * flat, (first - a) * flat
        if r == rank and q == rank:
            for f in range(len(cur)):
                nxt[f][do:do + n].copy_(cur[f][so:so + n])
        elif r == rank:
            ops += [dist.P2POp(dist.isend, cur[f][so:so + n], q) for f in range(len(cur))]
        elif q == rank:
            ops += [dist.P2POp(dist.irecv, nxt[f][do:do + n], r) for f in range(len(cur))]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def raw_tensor(ptr, n, torch, dtype="<f8"):
    """torch view of device memory owned by the library (e.g. the dt-min slots)."""

    class _Raw:
        pass

    r = _Raw()
    r.__cuda_array_interface__ = {"shape": (n,), "typestr": dtype, "data": (int(ptr), False),
                                  "version": 2}
    return torch.as_tensor(r, device="cuda")


class ShardedSolver:
    """One rank's share of the mesh on one GPU."""

    def __init__(self, cfg, host_tree, rank, world, device, dist, torch, capacity=None,
                 storage=B.STORAGE_PADDED, transport="nccl"):
        """transport: "nccl" = pack -> all_to_all_single -> unpack + all_reduce(min), driven from here;
        "p2p" = the library's peer-memory exchange (amrb_exchange_*: slabs stored straight into the peers'
        receive buffers over NVLink, flags + CFL minimum in peer mailboxes, K-step loop in C++)."""
        self.cfg, self.rank, self.world, self.dist, self.torch = cfg, rank, world, dist, torch
        self.device = device
        self.transport = transport
        self.ex = None
        self._peer_maps = []
        levels, rel, nbr, quad = host_tree.tables()
        self.plan = pl = ShardPlan(levels, rel, nbr, quad, rank, world, halo_layers=cfg.halo)
        self.ids = host_tree.ids()[pl.lo:pl.hi]
        self.lay = B.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage)
        # capacity: slots for owned + ghost patches; meshes that change need headroom
        self.capacity = max(pl.n_total, 1) if capacity is None else int(capacity)
        self.pool = B.DevicePool(self.lay, self.capacity, device)
        self.pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
        self.pool.set_topology(pl.levels, pl.rel, pl.nbr, pl.quad, n_total=pl.n_total)
        self.L = B.lib()
        self.stream = torch.cuda.ExternalStream(int(self.L.amrb_pool_stream(self.pool.h) or 0),
                                                device=device)
        self.comm_stream = torch.cuda.Stream(device=device)
        # the CFL all-reduce runs on its own stream and its own communicator, so that it overlaps
        # the next step's slab exchange instead of sitting between two steps
        self.ar_stream = torch.cuda.Stream(device=device)
        self.ar_group = None
        if (world > 1 and dist is not None and dist.is_initialized() and dist.get_backend() == "nccl"
                and os.environ.get("AMRB_AR_GROUP", "1") != "0"):
            self.ar_group = dist.new_group(ranks=list(range(world)))
        self.graphs = {}
        self._dtmin_cache = {}
        self._dtmin_base = 0
        self.launches = 0
        self._install_exchange()

    def _install_exchange(self):
        """device copies of the plan's entry lists and the slab buffers (after every new ShardPlan)"""
        torch, pl, cfg, device = self.torch, self.plan, self.cfg, self.device
        slab = self.L.amrb_pool_face_slab_doubles(self.pool.h, 0) * cfg.nvar
        self.slab = slab
        dev = torch.device("cuda", device)
        self.d_send_entries = torch.from_numpy(pl.send_entries.reshape(-1).copy()).to(dev)
        self.d_recv_entries = torch.from_numpy(pl.recv_entries.reshape(-1).copy()).to(dev)
        self.d_interior = torch.from_numpy(pl.interior.copy()).to(dev)
        self.d_boundary = torch.from_numpy(pl.boundary.copy()).to(dev)
        self.send_buf = torch.zeros(max(len(pl.send_entries) * slab, 1), dtype=torch.float64, device=dev)
        self.recv_buf = torch.zeros(max(len(pl.recv_entries) * slab, 1), dtype=torch.float64, device=dev)
        self.in_splits = [int(c * slab) for c in pl.send_counts]
        self.out_splits = [int(c * slab) for c in pl.recv_counts]
        self.exchanged_bytes = (sum(self.in_splits) + sum(self.out_splits)) * 8
        self.graphs = {}
        self._dtmin_cache = {}
        if self.transport == "p2p":
            self._install_p2p()

    # ---- peer-memory exchange (include/gpuamr_b200.h section 6b)
    def _install_p2p(self):
        import ctypes as C
        L, pl = self.L, self.plan
        self._drop_p2p()
        ex = C.c_void_p()
        sc = np.ascontiguousarray(pl.send_counts, np.int64)
        so = np.ascontiguousarray(pl.send_offsets, np.int64)
        se = np.ascontiguousarray(pl.send_entries, np.int32)
        re = np.ascontiguousarray(pl.recv_entries, np.int32)
        B.check(L.amrb_exchange_create(self.pool.h, self.rank, self.world, B._ptr(se), B._ptr(sc), B._ptr(so),
                                       B._ptr(re), len(re), C.byref(ex)))
        self.ex = ex
        bd = np.ascontiguousarray(pl.boundary, np.int32)
        it = np.ascontiguousarray(pl.interior, np.int32)
        B.check(L.amrb_exchange_set_lists(ex, B._ptr(bd), len(bd), B._ptr(it), len(it)))
        if self.dist is None:
            return                       # in-process cluster: LocalCluster connects the raw pointers
        handles = []
        for which in range(3):
            buf = (C.c_ubyte * 64)()
            B.check(L.amrb_ipc_export(L.amrb_exchange_buffer(ex, which), buf))
            handles.append(bytes(buf))
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, handles)
        err = None
        try:
            for r in range(self.world):
                if r == self.rank:
                    continue
                ptrs = []
                for hb in gathered[r]:
                    out = C.c_void_p()
                    src = (C.c_ubyte * 64).from_buffer_copy(hb)
                    B.check(L.amrb_ipc_open(src, C.byref(out)))
                    ptrs.append(out)
                    self._peer_maps.append(out)
                B.check(L.amrb_exchange_connect(ex, r, ptrs[0], ptrs[1], ptrs[2]))
        except B.AmrbError as e:         # e.g. no peer access between two GPUs of this box
            err = str(e)
        oks = [None] * self.world
        self.dist.all_gather_object(oks, err)
        if any(o is not None for o in oks):
            # every rank falls back together: the exchange driven from here over torch.distributed
            import sys
            if self.rank == 0:
                sys.stderr.write("peer-memory exchange unavailable (%s): NCCL transport instead\n"
                                 % next(o for o in oks if o is not None))
            self._drop_p2p()
            self.transport = "nccl"

    def _drop_p2p(self):
        if self.ex is not None:
            self.torch.cuda.synchronize()
            if self.dist is not None:
                self.dist.barrier()      # nobody may still be pushing into buffers about to be unmapped
            for m in self._peer_maps:
                self.L.amrb_ipc_close(m)
            self._peer_maps = []
            self.L.amrb_exchange_destroy(self.ex)
            self.ex = None

    def close(self):
        self._drop_p2p()
        self.pool.close()

    # ---- re-slicing after a reconstruct
    def field_views(self, which):
        """torch views of the whole current ('cur') / next ('nxt') field arrays of the pool"""
        get = self.L.amrb_pool_field if which == "cur" else self.L.amrb_pool_next_field
        n = self.capacity * self.pool.stored
        return [raw_tensor(get(self.pool.h, f), n, self.torch) for f in range(self.cfg.nvar)]

    def reshard(self, new_host_tree, old_size, plan):
        """re-slicing after new_host_tree.reconstruct() (every rank ran it on the same global flags):
        phase A over NCCL point-to-point copies, phase B locally.  The two phases are tested
        separately (migrate_old_patches over gloo on the CPU, finish_reshard through LocalCluster on
        one GPU); their composition has not been run on several GPUs yet."""
        kind, src, child = plan
        old_bounds = [(r * old_size) // self.world for r in range(self.world + 1)]
        rp = ReshardPlan(old_bounds, new_host_tree.size, kind, src, child, 1 << self.cfg.rank, self.world)
        if rp.staging_slots(self.rank) > self.capacity:   # before phase A writes into the next buffer
            raise B.AmrbError("re-slicing needs %d staging slots, pool capacity is %d"
                              % (rp.staging_slots(self.rank), self.capacity))
        self.halo_exchange()                    # copied patches carry their halos along
        self.torch.cuda.synchronize()
        migrate_old_patches(rp, self.rank, self.pool.stored, self.field_views("cur"),
                            self.field_views("nxt"), self.dist)
        self.torch.cuda.synchronize()
        self.finish_reshard(rp, new_host_tree)
        return rp

    def finish_reshard(self, rp, new_host_tree):
        """phase B: the incoming old patches [a, b) sit in the next buffer -> make them current, apply
        this rank's slice of the transfer plan, install the new ShardPlan (tables, ghost slots)"""
        B.check(self.L.amrb_pool_swap_buffers(self.pool.h))
        kind, src, child = rp.sub[self.rank]
        if len(kind):
            self.pool.apply_plan(kind, src, child)
        levels, rel, nbr, quad = new_host_tree.tables()
        self.plan = pl = ShardPlan(levels, rel, nbr, quad, self.rank, self.world, halo_layers=self.cfg.halo)
        if pl.n_total > self.capacity or rp.staging_slots(self.rank) > self.capacity:
            raise B.AmrbError("shard of %d patches (+ghosts) exceeds the pool capacity %d"
                              % (pl.n_total, self.capacity))
        self.ids = new_host_tree.ids()[pl.lo:pl.hi]
        self.pool.set_topology(pl.levels, pl.rel, pl.nbr, pl.quad, n_total=pl.n_total)
        self._install_exchange()

    # ---- data
    def upload_interior(self, data):
        for f in range(self.cfg.nvar):
            self.pool.upload_interior(f, data[f])

    def download_interior(self):
        cfg, n = self.cfg, self.plan.n_owned
        return np.stack([self.pool.download_interior(f, n).reshape((n,) + (cfg.size,) * cfg.rank)
                         for f in range(cfg.nvar)])

    # ---- exchange
    def _pack(self):
        n = len(self.plan.send_entries)
        if n:
            B.check(self.L.amrb_pool_pack_faces(self.pool.h, self.d_send_entries.data_ptr(), n,
                                                self.send_buf.data_ptr()))
            self.launches += 1

    def _unpack(self):
        n = len(self.plan.recv_entries)
        if n:
            B.check(self.L.amrb_pool_unpack_faces(self.pool.h, self.d_recv_entries.data_ptr(), n,
                                                  self.recv_buf.data_ptr()))
            self.launches += 1

    def _comm(self):
        """slabs -> peers on the side stream; returns after enqueueing (device-side ordering only)"""
        torch, dist = self.torch, self.dist
        self.comm_stream.wait_stream(self.stream)
        with torch.cuda.stream(self.comm_stream):
            dist.all_to_all_single(self.recv_buf[:sum(self.out_splits)], self.send_buf[:sum(self.in_splits)],
                                   self.out_splits, self.in_splits)

    def exchange(self):
        if self.ex is not None:
            B.check(self.L.amrb_exchange_halo(self.ex))
            self.launches += 3
            return
        self._pack()
        self._comm()
        self.stream.wait_stream(self.comm_stream)
        self._unpack()

    def halo_exchange(self):
        self.exchange()
        self.pool.halo_exchange()
        self.launches += 1

    # ---- stepping
    def _dtmin_tensor(self, k):
        """torch view of the dt-min slot entering step k.  The library re-allocates its scalar arrays when
        a batch needs more slots than any batch before (amrb_pool_batch_begin), so the views are keyed by
        the slot's CURRENT device address, never by k alone; recorded graphs hold the old addresses and
        are dropped with them."""
        ptr = int(self.L.amrb_pool_dtmin_slot(self.pool.h, k) or 0)
        if ptr == 0:
            raise B.AmrbError("no dt-min slot %d in the open batch" % k)
        if k == 0 and ptr != self._dtmin_base:
            self._dtmin_base = ptr
            self._dtmin_cache = {}
            self.graphs = {}
        t = self._dtmin_cache.get(ptr)
        if t is None:
            t = raw_tensor(ptr, 1, self.torch)
            self._dtmin_cache[ptr] = t
        return t

    def _allreduce_dtmin(self, k):
        """global CFL minimum of the state entering step k (one double), on the side stream;
        the step kernels of step k wait for it, the slab exchange of step k does not"""
        torch, dist = self.torch, self.dist
        t = self._dtmin_tensor(k)
        self.ar_stream.wait_stream(self.stream)
        with torch.cuda.stream(self.ar_stream):
            if self.ar_group is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.ar_group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)

    def advance_batch_async(self, steps, remaining=B.DBL_MAX, overlap=True):
        L, h, pl = self.L, self.pool.h, self.plan
        if self.ex is not None:
            l0 = int(L.amrb_exchange_launch_count(self.ex)) + self.pool.launch_count()
            B.check(L.amrb_exchange_advance_batch_async(self.ex, steps, remaining, 1 if overlap else 0))
            self.launches += int(L.amrb_exchange_launch_count(self.ex)) + self.pool.launch_count() - l0
            return
        B.check(L.amrb_pool_batch_begin(h, steps, remaining))
        self._allreduce_dtmin(0)
        for k in range(steps):
            self._pack()
            self._comm()
            self.stream.wait_stream(self.ar_stream)          # dt of this step is global now
            if overlap and len(pl.interior):
                B.check(L.amrb_pool_step_partial(h, self.d_interior.data_ptr(), len(pl.interior)))
                self.launches += 1
            self.stream.wait_stream(self.comm_stream)
            self._unpack()
            if overlap:
                if len(pl.boundary):
                    B.check(L.amrb_pool_step_partial(h, self.d_boundary.data_ptr(), len(pl.boundary)))
                    self.launches += 1
            else:
                B.check(L.amrb_pool_step_partial(h, None, 0))
                self.launches += 1
            self._allreduce_dtmin(k + 1)
            B.check(L.amrb_pool_step_commit(h))
        self.stream.wait_stream(self.ar_stream)
        self.exchange()
        B.check(L.amrb_pool_batch_end(h, 1))
        self.launches += 1

    def advance_batch_graph(self, steps, overlap=True):
        """Same batch, recorded once into a CUDA graph (kernels, NCCL exchanges and the event
        fork/joins between the three streams) and replayed: no per-launch host work inside the
        batch.  Needs an even step count (the current/next buffer pointers of the recorded
        launches must be back in place after one replay) and a preceding eager batch of the same
        length (allocations, carried dt-min slot)."""
        torch = self.torch
        if steps % 2 != 0:
            raise ValueError("graph replay needs an even number of steps")
        g = self.graphs.get((steps, overlap))
        if g is None:
            self.finish_advance_batch()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = self.launches
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                self.advance_batch_async(steps, overlap=overlap)
            self.graph_launches = self.launches - l0
            self.launches = l0
            self.finish_advance_batch()                      # host state only: nothing ran yet
            self.graphs[(steps, overlap)] = g
        with torch.cuda.stream(self.stream):
            g.replay()
        self.launches += self.graph_launches

    def finish_advance_batch(self, max_steps=0):
        out = self.pool.finish_advance_batch(max_steps)
        if self.ex is not None and self.L.amrb_exchange_timed_out(self.ex):
            raise B.AmrbError("peer-memory exchange: a rank stopped answering (receiver timed out)")
        return out


# ------------------------------------------------------------------------------- in-process cluster
class LocalCluster:
    """All W shards of one mesh inside ONE process on ONE GPU, driven phase by phase: the same
    ShardPlan, ghost slots, pack / unpack kernels, interior / boundary launches and per-step CFL
    minimum as the NCCL path, with the slab exchange and the all-reduce done by device-to-device
    copies.  Test vehicle for the sharding logic on a single-GPU box (the NCCL transport itself is
    covered by tests/test_multigpu_gpu.py on >= 2 GPUs)."""

    def __init__(self, cfg, host_tree, world, device, torch, capacity=None, storage=B.STORAGE_PADDED,
                 transport="copy"):
        """transport "copy": slabs moved by device-to-device copies of the packed buffers (the NCCL path's
        stand-in); "p2p": the library's peer-memory exchange with the peers' buffers connected as plain
        pointers (same device), driven push-all / wait-all because all ranks share this process."""
        self.torch, self.world, self.cfg = torch, world, cfg
        self.p2p = transport == "p2p"
        self.sols = [ShardedSolver(cfg, host_tree, r, world, device, None, torch, capacity=capacity,
                                   storage=storage, transport="p2p" if self.p2p else "nccl")
                     for r in range(world)]
        self._index_exchange()

    def _index_exchange(self):
        # where rank r's send segment for rank q starts / where q expects r's data
        self.send_off = [np.concatenate([[0], np.cumsum(s.in_splits)]).astype(int) for s in self.sols]
        self.recv_off = [np.concatenate([[0], np.cumsum(s.out_splits)]).astype(int) for s in self.sols]
        for r, s in enumerate(self.sols):
            for q, t in enumerate(self.sols):
                assert s.in_splits[q] == t.out_splits[r], "send / receive plans disagree"
        if self.p2p:
            L = self.sols[0].L
            for r, s in enumerate(self.sols):
                for q, t in enumerate(self.sols):
                    if q != r:
                        B.check(L.amrb_exchange_connect(s.ex, q, L.amrb_exchange_buffer(t.ex, 0),
                                                        L.amrb_exchange_buffer(t.ex, 1),
                                                        L.amrb_exchange_buffer(t.ex, 2)))

    def _sync(self):
        self.torch.cuda.synchronize()

    def _p2p_exchange(self, with_dt=0, k=0):
        L = self.sols[0].L
        for s in self.sols:                                  # every push is enqueued before any wait
            B.check(L.amrb_exchange_push(s.ex, with_dt, k))
        for s in self.sols:
            B.check(L.amrb_exchange_wait(s.ex, with_dt, k))

    def exchange(self):
        if self.p2p:
            self._p2p_exchange()
            self._sync()
            return
        for s in self.sols:
            s._pack()
        self._sync()
        for r, s in enumerate(self.sols):
            for q, t in enumerate(self.sols):
                n = s.in_splits[q]
                if n:
                    t.recv_buf[self.recv_off[q][r]:self.recv_off[q][r] + n].copy_(
                        s.send_buf[self.send_off[r][q]:self.send_off[r][q] + n])
        self._sync()
        for s in self.sols:
            s._unpack()
        self._sync()

    def halo_exchange(self):
        self.exchange()
        for s in self.sols:
            s.pool.halo_exchange()
        self._sync()

    def _reduce_dtmin(self, k):
        ts = [s._dtmin_tensor(k) for s in self.sols]
        m = self.torch.stack([t.reshape(()) for t in ts]).min()
        for t in ts:
            t.fill_(m)
        self._sync()

    def advance_batch(self, steps, overlap=True):
        """same launch sequence per rank as ShardedSolver.advance_batch_async"""
        L = self.sols[0].L
        for s in self.sols:
            B.check(L.amrb_pool_batch_begin(s.pool.h, steps, B.DBL_MAX))
        self._sync()
        if not self.p2p:
            self._reduce_dtmin(0)
        for k in range(steps):
            if self.p2p:
                self._p2p_exchange(1, k)                     # slabs + CFL minimum of slot k in one exchange
            else:
                self.exchange()
            for s in self.sols:
                pl, h = s.plan, s.pool.h
                if overlap:
                    if len(pl.interior):
                        B.check(L.amrb_pool_step_partial(h, s.d_interior.data_ptr(), len(pl.interior)))
                    if len(pl.boundary):
                        B.check(L.amrb_pool_step_partial(h, s.d_boundary.data_ptr(), len(pl.boundary)))
                else:
                    B.check(L.amrb_pool_step_partial(h, None, 0))
            self._sync()
            if not self.p2p:
                self._reduce_dtmin(k + 1)
            for s in self.sols:
                B.check(L.amrb_pool_step_commit(s.pool.h))
        self.exchange()
        for s in self.sols:
            B.check(L.amrb_pool_batch_end(s.pool.h, 1))
        self._sync()
        return [s.finish_advance_batch(steps) for s in self.sols]

    def patch_max_flags(self, field, refine_thr, coarsen_thr, min_level, max_level):
        """the criterion on every shard, concatenated in global (Morton) order"""
        out = [s.pool.patch_max_flags(field, refine_thr, coarsen_thr, min_level, max_level)[:s.plan.n_owned]
               for s in self.sols]
        return np.concatenate(out)

    def reshard(self, host_tree, old_size, plan):
        """after host_tree.reconstruct(): move old patches to the ranks that need them (phase A, here
        device-to-device copies; the NCCL form sends the same contiguous ranges), then every shard
        applies its slice of the plan and installs its new tables (phase B)"""
        kind, src, child = plan
        flat = self.sols[0].pool.stored
        old_bounds = [s.plan.lo for s in self.sols] + [old_size]
        rp = ReshardPlan(old_bounds, host_tree.size, kind, src, child, 1 << self.cfg.rank, self.world)
        self.halo_exchange()                    # copied patches carry their halos along
        cur = [s.field_views("cur") for s in self.sols]
        nxt = [s.field_views("nxt") for s in self.sols]
        for (r, q), (first, count) in rp.moves.items():
            so = (first - int(old_bounds[r])) * flat
            do = (first - rp.need[q][0]) * flat
            for f in range(self.cfg.nvar):
                nxt[q][f][do:do + count * flat].copy_(cur[r][f][so:so + count * flat])
        self._sync()
        for s in self.sols:
            s.finish_reshard(rp, host_tree)
        self._index_exchange()
        self._sync()
        return rp

    def close(self):
        for s in self.sols:
            s.close()


# ---------------------------------------------------------------------------------------- bench
def weak_scaled_tree(wl, world):
    """C2-family mesh with ~world x 2272 patches of 64x64 Euler cells: base level 5 + floor(log4 N),
    two refinement rings (r, r/2); r found by bisection on the patch count."""
    base = 5
    n = world
    while n >= 4:
        base += 1
        n //= 4
    cfg = wl.Config(2, 64, 1, 9, B.EQ_EULER)
    target = 2272 * world

    def build(r):
        return wl.build_static_tree(cfg, base, (r, r / 2.0))

    lo, hi = 0.05, 0.75
    best = None
    for _ in range(18):
        mid = 0.5 * (lo + hi)
        t = build(mid)
        if best is None or abs(t.size - target) < abs(best[1].size - target):
            best = (mid, t)
        if t.size < target:
            lo = mid
        else:
            hi = mid
    return cfg, best[1], base, best[0]


def state_checksums(torch, views, bounds, stored):
    """order-independent integer checksums of the bit patterns of each field over each patch range
    [bounds[r], bounds[r+1]): int64 sums with wrap-around; bit-identical states <-> identical sums"""
    out = torch.zeros((len(views), len(bounds) - 1), dtype=torch.int64, device=views[0].device)
    for f, v in enumerate(views):
        iv = v.view(torch.int64)
        for r in range(len(bounds) - 1):
            out[f, r] = iv[int(bounds[r]) * stored:int(bounds[r + 1]) * stored].sum()
    return out


def check_parity(torch, dist, bench_mod, amrb, wl, cfg, host, sol, rank, world, local, storage, steps=3,
                 tol=1e-12):
    """Correctness evidence carried by the bench line itself: the sharded run (slab exchange, ghost slots,
    per-step CFL minimum across ranks) against ONE pool holding the whole mesh on rank 0, same initial
    condition, `steps` steps.  Rank 0 sends every rank the reference state of its Morton range; each rank
    compares element by element: max |a - b| / max |b| per field (SURVEY 8c metric) must be <= tol, the dt
    sequence must agree to tol.  Also reports whether the states are bit-identical (they are whenever the
    shards and the single pool launch the same kernel instantiation; the 2D marching kernel picks its band
    height from the launch size, which changes rounding in the last bit)."""
    ids_all = host.ids()
    P, stored, nvar = len(ids_all), sol.pool.stored, cfg.nvar
    bench_mod.fill_ic(torch, amrb, wl, sol.pool, sol.ids, cfg, local)
    sol.halo_exchange()
    sol.advance_batch_async(steps)
    _, n_s, dts_s = sol.finish_advance_batch(steps)
    torch.cuda.synchronize()
    n_own = sol.plan.n_owned
    own = [v[:n_own * stored] for v in sol.field_views("cur")]
    bounds = [int(b) for b in sol.plan.bounds]
    detail, views, pool = {}, None, None
    dts_1 = None
    if rank == 0:
        lay = B.make_layout(cfg.rank, cfg.size, cfg.halo, cfg.eq, cfg.depth, storage)
        pool = B.DevicePool(lay, P, local)
        pool.set_physics([cfg.length] * 3, cfg.gamma, cfg.cfl)
        pool.set_topology_from_ids(ids_all)
        bench_mod.fill_ic(torch, amrb, wl, pool, ids_all, cfg, local)
        pool.halo_exchange()
        pool.advance_batch_async(steps)
        _, n_1, dts_1 = pool.finish_advance_batch(steps)
        torch.cuda.synchronize()
        views = bench_mod.pool_field_views(torch, amrb, pool, P * stored, nvar)
    # element-wise comparison, one field of one rank's range at a time (chunks of <= 2^27 doubles)
    CH = 1 << 27
    stats = torch.zeros(3 * nvar, dtype=torch.float64, device="cuda")     # per field: max|a-b|, max|b|, #different
    buf = torch.empty(min(CH, max(n_own * stored, 1)), dtype=torch.float64, device="cuda") if rank != 0 else None
    for r in range(world):
        lo, hi = bounds[r] * stored, bounds[r + 1] * stored
        for f in range(nvar):
            for c0 in range(lo, hi, CH):
                c1 = min(hi, c0 + CH)
                if rank == 0 and r == 0:
                    ref, mine = views[f][c0:c1], own[f][c0 - lo:c1 - lo]
                elif rank == 0:
                    dist.send(views[f][c0:c1].contiguous(), dst=r)
                    continue
                elif rank == r:
                    ref = buf[:c1 - c0]
                    dist.recv(ref, src=0)
                    mine = own[f][c0 - lo:c1 - lo]
                else:
                    continue
                d = (mine - ref).abs()
                stats[3 * f] = torch.maximum(stats[3 * f], d.max())
                stats[3 * f + 1] = torch.maximum(stats[3 * f + 1], ref.abs().max())
                stats[3 * f + 2] += (mine.view(torch.int64) != ref.view(torch.int64)).sum().to(torch.float64)
    mx = stats.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(stats)                                   # the #different entries add up
    ok = True
    if rank == 0:
        err = max(float(mx[3 * f]) / max(float(mx[3 * f + 1]), 1e-300) for f in range(nvar))
        ndiff = int(sum(float(stats[3 * f + 2]) for f in range(nvar)))
        a, b = np.asarray(dts_s), np.asarray(dts_1)
        dt_ok = bool(n_1 == n_s and len(a) == len(b) and np.allclose(a, b, rtol=tol, atol=0))
        ok = bool(err <= tol) and dt_ok
        detail = {"checked": True, "steps": int(steps), "against": "one pool holding the whole mesh on rank 0",
                  "state_max_err_over_field_max": err, "tolerance": tol, "state_bit_identical": ndiff == 0,
                  "values_differing_in_the_last_bits": ndiff, "dt_sequence_matches": dt_ok,
                  "dt_sequence_bit_identical": bool(np.array_equal(a, b))}
        pool.close()
        del views
        torch.cuda.empty_cache()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(int(flag.item())), detail


def run_bench(args, METRIC, UNIT):
    """N > 1.  Default: STRONG scaling of the C3 mesh (the same ~1.04e9-cell 3D tree at every N, one Morton
    range per GPU, interior-only pools).  --workload c2: the weak-scaled C2 family of round 1."""
    import importlib
    import json
    import sys

    import torch
    import torch.distributed as dist

    from . import workloads as wl

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench as bench_mod
    amrb = importlib.import_module("gpu-amr_b200")

    strong = args.workload != "c2"
    if strong:
        cfg = wl.c3_config()
        base = wl.C3["base_level"] if args.workload == "c3" else int(args.workload.split("L")[-1])
        host = wl.build_static_tree(cfg, base, wl.C3["ball_radii"])
        storage, radius = B.STORAGE_INTERIOR, None
    else:
        cfg, host, base, radius = weak_scaled_tree(wl, world)
        storage = B.STORAGE_PADDED
    sol = ShardedSolver(cfg, host, rank, world, local, dist, torch, storage=storage,
                        transport=os.environ.get("AMRB_TRANSPORT", "p2p"))
    P = host.size
    cells = P * cfg.data
    parity_ok, parity = check_parity(torch, dist, bench_mod, amrb, wl, cfg, host, sol, rank, world, local, storage)
    if not parity_ok:
        if rank == 0:
            sys.stderr.write("PARITY FAILURE (sharded vs single pool): %s\n" % json.dumps(parity))
        dist.barrier()
        sol.close()
        dist.destroy_process_group()
        raise SystemExit(3)
    bench_mod.fill_ic(torch, amrb, wl, sol.pool, sol.ids, cfg, local)
    sol.halo_exchange()
    K, W = args.steps, max(args.warmup, 3)
    sol.advance_batch_async(W)
    sol.finish_advance_batch()
    sol.advance_batch_async(K)
    sol.finish_advance_batch()
    # exchange schedule: (a) interior patches advance while the slabs travel, boundary patches after
    # the receive (two step launches), or (b) exchange first, then one step launch over all patches.
    # (a) wins when a rank's share is large, (b) when it is small (a launch over few boundary
    # patches still costs a full pipeline fill); picked here by timing a short batch of each.
    mode_ms = {}
    for ov in (True, False):
        for rep in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sol.stream)
            sol.advance_batch_async(6, overlap=ov)
            e1.record(sol.stream)
            torch.cuda.synchronize()
            sol.finish_advance_batch()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mode_ms[ov] = float(t.item())
    overlap = mode_ms[True] <= mode_ms[False]
    if os.environ.get("AMRB_OVERLAP"):
        overlap = os.environ["AMRB_OVERLAP"] != "0"
    advance = lambda k: sol.advance_batch_async(k, overlap=overlap)  # noqa: E731

    clocks = bench_mod.ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    reps, total = [], 0.0
    t0w = time.time()
    while len(reps) < 3 or (total < 500.0 and len(reps) < 400):
        l0 = sol.launches
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sol.stream)
        advance(K)
        e1.record(sol.stream)
        torch.cuda.synchronize()
        dt_sum, executed, _ = sol.finish_advance_batch()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # max over ranks
        reps.append(float(ms.item()))
        total += reps[-1]
        launches = sol.launches - l0
    t1w = time.time()
    ms_total = sorted(reps)[len(reps) // 2]
    value = cells * K / (ms_total * 1e-3)

    # per-phase device times of one more K-step batch (peer-memory transport, plain schedule), max over ranks
    phases = None
    if sol.ex is not None:
        import ctypes as C
        B.check(sol.L.amrb_exchange_set_timing(sol.ex, 1))
        dist.barrier()
        sol.advance_batch_async(K, overlap=False)
        sol.finish_advance_batch()
        out4 = (C.c_double * 4)()
        B.check(sol.L.amrb_exchange_get_timing(sol.ex, out4))
        B.check(sol.L.amrb_exchange_set_timing(sol.ex, 0))
        ph = torch.tensor(list(out4), dtype=torch.float64, device="cuda") / K
        pmax, pmin = ph.clone(), ph.clone()
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        phases = {n: {"max_over_ranks": float(pmax[i]), "min_over_ranks": float(pmin[i])}
                  for i, n in enumerate(("push", "wait", "unpack", "step_kernel"))}

    # e2e: pinned-host state -> device, halo, K steps, state back to the host (per rank its shard)
    n_own, stored = sol.plan.n_owned, sol.pool.stored
    fbytes = n_own * stored * 8
    bufs, how = bench_mod.host_state_buffers(torch, fbytes, cfg.nvar)
    L = sol.L
    for f in range(cfg.nvar):
        B.check(L.amrb_copy_device_to_host(bufs[f].data_ptr(), L.amrb_pool_field(sol.pool.h, f), fbytes))
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sol.stream)
    for f in range(cfg.nvar):
        B.check(L.amrb_copy_host_to_device_async(L.amrb_pool_field(sol.pool.h, f), bufs[f].data_ptr(),
                                                 fbytes, L.amrb_pool_stream(sol.pool.h)))
    sol.pool.mark_dirty()                                    # raw writes: drop the carried dt-min
    sol.halo_exchange()
    sol.advance_batch_async(K, overlap=overlap)
    for f in range(cfg.nvar):
        B.check(L.amrb_copy_device_to_host_async(bufs[f].data_ptr(), L.amrb_pool_field(sol.pool.h, f),
                                                 fbytes, L.amrb_pool_stream(sol.pool.h)))
    e1.record(sol.stream)
    torch.cuda.synchronize()
    sol.finish_advance_batch()
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    state_bytes = cfg.nvar * fbytes
    tot = torch.tensor([float(state_bytes), float(sol.exchanged_bytes // 2), float(len(sol.plan.ghost_global)),
                        float(len(sol.plan.boundary))], dtype=torch.float64, device="cuda")
    mx = tot.clone()
    dist.all_reduce(tot)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        clk = clocks.stop(t0w, t1w)
        peaks, peak_src = bench_mod.measured_peaks()
        b_alg = 2 * cfg.nvar * 8
        achieved = value * b_alg / 1e9 / world
        if strong and args.workload == "c3":
            cfgobj = bench_mod.workload_config(wl, "c3", world, cells)
        elif strong:
            cfgobj = {"workload": "development: C3 family at base level %d" % base, "cells": int(cells)}
        else:
            cfgobj = {"workload": "C2 family, weak-scaled: 2D static multi-level tree, Euler fp64, 64x64 "
                                  "patches halo 1, base level %d + 2 rings (r=%.4f L), Morton-range "
                                  "partition over %d GPUs" % (base, radius, world), "cells": int(cells),
                      "l2_policy": "inputs larger than L2 (0.8 GB of state per GPU vs 126 MB)"}
        cfgobj.update({"patches": int(P), "cells_per_gpu": int(cells // world), "executed_steps": int(executed),
                       "partition": "contiguous Morton ranges, equal patch counts, %d GPUs" % world,
                       "device_layout": "interior-only [P][S^3] per field" if storage else "padded",
                       "launch_mode": "eager launches",
                       "transport": "peer-memory push over NVLink (slabs + CFL minimum + flag in one kernel), "
                                    "K-step loop in the library" if sol.transport == "p2p" else
                                    "NCCL all_to_all_single + all_reduce(min) driven from Python",
                       "exchange_schedule": (("boundary patches first, slab push of the next step on a side stream under "
                                              "the interior launch" if overlap else
                                              "push, wait, unpack, one launch over all patches")
                                             if sol.transport == "p2p" else
                                             ("interior patches overlap the slab exchange" if overlap else
                                              "exchange, then one launch over all patches")),
                       "schedule_probe_ms_per_6_steps": {"overlap": mode_ms[True], "single_launch": mode_ms[False]},
                       "ms_per_step_by_phase": phases,
                       "ghost_patches_max_rank": int(mx[2].item()), "boundary_patches_max_rank": int(mx[3].item()),
                       "ghost_bytes_per_step_all_ranks": float(tot[1].item())})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfgobj,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                         "kernel": "whole step per GPU (slab push / exchange + wait + unpack + fused step kernel)",
                         "algorithmic_bytes_per_cell": b_alg, "peak_source": peak_src},
            "cpu_baseline": None, "parity": parity,
            "e2e": {"value": cells * K / (float(ems.item()) * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": float(tot[0].item()) / K, "d2h_bytes_per_step": float(tot[0].item()) / K + 8,
                    "ms_total": float(ems.item()), "host_buffers": how},
            "gpu_launches": int(launches), "clocks": clk, "batch_ms": bench_mod.spread(reps),
            "timed_region_s": total * 1e-3,
        }
        print(json.dumps(line))
    dist.barrier()
    sol.close()
    dist.destroy_process_group()
