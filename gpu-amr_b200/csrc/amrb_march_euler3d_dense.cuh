// Fused Euler step over the INTERIOR-ONLY 3D pool layout (amrb_layout.storage = AMRB_STORAGE_INTERIOR):
// field f of patch p is S^3 contiguous doubles, ghosts are never stored anywhere.
//
// Same warp-autonomous plane-marching pipeline as euler3d_march_kernel (amrb_march_euler3d.cuh; read its
// header first), re-laid for the dense layout.  What the layout changes:
//   * HBM traffic: 4 096 B in + 4 096 B out per 8^3 field-patch instead of 6 400 + 6 400 B of the
//     padded 10^3 layout (whose ceiling was 0.64 of the HBM roofline); a 1.07e9-cell mesh is 86 GB;
//   * the ring: CR whole interior planes of a field-patch are one contiguous run (CR * S*S doubles;
//     2 KB for four 8 x 8 planes) -> one TMA bulk copy per field per chunk, 5 per 10 KB stage; copies of
//     >= 2 KB keep a warp's serial bulk-copy completion rate out of the way (tools/tma_bench.cu);
//   * stores: the lane's (A | B) pair is 16-byte aligned in the dense row, a plane is 512 contiguous
//     bytes of fully written sectors: no ghost-row / ghost-column copies, no pairing shuffle;
//   * boundary faces: the interior cell of a lateral boundary face is read from the staged plane in
//     shared memory (the padded kernel fetched it from the pool again), only the ghost cell is gathered
//     from the neighbor patch interior through the halo tables.
// Ghost gathers follow halo_source<3, S, H, 0> (amrb_kernels.cuh) = the reference's same_t / finer_t /
// coarser_t operators (include/ndtree/patch_utils.hpp:303-441), restriction summed last-dim-fastest.
// Arithmetic: include/solver/EulerPhysics.hpp:74-129, amr_solver.hpp:265-353 (see the padded kernel).
#pragma once
#include "amrb_march_euler3d.cuh"

namespace amrb
{

template <int S, int CR, int NS, int WPC>
struct March3DenseCfg
{
    static constexpr int NV    = 5;
    static constexpr int SS    = S * S;
    static constexpr int N     = S * S * S;            // doubles per field-patch
    static constexpr int NBX   = S / 8, NBY = S / 8;
    static constexpr int NB    = NBX * NBY;            // tasks (8 x 8 column blocks) per patch
    static constexpr bool WHOLE = (NB == 1);           // stream whole planes with TMA bulk copies
    static constexpr int PLD   = 64;                   // doubles per staged field-plane: the block's 8 x 8 cells
    static constexpr int NCH   = S / CR;               // chunks per task
    static constexpr int FS    = CR * PLD;             // field stride inside a stage
    static constexpr int STAGE = NV * FS;
    static constexpr int RING  = NS * STAGE;
    static constexpr int BFW   = 6;                    // doubles per parked boundary flux (5 + pad)
    static constexpr int BF    = 2 * 32 * BFW;         // double-buffered, 32 lateral faces per plane
    static constexpr int ST    = 5 * 32;               // ghost cells of the next plane's faces in flight
    static constexpr int GZ    = 5 * 64;               // ghost plane below the NEXT task's first plane, in flight
    static constexpr int WARP_DOUBLES = RING + BF + ST + GZ;
    static constexpr size_t SMEM      = (size_t)WPC * WARP_DOUBLES * sizeof(double);
    static_assert(S % 8 == 0, "8 x 8 column blocks");
    static_assert(S % CR == 0, "chunk shape");
    static_assert(NS * CR <= S, "the ring never reaches beyond the next task");
    static_assert((FS * 8) % 16 == 0 && (WARP_DOUBLES * 8) % 16 == 0, "bulk copy alignment");
};

// OPT bits: 1 = two planes per loop trip (ping-pong plane states); 2 = task prologue pipelined into the previous
// task (ghost cells of plane 0's faces and the ghost plane below gathered with cp.async during the previous task's
// last plane flux + stores: no global-memory latency at a task switch); 4 = one basic block per plane (the
// next-step wave speeds of plane 0's dummy pass are computed and discarded instead of branched around);
// 8 = the lower z-face flux enters the accumulators first (10 fewer doubles live across the x / y faces)
constexpr int kOptPingPong = 1, kOptPrefetch = 2, kOptOneBlock = 4, kOptEarlyZ = 8;

template <int S, int CR, int NS, int WPC, int OPT = 0>
__device__ __forceinline__ void euler3d_dense_body(const StepArgs& a, int n_items)
{
    constexpr bool PINGPONG = (OPT & kOptPingPong) != 0;
    constexpr bool PF       = (OPT & kOptPrefetch) != 0;
    constexpr bool ONEBLOCK = (OPT & kOptOneBlock) != 0;
    constexpr bool EARLYZ   = (OPT & kOptEarlyZ) != 0;
    using C           = March3DenseCfg<S, CR, NS, WPC>;
    constexpr int NV  = 5;
    constexpr int SS  = C::SS;
    constexpr int N   = C::N;
    constexpr int PLD = C::PLD;
    constexpr int FS  = C::FS;
    constexpr int HF  = S / 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WPC * NS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int xq = lane & 3, yy = lane >> 2;   // marching role: x pair, y row of the block
    const int side = lane >> 3, bt = lane & 7; // boundary-face role: side (x-,x+,y-,y+), position
    double*   ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::WARP_DOUBLES;
    double*   sBF  = ring + C::RING;
    double*   sST  = sBF + C::BF + lane; // this lane's column of the ghost staging buffer
    double*   sGZ  = sBF + C::BF + C::ST + 2 * lane; // this lane's (A | B) slots of the staged ghost plane
    uint64_t* bar  = bars + warp * NS;

    // ---- this warp's tasks: first one static (warp gw takes task gw: the chip starts on one Morton
    // window), every further one drawn from the device counter a.queue when there is one
    const int n_tasks = n_items * C::NB;
    const int gw = blockIdx.x * WPC + warp, nw_all = gridDim.x * WPC;
    int          tau_cur = (gw < n_tasks) ? gw : n_tasks; // n_tasks = "none"
    int          tau_nxt = n_tasks;
    unsigned int nxt_raw = 0; // lane 0: the ticket drawn for the task after the current one
    int          kc = 0;      // sequence number of the current task
    // Tickets run one task ahead of their use: the one read at plane 2 of task k (-> task k+1) was drawn at plane 2
    // of task k-1, and the next one is drawn right after the read.  (Drawn at the end of a task and read "later",
    // the compiler moved the read up to the atomic: its whole latency, 1.6 % of the warp time, per task.)
    auto draw_ticket = [&]() {
        if (a.queue != nullptr && lane == 0) nxt_raw = atomicAdd(a.queue, 1u);
    };
    auto take_next = [&]() {
        if (a.queue == nullptr)
            tau_nxt = tau_cur + nw_all;
        else
        {
            const unsigned int t = __shfl_sync(0xffffffffu, nxt_raw, 0);
            tau_nxt              = (t < 0x40000000u) ? nw_all + (int)t : n_tasks;
            draw_ticket();
        }
    };
    static_assert(S - NS * CR + CR - 1 >= 2, "the ring reaches the next task only after plane 2 (take_next)");
    if (tau_cur < n_tasks) draw_ticket();

    // halo tables of a task's patch, one 32-bit piece per lane (lanes 0-23: 6 x 4 neighbor indices,
    // lane 24: level, lanes 25-30: the 6 relation bytes), prefetched for the NEXT task
    auto tab_load = [&](int tau) -> int {
        const int item = tau / C::NB;
        const int q    = a.list ? a.list[item] : item;
        return tab_piece3(a.nbr, a.level, a.meta, q, lane); // lane 31: q (a.list resolved with the prefetch)
    };
    int tab = (tau_cur < n_tasks) ? tab_load(tau_cur) : 0, tab_nxt = 0;

    auto task_at = [&](int tau, int tabv, int& p, int& bx, int& by) {
        const int blk = tau % C::NB;
        bx            = blk % C::NBX;
        by            = blk / C::NBX;
        p             = __shfl_sync(0xffffffffu, tabv, 31);
    };

    if (lane == 0)
    {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bar[s], C::WHOLE ? 1 : 32);
    }
    __syncwarp();

    // ---- producer side of the ring (lane 0 issues the bulk copies; all lanes track the counters)
    int  ik = 0, ic = 0, ist = 0; // next chunk to issue: task, chunk in task, stage
    auto issue_next = [&]() {
        const int tau = (ik == kc) ? tau_cur : tau_nxt;
        if (tau >= n_tasks) return;
        int p, bx, by;
        task_at(tau, (ik == kc) ? tab : tab_nxt, p, bx, by);
        double* dst = ring + ist * C::STAGE;
        if constexpr (C::WHOLE)
        {
            if (lane == 0)
            {
                mbar_expect_tx(&bar[ist], C::STAGE * 8);
                const size_t go = (size_t)p * N + (size_t)(ic * CR) * SS;
#pragma unroll
                for (int f = 0; f < NV; ++f) bulk_g2s(dst + f * FS, a.cur.p[f] + go, FS * 8, &bar[ist]);
            }
        }
        else
        {
            // wider patches: the block's 8 rows x 8 cells of a field-plane (64-byte row pieces at the patch's row
            // pitch) moved as ONE 16-byte cp.async piece per lane into the same dense 8 x 8 tile the 8^3 path
            // stages (the cells left and right of the block are boundary faces of the task: they come through the
            // ghost gathers), then the lanes arrive on the stage's mbarrier (32 arrivals per phase)
            const size_t go = (size_t)p * N + (size_t)(ic * CR) * SS + (size_t)(8 * by + (lane >> 2)) * S + 8 * bx +
                              2 * (lane & 3);
#pragma unroll
            for (int f = 0; f < NV; ++f)
#pragma unroll
                for (int j = 0; j < CR; ++j)
                    cp_async16(dst + f * FS + j * PLD + 2 * lane, a.cur.p[f] + go + (size_t)j * SS);
            cp_async_mbar_arrive(&bar[ist]);
        }
        if (++ic == C::NCH)
        {
            ic = 0;
            ++ik;
        }
        if (++ist == NS) ist = 0;
    };
#pragma unroll
    for (int s = 0; s < NS; ++s) issue_next();

    // ---- step scalars
    double       rem_after;
    const double dt = resolve_step_dt(a.sc, rem_after);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.sc.dtmin_in != nullptr)
    {
        *a.sc.dt_taken      = dt;
        *a.sc.remaining_out = rem_after;
    }
    const double g = a.gamma, gm1 = a.gamma - 1.0;
    double       cand = DBL_MAX; // min dx/speed over finished levels
    double       sxm = 0.0, sym = 0.0, szm = 0.0;
    int          lvl_prev = -1;

    int cst = 0, cph = 0; // consumer: stage and phase parity of the next chunk to wait for

    // ---- boundary-face role of this lane for a task: where the ghost cell of plane z comes from, resolved
    // ONCE per task (interior coordinates; gy / gx may be -1 or S = across a face)
    GhostSrc3 gs;
    auto resolve_gs = [&](int tabv, int p, int bx, int by) {
        const int  x0 = 8 * bx, y0 = 8 * by;
        const int  bd       = (side < 2) ? 4 + side : side; // tree direction of the side
        const bool internal = (side == 0)   ? (bx > 0)
                              : (side == 1) ? (bx < C::NBX - 1)
                              : (side == 2) ? (by > 0)
                                            : (by < C::NBY - 1);
        int g_y, g_x, i_y, i_x; // interior coordinates of the ghost / interior cell of the face
        if (side < 2)
        {
            g_y = i_y = y0 + bt;
            g_x       = side ? x0 + 8 : x0 - 1;
            i_x       = side ? x0 + 7 : x0;
        }
        else
        {
            g_x = i_x = x0 + bt;
            g_y       = (side == 3) ? y0 + 8 : y0 - 1;
            i_y       = (side == 3) ? y0 + 7 : y0;
        }
        const int bm = tab_meta3(tabv, p, bd);
        int4      bnb;
        bnb.x = __shfl_sync(0xffffffffu, tabv, bd * 4);
        bnb.y = __shfl_sync(0xffffffffu, tabv, bd * 4 + 1);
        bnb.z = __shfl_sync(0xffffffffu, tabv, bd * 4 + 2);
        bnb.w = __shfl_sync(0xffffffffu, tabv, bd * 4 + 3);
        const int rel = internal ? 0 : (bm & 3);
        int f_y = g_y, f_x = g_x; // mirrored into the neighbor's frame (patch_utils.hpp:322-327)
        if (!internal)
        {
            if (side < 2)
                f_x += (side & 1) ? -S : S;
            else
                f_y += (side & 1) ? -S : S;
        }
        gs.finer  = 0;
        gs.pair   = 0;
        gs.zshift = 0;
        gs.zbase  = 0;
        gs.q0 = gs.q1 = p;
        if (internal)
        {
            gs.off  = g_y * S + g_x; // block side inside the patch: the own patch's cell
            gs.pair = (side >= 2);
        }
        else if (rel == 1)
        {
            gs.q0 = gs.q1 = bnb.x; // same_t (patch_utils.hpp:315-332)
            gs.off        = f_y * S + f_x;
            gs.pair       = (side >= 2); // a contiguous row of 8 cells: two per copy, even lanes only
        }
        else if (rel == 3)
        {
            // coarser_t: injection of the covering coarse cell (patch_utils.hpp:388-441)
            const int qz = (bm >> 2) & 1, qy = (bm >> 3) & 1, qx = (bm >> 4) & 1;
            gs.q0 = gs.q1 = bnb.x;
            gs.off    = (qy * HF + f_y / 2) * S + (qx * HF + f_x / 2);
            gs.zbase  = qz * HF;
            gs.zshift = 1;
        }
        else if (rel == 2)
        {
            // finer_t: mean of 8 fine cells (patch_utils.hpp:334-386); finer-neighbor index =
            // z half (bit 0) + 2 x half along the other tangential dim (neighbor.hpp:316-337)
            const int t = ((side < 2) ? g_y : g_x) / HF;
            gs.q0       = t ? bnb.z : bnb.x;
            gs.q1       = t ? bnb.w : bnb.y;
            gs.off      = ((f_y * 2) % S) * S + ((f_x * 2) % S);
            gs.finer    = 1;
        }
        else
            gs.off = i_y * S + i_x; // relation "none" (never in a periodic balanced tree): zero gradient
    };
    // ghost cell of this lane's boundary face of plane z -> staging column (asynchronously)
    auto bnd_issue = [&](int z) {
        double* st = sST;
        if (gs.finer)
        {
            const size_t o = (size_t)(z < HF ? gs.q0 : gs.q1) * N + (size_t)(2 * (z % HF)) * SS + gs.off;
            double       t[NV];
            fine_mean5(a.cur, o, S, SS, t);
#pragma unroll
            for (int f = 0; f < NV; ++f) st[f * 32] = t[f];
        }
        else
        {
            // the load / store unit works per REQUEST (scattered 8-byte copies, not bytes or cache lines, are what
            // these gathers cost: profiles/r02_summary.md): the contiguous y-side rows go two cells per copy
            const size_t o = (size_t)gs.q0 * N + (size_t)(gs.zbase + (z >> gs.zshift)) * SS + gs.off;
            if (gs.pair)
            {
                if ((bt & 1) == 0)
                {
#pragma unroll
                    for (int f = 0; f < NV; ++f) cp_async16(st + f * 32, a.cur.p[f] + o);
                }
            }
            else
            {
#pragma unroll
                for (int f = 0; f < NV; ++f) cp_async8(st + f * 32, a.cur.p[f] + o);
            }
        }
    };
    // source of the ghost cells of this lane's column pair across a z face (d = 0 below, 1 above)
    struct ZSrc
    {
        size_t o;
        int    dB, fin;
    };
    auto zsource = [&](int d, int tabv, int p, int bx, int by) -> ZSrc {
        const int m = tab_meta3(tabv, p, d);
        int4      nb;
        nb.x = __shfl_sync(0xffffffffu, tabv, d * 4);
        nb.y = __shfl_sync(0xffffffffu, tabv, d * 4 + 1);
        nb.z = __shfl_sync(0xffffffffu, tabv, d * 4 + 2);
        nb.w = __shfl_sync(0xffffffffu, tabv, d * 4 + 3);
        const int rel = m & 3;
        const int y = 8 * by + yy, x = 8 * bx + 2 * xq; // interior coordinates of cell A
        const int zf = d ? 0 : S - 1;                   // mirrored source plane of a same-level neighbor
        int       q = p, dB = 1, fin = 0;
        int       off = (d ? S - 1 : 0) * SS + y * S + x; // "none": the own boundary cell
        if (rel == 1)
        {
            q   = nb.x;
            off = zf * SS + y * S + x;
        }
        else if (rel == 3)
        {
            const int qz = (m >> 2) & 1, qy = (m >> 3) & 1, qx = (m >> 4) & 1;
            q   = nb.x;
            off = (qz * HF + zf / 2) * SS + (qy * HF + y / 2) * S + (qx * HF + x / 2);
            dB  = 0;
        }
        else if (rel == 2)
        {
            const int fi = y / HF + 2 * (x / HF);
            q            = (fi == 0) ? nb.x : (fi == 1) ? nb.y : (fi == 2) ? nb.z : nb.w;
            off          = ((zf * 2) % S) * SS + ((y * 2) % S) * S + ((x * 2) % S);
            dB           = 2;
            fin          = 1;
        }
        return ZSrc{ (size_t)q * N + off, dB, fin };
    };
    // ... into registers
    auto zghost = [&](int d, int tabv, int p, int bx, int by, double (&vA)[NV], double (&vB)[NV]) {
        const ZSrc zs = zsource(d, tabv, p, bx, by);
        if (zs.fin)
        {
            double tA[NV], tB[NV];
            fine_mean5(a.cur, zs.o, S, SS, tA);
            fine_mean5(a.cur, zs.o + 2, S, SS, tB);
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                vA[f] = tA[f];
                vB[f] = tB[f];
            }
        }
        else
        {
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                vA[f] = __ldg(a.cur.p[f] + zs.o);
                vB[f] = __ldg(a.cur.p[f] + zs.o + zs.dB);
            }
        }
    };
    // ... into the staged ghost plane (asynchronously; the lane reads back only its own slots)
    auto zghost_stage = [&](int tabv, int p, int bx, int by) {
        const ZSrc zs = zsource(0, tabv, p, bx, by);
        if (zs.fin)
        {
            double tA[NV], tB[NV];
            fine_mean5(a.cur, zs.o, S, SS, tA);
            fine_mean5(a.cur, zs.o + 2, S, SS, tB);
#pragma unroll
            for (int f = 0; f < NV; ++f) *reinterpret_cast<double2*>(sGZ + f * 64) = make_double2(tA[f], tB[f]);
        }
        else
        {
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                cp_async8(sGZ + f * 64, a.cur.p[f] + zs.o);
                cp_async8(sGZ + f * 64 + 1, a.cur.p[f] + zs.o + zs.dB);
            }
        }
    };
    // everything of a task that comes from global memory besides the ring: issued one task ahead
    auto prefetch_task = [&](int tau, int tabv) {
        int p, bx, by;
        task_at(tau, tabv, p, bx, by);
        resolve_gs(tabv, p, bx, by);
        bnd_issue(0);
        zghost_stage(tabv, p, bx, by);
        cp_async_commit();
    };
    if constexpr (PF)
    {
        if (tau_cur < n_tasks) prefetch_task(tau_cur, tab);
    }

    while (tau_cur < n_tasks)
    {
        int p, bx, by;
        task_at(tau_cur, tab, p, bx, by);
        const int lvl = __shfl_sync(0xffffffffu, tab, 24);
        if (lvl != lvl_prev && lvl_prev >= 0)
        {
            if (sxm > 1e-12) cand = fmin(cand, a.dx[lvl_prev][0] / sxm);
            if (sym > 1e-12) cand = fmin(cand, a.dx[lvl_prev][1] / sym);
            if (szm > 1e-12) cand = fmin(cand, a.dx[lvl_prev][2] / szm);
            sxm = sym = szm = 0.0;
        }
        lvl_prev         = lvl;
        const double hx  = -0.5 * (dt / a.dx[lvl][0]); // -0.5 dt/dx, x = fastest layout dim
        const double hy  = -0.5 * (dt / a.dx[lvl][1]);
        const double hz  = -0.5 * (dt / a.dx[lvl][2]);
        const double nhz = -hz;
        const size_t pb  = (size_t)p * N;
        const int    x0 = 8 * bx, y0 = 8 * by; // interior coordinates of the block's first cell
        if constexpr (!PF) resolve_gs(tab, p, bx, by);
        // interior cell of this lane's boundary face inside the staged block rows
        const int ioff_s = (side < 2) ? bt * 8 + (side ? 7 : 0) : ((side == 3) ? 7 : 0) * 8 + bt;
        // flux of this lane's boundary face of plane z: ghost cell from the staging column, interior
        // cell from the staged plane (`pl` = first field of that plane inside the ring) -> sBF[z & 1].
        // G = F(ghost) + F(interior) -/+ smax (U_interior - U_ghost), one branch-free form for all sides.
        const double bsgn = (side & 1) ? -1.0 : 1.0;
        auto bnd_flux = [&](int z, const double* pl) {
            const double* st = sST;
            Cell3         gc, ic_;
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                gc.u[f]  = st[f * 32];
                ic_.u[f] = pl[f * FS + ioff_s];
            }
            prims3(gc, g, gm1);
            prims3(ic_, g, gm1);
            double F[NV];
            flux3_bnd(gc, ic_, side >> 1, bsgn, F);
            // parked as [buffer][chunk][face] double2: 16-byte slots of consecutive faces are consecutive in
            // shared memory (the former face-major layout, 48 bytes per face, made lanes 0 / 8 / 16 / 24
            // collide on every store and the two sides of a face pair on every load: 29 % of all shared
            // wavefronts were conflict replays, profiles/r02c_euler3d_dense_ncu_summary.txt)
            double2* o = reinterpret_cast<double2*>(sBF) + (z & 1) * 96 + lane;
            o[0]       = make_double2(F[0], F[1]);
            o[32]      = make_double2(F[2], F[3]);
            o[64]      = make_double2(F[4], 0.0);
        };
        // ---- task prologue: ghost plane below, boundary fluxes of plane 0 (needs the task's first
        // chunk: it was requested while the previous task was marched)
        double gzA[NV], gzB[NV];
        if constexpr (PF)
        {
            // both were gathered into shared memory while the previous task finished
            mbar_wait(&bar[cst], cph);
            cp_async_wait<0>();
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                const double2 v = *reinterpret_cast<const double2*>(sGZ + f * 64);
                gzA[f]          = v.x;
                gzB[f]          = v.y;
            }
        }
        else
        {
            bnd_issue(0);
            cp_async_commit();
            zghost(0, tab, p, bx, by, gzA, gzB);
            mbar_wait(&bar[cst], cph);
            cp_async_wait<0>();
        }
        bnd_flux(0, ring + cst * C::STAGE);
        __syncwarp();

        struct PlaneState
        {
            Cell3  A, B;               // records of the lane's two cells
            double accA[NV], accB[NV]; // U + hx dFx + hy dFy - hz Fz(low)
        };
        PlaneState s0, s1;
        size_t     go = pb + (size_t)(y0 + yy) * S + x0 + 2 * xq; // the lane's pair in plane 0
        int        sl = 0;                                        // slot of the streamed plane in its chunk

        // finish a plane: add the upper z-face flux, store the (A | B) pair, wave speeds of the new state
        auto finish = [&](const PlaneState& pv, const double (&GzA)[NV], const double (&GzB)[NV], bool fin) {
            double rA[NV], rB[NV];
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                rA[f] = fma(hz, GzA[f], pv.accA[f]);
                rB[f] = fma(hz, GzB[f], pv.accB[f]);
                if (fin) *reinterpret_cast<double2*>(a.nxt.p[f] + go) = make_double2(rA[f], rB[f]);
            }
            if constexpr (!ONEBLOCK)
            {
                if (!fin) return; // plane 0: `pv` is the ghost plane below (warp-uniform branch)
                go += SS;
            }
            else
                go += fin ? SS : 0; // plane 0's dummy pass runs through the same instructions
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
            {
                const double* n    = c2 ? rB : rA;
                const double  irho = rcp_nr2(n[0]);
                double        K    = n[1] * n[1];
                K                  = fma(n[2], n[2], K);
                K                  = fma(n[3], n[3], K);
                K *= 0.5 * irho;
                const double pr = gm1 * (n[4] - K);
                const double cs = sqrt_nr2(g * pr * irho);
                const double vx = fabs(n[1] * irho) + cs, vy = fabs(n[2] * irho) + cs, vz = fabs(n[3] * irho) + cs;
                if constexpr (!ONEBLOCK)
                {
                    sxm = pos_max(sxm, vx);
                    sym = pos_max(sym, vy);
                    szm = pos_max(szm, vz);
                }
                else
                {
                    // the dummy pass may produce anything (NaN included): selected away, never compared
                    sxm = fin ? pos_max(sxm, vx) : sxm;
                    sym = fin ? pos_max(sym, vy) : sym;
                    szm = fin ? pos_max(szm, vz) : szm;
                }
            }
        };

        // one interior plane z: `pv` = state of plane z-1 (or of the ghost plane below), `nw` = state of
        // plane z.  ONE instance of this body exists in the kernel (rolled loop, instruction cache).
        auto plane_step = [&](int z, const PlaneState& pv, PlaneState& nw, bool fin, bool last) {
            const int BUF = z & 1;
            if (!last)
            {
                bnd_issue(z + 1); // ghost cells of the next plane's faces: in flight during this plane
                cp_async_commit();
            }
            else
                zghost(1, tab, p, bx, by, gzA, gzB); // in flight during the last plane
            if (sl == 0) mbar_wait(&bar[cst], cph);
            const double* src = ring + cst * C::STAGE + sl * PLD + yy * 8 + 2 * xq; // the staged tile is 8 x 8
            const int     lo  = (xq > 0) ? -1 : 0; // left cell (clamped: replaced by a parked flux)
            double        Lu[NV];
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                const double2 v0 = *reinterpret_cast<const double2*>(src + f * FS);
                Lu[f]            = src[f * FS + lo];
                nw.A.u[f]        = v0.x;
                nw.B.u[f]        = v0.y;
            }
            prims3(nw.A, g, gm1);
            prims3(nw.B, g, gm1);
            double GzA[NV], GzB[NV];
            flux3<2>(pv.A, nw.A, GzA);
            flux3<2>(pv.B, nw.B, GzB);
            finish(pv, GzA, GzB, fin);
            if constexpr (EARLYZ)
            {
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    nw.accA[f] = fma(nhz, GzA[f], nw.A.u[f]);
                    nw.accB[f] = fma(nhz, GzB[f], nw.B.u[f]);
                }
            }
            const double2* bfp = reinterpret_cast<const double2*>(sBF) + BUF * 96;
            // ---- x faces
            {
                const double2* bf = bfp + ((xq >> 1) * 8 + yy);
                const double2  b0 = bf[0], b1 = bf[32], b2 = bf[64];
                const double   bfl[NV] = { b0.x, b0.y, b1.x, b1.y, b2.x };
                Cell3          L;
#pragma unroll
                for (int f = 0; f < NV; ++f) L.u[f] = Lu[f];
                L.p  = __shfl_up_sync(0xffffffffu, nw.B.p, 1);
                L.a  = __shfl_up_sync(0xffffffffu, nw.B.a, 1);
                L.ir = __shfl_up_sync(0xffffffffu, nw.B.ir, 1);
                double GL[NV], GM[NV];
                flux3<0>(L, nw.A, GL);
                flux3<0>(nw.A, nw.B, GM);
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    if (xq == 0) GL[f] = bfl[f];
                    double GR = __shfl_down_sync(0xffffffffu, GL[f], 1);
                    if (xq == 3) GR = bfl[f];
                    nw.accA[f] = fma(hx, GM[f] - GL[f], EARLYZ ? nw.accA[f] : nw.A.u[f]);
                    nw.accB[f] = fma(hx, GR - GM[f], EARLYZ ? nw.accB[f] : nw.B.u[f]);
                }
            }
            // ---- y faces
            {
                const int yl = (yy > 0) ? -8 : 0; // staged row below (clamped: replaced by a parked flux)
                Cell3     YA, YB;
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    const double2 v0 = *reinterpret_cast<const double2*>(src + f * FS + yl);
                    YA.u[f]          = v0.x;
                    YB.u[f]          = v0.y;
                }
                YA.p  = __shfl_up_sync(0xffffffffu, nw.A.p, 4);
                YA.a  = __shfl_up_sync(0xffffffffu, nw.A.a, 4);
                YA.ir = __shfl_up_sync(0xffffffffu, nw.A.ir, 4);
                YB.p  = __shfl_up_sync(0xffffffffu, nw.B.p, 4);
                YB.a  = __shfl_up_sync(0xffffffffu, nw.B.a, 4);
                YB.ir = __shfl_up_sync(0xffffffffu, nw.B.ir, 4);
                double GyA[NV], GyB[NV];
                flux3<1>(YA, nw.A, GyA);
                flux3<1>(YB, nw.B, GyB);
                const double2* bf = bfp + ((2 + (yy >> 2)) * 8 + 2 * xq);
                const double2  c0 = bf[0], c1 = bf[32], c2 = bf[64], d0 = bf[1], d1 = bf[33], d2 = bf[65];
                const double  bA[NV] = { c0.x, c0.y, c1.x, c1.y, c2.x };
                const double  bB[NV] = { d0.x, d0.y, d1.x, d1.y, d2.x };
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    if (yy == 0)
                    {
                        GyA[f] = bA[f];
                        GyB[f] = bB[f];
                    }
                    double upA = __shfl_down_sync(0xffffffffu, GyA[f], 4);
                    double upB = __shfl_down_sync(0xffffffffu, GyB[f], 4);
                    if (yy == 7)
                    {
                        upA = bA[f];
                        upB = bB[f];
                    }
                    nw.accA[f] = fma(hy, upA - GyA[f], nw.accA[f]);
                    nw.accB[f] = fma(hy, upB - GyB[f], nw.accB[f]);
                }
            }
            // ---- lower z face
            if constexpr (!EARLYZ)
            {
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    nw.accA[f] = fma(nhz, GzA[f], nw.accA[f]);
                    nw.accB[f] = fma(nhz, GzB[f], nw.accB[f]);
                }
            }
            if (!last)
            {
                // boundary fluxes of plane z+1: its ghost cells have landed in the staging column; its
                // interior cells sit in the ring -- same stage, or the next one (wait for it here, the
                // wait at the top of the next plane then falls through)
                cp_async_wait<0>();
                int stg = cst, slot = sl + 1;
                if (slot == CR)
                {
                    slot = 0;
                    stg  = (cst + 1 == NS) ? 0 : cst + 1;
                    mbar_wait(&bar[stg], (cst + 1 == NS) ? (cph ^ 1) : cph);
                }
                bnd_flux(z + 1, ring + stg * C::STAGE + slot * PLD);
            }
            __syncwarp();
            if (++sl == CR)
            {
                // every lane has consumed its values of this stage's last plane: refill it
                sl = 0;
                issue_next();
                if (++cst == NS)
                {
                    cst = 0;
                    cph ^= 1;
                }
            }
        };
        // ghost plane below -> records (accumulators: any finite values, never stored)
#pragma unroll
        for (int f = 0; f < NV; ++f)
        {
            s0.A.u[f] = s0.accA[f] = gzA[f];
            s0.B.u[f] = s0.accB[f] = gzB[f];
        }
        prims3(s0.A, g, gm1);
        prims3(s0.B, g, gm1);
        if constexpr (PINGPONG)
        {
            // two planes per trip, the two plane states swapping roles: no copy of the 26 carried doubles per
            // plane (6 % of the issued instructions were those moves), at twice the loop body
#pragma unroll 1
            for (int z = 0; z < S; z += 2)
            {
                if (z == 2)
                {
                    take_next();
                    if (tau_nxt < n_tasks) tab_nxt = tab_load(tau_nxt);
                }
                plane_step(z, s0, s1, z > 0, false);
                plane_step(z + 1, s1, s0, true, z + 1 == S - 1);
            }
        }
        else
        {
#pragma unroll 1
            for (int z = 0; z < S; ++z)
            {
                if (z == 2)
                {
                    take_next(); // the next task: request its halo tables
                    if (tau_nxt < n_tasks) tab_nxt = tab_load(tau_nxt);
                }
                plane_step(z, s0, s1, z > 0, z == S - 1);
                s0 = s1;
            }
        }
        if constexpr (PF)
        {
            // the next task's boundary ghosts and ghost plane below: in flight during this task's last
            // z-face flux and stores (gs now belongs to the next task; sST / sGZ are free since plane S-2)
            if (tau_nxt < n_tasks) prefetch_task(tau_nxt, tab_nxt);
        }
        // ghost plane above: z-face flux into plane S-1, finish it
        {
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                s1.A.u[f] = gzA[f];
                s1.B.u[f] = gzB[f];
            }
            prims3(s1.A, g, gm1);
            prims3(s1.B, g, gm1);
            double GzA[NV], GzB[NV];
            flux3<2>(s0.A, s1.A, GzA);
            flux3<2>(s0.B, s1.B, GzB);
            finish(s0, GzA, GzB, true);
        }
        __syncwarp(); // sBF is rewritten by the next task
        ++kc;
        tau_cur = tau_nxt;
        tab     = tab_nxt;
    }

    if (a.sc.dtmin_out != nullptr)
    {
        if (lvl_prev >= 0)
        {
            if (sxm > 1e-12) cand = fmin(cand, a.dx[lvl_prev][0] / sxm);
            if (sym > 1e-12) cand = fmin(cand, a.dx[lvl_prev][1] / sym);
            if (szm > 1e-12) cand = fmin(cand, a.dx[lvl_prev][2] / szm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0 && kc > 0) atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
    }
}

// occupancy by CTAs per SM (register budget = 64 K / (MINB x CTA threads rounded up to 128))
template <int S, int CR, int NS, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
euler3d_dense_kernel(const __grid_constant__ StepArgs a, int n_items)
{
    euler3d_dense_body<S, CR, NS, WPC>(a, n_items);
}
// two planes per loop trip (ping-pong plane states)
template <int S, int CR, int NS, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
euler3d_dense_kernel_pp(const __grid_constant__ StepArgs a, int n_items)
{
    euler3d_dense_body<S, CR, NS, WPC, kOptPingPong>(a, n_items);
}
// the body's OPT variants (A/B runs)
template <int S, int CR, int NS, int WPC, int MINB, int OPT>
__global__ void __launch_bounds__(WPC * 32, MINB)
euler3d_dense_kernel_o(const __grid_constant__ StepArgs a, int n_items)
{
    euler3d_dense_body<S, CR, NS, WPC, OPT>(a, n_items);
}

} // namespace amrb
