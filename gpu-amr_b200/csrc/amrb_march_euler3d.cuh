// Fused Euler step, third generation (3D): warp-autonomous plane marching.
//
// The 3D form of euler2d_march_kernel (amrb_march_euler.cuh):
//   * task = one 8 x 8 (x, y) column block of one patch, all S planes along z (the slowest layout
//     dim); an 8^3 patch is one task, a 16^3 patch four.  A warp walks its tasks back to back and
//     never meets a block barrier.
//   * the planes of a task stream through a warp-private shared-memory ring of NS stages: for 8^3
//     patches TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), one per field moving CR
//     whole padded planes (contiguous in the pool); for wider patches the block's 8 padded rows of
//     a field-plane are too small for a warp's bulk-copy rate (tools/tma_bench.cu), so all lanes
//     move 16-byte cp.async pieces and arrive on the stage's mbarrier.  Only the S interior planes
//     are streamed; the ghost planes below / above are never read from the pool.
//   * tasks: the first one of a warp is static (warp w of W takes task w: the chip starts on one
//     Morton window), every further one is drawn from a device counter, so that warps slowed down
//     by coarse/fine faces do not drift out of the window and ghost gathers keep hitting L2; the
//     halo tables of the NEXT task are fetched one 32-bit piece per lane while the current task is
//     marched and handed round by shuffles.
//   * lane = (y row 0..7) x (x pair 0..3): a lane owns two x-adjacent cells of the plane and marches
//     them along z.  The cell record (U, p, a, 1/rho) is derived once, in registers; the z-face
//     flux is carried in registers (folded into the running update); x-faces as in 2D (left
//     lane's p / a / 1/rho and the right lane's left flux by warp shuffle); the lower y-face uses
//     the record of lane-4 (U from the staged plane, p / a / 1/rho by shuffle), the upper y-face
//     flux is lane+4's lower flux by shuffle: every interior face flux is computed exactly once.
//   * ghost cells are never written: the 32 lateral boundary faces of a plane (8 per side) are one
//     per lane -- lane = side x position gathers the ghost cell straight from the neighbor patch
//     interior through the halo tables (same / coarser injection / finer restriction, reference
//     summation order) or, for a block boundary inside the patch, from the patch itself, computes
//     that face's flux and parks it in a double-buffered 1.5 KB array; the loads for plane z+1
//     are issued before plane z is computed.  Ghost planes across the z faces are gathered into
//     the registers of the lanes that march those columns.
//   * stores cover whole padded planes (ghost rows / columns get copies of the adjacent cell):
//     the interior planes of a field-patch are then one contiguous run of fully written 32-byte
//     sectors (see the 2D kernel for the measured reason).
//
// Arithmetic follows include/solver/EulerPhysics.hpp:74-129 and amr_solver.hpp:265-353 of the
// reference; 0.5 of the Rusanov flux is folded into dt/dx (exact); the update is accumulated as
// ((U + hx dFx) + hy dFy) - hz Fz_low + hz Fz_up with fused multiply-adds (rounding-level
// difference from the reference's association; parity bound 1e-12 field-max-normalised).
#pragma once
#include "amrb_march_euler.cuh"

namespace amrb
{

// per-thread 8-byte asynchronous copy global -> shared (SASS LDGSTS): no register staging, the
// load is in flight as soon as it is issued
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// 16-byte variant and "arrive on an mbarrier once my earlier cp.async have landed" (count not
// incremented: the barrier is initialised with one arrival per lane)
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct Cell3
{
    double u[5];
    double p, a, ir;
};

__device__ __forceinline__ void prims3(Cell3& c, double g, double gm1)
{
    c.ir     = rcp_nr2(c.u[0]);
    double K = c.u[1] * c.u[1];
    K        = fma(c.u[2], c.u[2], K);
    K        = fma(c.u[3], c.u[3], K);
    K *= 0.5 * c.ir;
    c.p = gm1 * (c.u[4] - K);
    c.a = sqrt_nr2(g * c.p * c.ir);
}

// G = F(L) + F(R) - smax (U_R - U_L) across a face normal to solver direction DS (0 x, 1 y, 2 z), in wave form
// (see flux2, amrb_march_euler.cuh): G_k = U_k,L (u_L + smax) + U_k,R (u_R - smax) [+ pressure terms]
template <int DS>
__device__ __forceinline__ void flux3(const Cell3& L, const Cell3& R, double (&G)[5])
{
    const double uL = L.u[1 + DS] * L.ir, uR = R.u[1 + DS] * R.ir;
    const double sm = pos_max(fabs(uL) + L.a, fabs(uR) + R.a);
    const double wp = uL + sm, wm = uR - sm;
    G[0]            = fma(R.u[0], wm, L.u[0] * wp);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        G[1 + k] = (k == DS) ? fma(R.u[1 + k], wm, fma(L.u[1 + k], wp, L.p + R.p)) : fma(R.u[1 + k], wm, L.u[1 + k] * wp);
    G[4] = fma(R.u[4], wm, fma(L.u[4], wp, fma(L.p, uL, R.p * uR)));
}
// the same across a lateral boundary face of a plane: ghost cell `gc`, interior cell `ic`, face normal x (ds = 0) or
// y (ds = 1) chosen at run time (the 32 faces of a plane are spread over the lanes of a warp), bsgn = +1 when the
// ghost cell is the lower one.  G = F(gc) + F(ic) + bsgn smax (U_gc - U_ic), one branch-free form for all sides.
__device__ __forceinline__ void flux3_bnd(const Cell3& gc, const Cell3& ic, int ds, double bsgn, double (&F)[5])
{
    const double mG = ds ? gc.u[2] : gc.u[1], mI = ds ? ic.u[2] : ic.u[1];
    const double uG = mG * gc.ir, uI = mI * ic.ir;
    const double sm = bsgn * pos_max(fabs(uG) + gc.a, fabs(uI) + ic.a);
    const double wG = uG + sm, wI = uI - sm;
    const double ps = gc.p + ic.p;
    F[0]            = fma(ic.u[0], wI, gc.u[0] * wG);
    F[1]            = fma(ic.u[1], wI, fma(gc.u[1], wG, ds ? 0.0 : ps));
    F[2]            = fma(ic.u[2], wI, fma(gc.u[2], wG, ds ? ps : 0.0));
    F[3]            = fma(ic.u[3], wI, gc.u[3] * wG);
    F[4]            = fma(ic.u[4], wI, fma(gc.u[4], wG, fma(gc.p, uG, ic.p * uI)));
}

// Restriction of one ghost cell for all 5 fields: mean of the 2^3 fine cells at offset `o`, summed
// last-dim-fastest like hypercube_offset (patch_utils.hpp:203-234, intergrid_operator.hpp:92-106).
// Kept out of line: coarse/fine faces are a small minority of all faces and the marching loop has
// to stay inside the instruction cache.
static __device__ __noinline__ void fine_mean5(const FieldPtrs& cur, size_t o, int P, int PP, double* out)
{
#pragma unroll 1
    for (int f = 0; f < 5; ++f)
    {
        const double* __restrict__ s = cur.p[f] + o;
        double sum = 0.0;
        sum += __ldg(s);
        sum += __ldg(s + 1);
        sum += __ldg(s + P);
        sum += __ldg(s + P + 1);
        sum += __ldg(s + PP);
        sum += __ldg(s + PP + 1);
        sum += __ldg(s + PP + P);
        sum += __ldg(s + PP + P + 1);
        out[f] = sum / 8.0;
    }
}

// where a boundary ghost cell of plane z is gathered from (resolved once per task)
struct GhostSrc3
{
    int q0, q1; // source patch for z in the lower / upper half of the patch (finer), else equal
    int off;    // (y, x) offset inside the source plane
    int zbase;  // single source: plane zbase + (z >> zshift); finer: plane H + 2 (z mod S/2)
    int zshift;
    int finer;
    int pair;   // dense kernel: y-side ghost row of a same-level neighbor (contiguous: two cells per copy)
};

template <int S, int H, int CR, int NS, int WPC>
struct March3Cfg
{
    using G                    = Geo<3, S, H>;
    static constexpr int NV    = 5;
    static constexpr int P     = G::P;
    static constexpr int PP    = P * P;
    static constexpr int NBX   = S / 8, NBY = S / 8;
    static constexpr int NB    = NBX * NBY;            // tasks per patch
    static constexpr bool WHOLE = (NB == 1);           // stream whole padded planes
    static constexpr int PLD   = WHOLE ? PP : 8 * P;   // doubles per staged field-plane
    static constexpr int ROW0  = WHOLE ? H : 0;        // staged row of the block's first row
    static constexpr int NCH   = S / CR;               // chunks per task
    static constexpr int STAGE = NV * CR * PLD;
    static constexpr int RING  = NS * STAGE;
    static constexpr int BFW   = 6;                    // doubles per parked boundary flux (5 + pad)
    static constexpr int BF    = 2 * 32 * BFW;         // double-buffered, 32 faces per plane
    static constexpr int ST    = 10 * 32;              // boundary-face inputs in flight: [10][32]
    static constexpr int WARP_DOUBLES = RING + BF + ST;
    static constexpr size_t SMEM      = (size_t)WPC * WARP_DOUBLES * sizeof(double);
    static_assert(S % 8 == 0 && S % 2 == 0, "8 x 8 column blocks");
    static_assert(H == 1, "odd ghost width: (ghost|first) and (second|right) pairs are 16-byte aligned");
    static_assert(S % CR == 0, "chunk shape");
    static_assert(NS * CR <= S, "the ring never reaches beyond the next task");
    static_assert((PLD * 8) % 16 == 0 && (WARP_DOUBLES * 8) % 16 == 0, "bulk copy alignment");
};

template <int S, int H, int CR, int NS, int WPC, int MINB, int CPRING = 0>
__global__ void __launch_bounds__(WPC * 32, MINB)
euler3d_march_kernel(const __grid_constant__ StepArgs a, int n_items)
{
    using C           = March3Cfg<S, H, CR, NS, WPC>;
    using G           = Geo<3, S, H>;
    constexpr int NV  = 5;
    constexpr int P   = C::P;
    constexpr int PP  = C::PP;
    constexpr int PLD = C::PLD;
    constexpr int FS  = CR * PLD; // field stride inside a stage

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WPC * NS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int xq = lane & 3, yy = lane >> 2;   // marching role: x pair, y row of the block
    const int side = lane >> 3, bt = lane & 7; // boundary-face role: side (x-,x+,y-,y+), position
    double*   ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::WARP_DOUBLES;
    double*   sBF  = ring + C::RING;
    double*   sST  = sBF + C::BF + lane; // this lane's column of the staging buffer
    uint64_t* bar  = bars + warp * NS;

    // ---- this warp's tasks.  Static map: warp gw of nw_all takes tasks gw, gw + nw_all, ...  (the chip
    // sweeps one Morton window at a time).  With a.queue the first task is still gw, every further one
    // is drawn from a device counter when the previous one is finished: the window stays tight even
    // when tasks differ in cost (coarse/fine faces), which keeps the ghost gathers in L2.
    const int n_tasks = n_items * C::NB;
    const int gw = blockIdx.x * WPC + warp, nw_all = gridDim.x * WPC; // warp of the grid
    int          tau_cur = (gw < n_tasks) ? gw : n_tasks; // n_tasks = "none"
    int          tau_nxt = n_tasks;
    unsigned int nxt_raw = 0;     // lane 0: value drawn from the counter, not yet broadcast
    bool         nxt_known = true;
    int          kc = 0;          // sequence number of the current task
    // draw the task after tau_cur; the atomic's result is only broadcast when it is first needed
    // (resolve_next), so its latency hides behind the planes in between
    auto fetch_next = [&]() {
        if (a.queue == nullptr)
        {
            tau_nxt   = tau_cur + nw_all;
            nxt_known = true;
        }
        else
        {
            if (lane == 0) nxt_raw = atomicAdd(a.queue, 1u);
            nxt_known = false;
        }
    };
    auto resolve_next = [&]() {
        if (nxt_known) return;
        const unsigned int t = __shfl_sync(0xffffffffu, nxt_raw, 0);
        tau_nxt              = (t < 0x40000000u) ? nw_all + (int)t : n_tasks;
        nxt_known            = true;
    };
    if (tau_cur < n_tasks) fetch_next();

    // halo tables of a task's patch, one 32-bit piece per lane (lanes 0-23: the 6 x 4 neighbor indices,
    // lane 24: level, lanes 25-30: the 6 relation bytes): loaded for the NEXT task while the current
    // one is marched and handed round by shuffles, so that no task starts with a table lookup
    auto tab_load = [&](int tau) -> int {
        const int item = tau / C::NB;
        const int q    = a.list ? a.list[item] : item;
        static_assert(G::NDIR * G::KF == 24, "3D tables");
        return tab_piece3(a.nbr, a.level, a.meta, q, lane);
    };
    int tab = (tau_cur < n_tasks) ? tab_load(tau_cur) : 0, tab_nxt = 0;

    auto task_at = [&](int tau, int& p, int& bx, int& by) {
        const int item = tau / C::NB;
        const int blk  = tau % C::NB;
        bx             = blk % C::NBX;
        by             = blk / C::NBX;
        p              = a.list ? a.list[item] : item;
    };

    if (lane == 0)
    {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bar[s], (C::WHOLE && !CPRING) ? 1 : 32);
    }
    __syncwarp();

    // ---- producer side of the ring (lane 0 issues; all lanes track the counters)
    int  ik = 0, ic = 0, ist = 0; // next chunk to issue: task, chunk in task, stage
    auto issue_next = [&]() {
        // the producer is at most NS chunks ahead: inside the current task or the next one
        if (ik != kc) resolve_next();
        const int tau = (ik == kc) ? tau_cur : tau_nxt;
        if (tau >= n_tasks) return;
        int p, bx, by;
        task_at(tau, p, bx, by);
        double* dst = ring + ist * C::STAGE;
        if constexpr (C::WHOLE && CPRING)
        {
            // (experiment) the same contiguous planes moved as 16-byte cp.async pieces by all lanes
            const size_t go = (size_t)p * G::FLAT + (size_t)(H + ic * CR) * PP;
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                const double* src = a.cur.p[f] + go;
                double*       d   = dst + f * FS;
#pragma unroll
                for (int i = lane; i < FS / 2; i += 32) cp_async16(d + 2 * i, src + 2 * i);
            }
            cp_async_mbar_arrive(&bar[ist]);
        }
        else if constexpr (C::WHOLE)
        {
            // 8^3 patches: CR whole padded planes of a field are contiguous -> one TMA bulk copy each
            if (lane == 0)
            {
                mbar_expect_tx(&bar[ist], C::STAGE * 8);
                const size_t go = (size_t)p * G::FLAT + (size_t)(H + ic * CR) * PP;
#pragma unroll
                for (int f = 0; f < NV; ++f)
                    bulk_g2s(dst + f * FS, a.cur.p[f] + go, FS * 8, &bar[ist]);
            }
        }
        else
        {
            // wider patches: the block's 8 padded rows of a field-plane are only 8 P doubles; a warp's
            // bulk copies complete one at a time (tools/tma_bench.cu), so copies this small cannot feed
            // the march.  All lanes move 16-byte pieces with cp.async instead and arrive on the stage's
            // mbarrier when their pieces have landed (32 arrivals per phase).
            const size_t go =
                (size_t)p * G::FLAT + (size_t)(H + ic * CR) * PP + (size_t)(H + 8 * by) * P;
#pragma unroll
            for (int f = 0; f < NV; ++f)
#pragma unroll
                for (int j = 0; j < CR; ++j)
                {
                    const double* src = a.cur.p[f] + go + (size_t)j * PP;
                    double*       d   = dst + f * FS + j * PLD;
#pragma unroll
                    for (int i = lane; i < PLD / 2; i += 32) cp_async16(d + 2 * i, src + 2 * i);
                }
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

    while (tau_cur < n_tasks)
    {
        int p, bx, by;
        task_at(tau_cur, p, bx, by);
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
        const size_t pb  = (size_t)p * G::FLAT;
        const int    x0 = 8 * bx, y0 = 8 * by;

        // ---- boundary-face role of this lane for the task: where the ghost cell of plane z
        // comes from, resolved ONCE per task into a compact descriptor (the halo tables are not
        // touched inside the plane loop)
        const int  bd       = (side < 2) ? 4 + side : side; // tree direction of the side
        const bool internal = (side == 0)   ? (bx > 0)
                              : (side == 1) ? (bx < C::NBX - 1)
                              : (side == 2) ? (by > 0)
                                            : (by < C::NBY - 1);
        int g_y, g_x, i_y, i_x; // padded (y, x) of the ghost / interior cell of the face
        if (side < 2)
        {
            g_y = i_y = H + y0 + bt;
            g_x       = side ? H + x0 + 8 : H + x0 - 1;
            i_x       = side ? H + x0 + 7 : H + x0;
        }
        else
        {
            g_x = i_x = H + x0 + bt;
            g_y       = (side == 3) ? H + y0 + 8 : H + y0 - 1;
            i_y       = (side == 3) ? H + y0 + 7 : H + y0;
        }
        const int ioff = i_y * P + i_x;
        GhostSrc3 gs;
        {
            // relation byte and the four neighbor indices of the side, from the prefetched tables
            const int bm = tab_meta3(tab, p, bd);
            int4      bnb;
            bnb.x = __shfl_sync(0xffffffffu, tab, bd * 4);
            bnb.y = __shfl_sync(0xffffffffu, tab, bd * 4 + 1);
            bnb.z = __shfl_sync(0xffffffffu, tab, bd * 4 + 2);
            bnb.w = __shfl_sync(0xffffffffu, tab, bd * 4 + 3);
            const int  rel = (!internal && a.lazy_halo) ? (bm & 3) : 0;
            int f_y = g_y, f_x = g_x; // mirrored into the neighbor's frame (patch_utils.hpp:322-327)
            if (side < 2)
                f_x += (side & 1) ? -S : S;
            else
                f_y += (side & 1) ? -S : S;
            gs.finer  = 0;
            gs.pair   = 0;
            gs.zshift = 0;
            gs.zbase  = H;
            if (rel == 0)
            {
                // own patch: block side inside the patch, relation 'none', or materialised halos
                gs.q0 = gs.q1 = p;
                gs.off        = g_y * P + g_x;
            }
            else if (rel == 1)
            {
                gs.q0 = gs.q1 = bnb.x; // same_t (patch_utils.hpp:315-332)
                gs.off        = f_y * P + f_x;
            }
            else if (rel == 3)
            {
                // coarser_t: injection of the covering coarse cell (patch_utils.hpp:388-441)
                const int qz = (bm >> 2) & 1, qy = (bm >> 3) & 1, qx = (bm >> 4) & 1;
                gs.q0 = gs.q1 = bnb.x;
                gs.off = (H + qy * (S / 2) + (f_y - H) / 2) * P + (H + qx * (S / 2) + (f_x - H) / 2);
                gs.zbase  = H + qz * (S / 2);
                gs.zshift = 1;
            }
            else
            {
                // finer_t: mean of 8 fine cells (patch_utils.hpp:334-386); finer-neighbor index =
                // z half (bit 0) + 2 x half along the other tangential dim (neighbor.hpp:316-337)
                const int t = ((side < 2) ? (g_y - H) : (g_x - H)) / (S / 2);
                gs.q0       = t ? bnb.z : bnb.x;
                gs.q1       = t ? bnb.w : bnb.y;
                gs.off      = ((((f_y - H) * 2) % S) + H) * P + (((f_x - H) * 2) % S) + H;
                gs.finer    = 1;
            }
        }
        // issue the loads of this lane's boundary face of plane z (ghost + interior cell, all
        // fields) into staging buffer z & 1; they land asynchronously (cp.async group)
        auto bnd_issue = [&](int z) {
            double* st = sST;
            if (gs.finer)
            {
                const size_t o = (size_t)(z < S / 2 ? gs.q0 : gs.q1) * G::FLAT +
                                 (size_t)(H + 2 * (z % (S / 2))) * PP + gs.off;
                double t[NV];
                fine_mean5(a.cur, o, P, PP, t);
#pragma unroll
                for (int f = 0; f < NV; ++f) st[f * 32] = t[f];
            }
            else
            {
                const size_t o =
                    (size_t)gs.q0 * G::FLAT + (size_t)(gs.zbase + (z >> gs.zshift)) * PP + gs.off;
#pragma unroll
                for (int f = 0; f < NV; ++f) cp_async8(st + f * 32, a.cur.p[f] + o);
            }
            const size_t zo = pb + (size_t)(H + z) * PP + ioff;
#pragma unroll
            for (int f = 0; f < NV; ++f) cp_async8(st + (NV + f) * 32, a.cur.p[f] + zo);
        };
        // flux of this lane's boundary face of plane z from the staged cells -> sBF[z & 1].
        // G = F(ghost) + F(interior) -/+ smax (U_interior - U_ghost): the Rusanov flux is symmetric
        // in the two cells except for the sign of the dissipation term (low sides: ghost is the
        // left cell), so one branch-free evaluation serves all four sides.
        const double bsgn = (side & 1) ? -1.0 : 1.0;
        auto bnd_flux = [&](int z) {
            const double* st = sST;
            Cell3         gc, ic_;
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                gc.u[f]  = st[f * 32];
                ic_.u[f] = st[(NV + f) * 32];
            }
            prims3(gc, g, gm1);
            prims3(ic_, g, gm1);
            double F[NV];
            flux3_bnd(gc, ic_, side >> 1, bsgn, F);
            double2* o = reinterpret_cast<double2*>(sBF + ((z & 1) * 32 + lane) * C::BFW);
            o[0]       = make_double2(F[0], F[1]);
            o[1]       = make_double2(F[2], F[3]);
            o[2]       = make_double2(F[4], 0.0);
        };
        // ghost cells of this lane's column pair across the z faces (d = 0 below, 1 above)
        auto zghost = [&](int d, double (&vA)[NV], double (&vB)[NV]) {
            const int m = tab_meta3(tab, p, d);
            int4      nb;
            nb.x = __shfl_sync(0xffffffffu, tab, d * 4);
            nb.y = __shfl_sync(0xffffffffu, tab, d * 4 + 1);
            nb.z = __shfl_sync(0xffffffffu, tab, d * 4 + 2);
            nb.w = __shfl_sync(0xffffffffu, tab, d * 4 + 3);
            const int rel = a.lazy_halo ? (m & 3) : 0;
            const int      y = H + y0 + yy, x = H + x0 + 2 * xq;
            const int      zi = d ? H + S : H - 1, zf = d ? H : H + S - 1; // ghost plane, mirrored
            int            q = p, dB = 1, fin = 0;
            int            off = zi * PP + y * P + x;
            if (rel == 1)
            {
                q   = nb.x;
                off = zf * PP + y * P + x;
            }
            else if (rel == 3)
            {
                const int qz = (m >> 2) & 1, qy = (m >> 3) & 1, qx = (m >> 4) & 1;
                q   = nb.x;
                off = (H + qz * (S / 2) + (zf - H) / 2) * PP + (H + qy * (S / 2) + (y - H) / 2) * P +
                      (H + qx * (S / 2) + (x - H) / 2);
                dB = 0;
            }
            else if (rel == 2)
            {
                const int fi = (y - H) / (S / 2) + 2 * ((x - H) / (S / 2));
                q            = (fi == 0) ? nb.x : (fi == 1) ? nb.y : (fi == 2) ? nb.z : nb.w;
                off = ((((zf - H) * 2) % S) + H) * PP + ((((y - H) * 2) % S) + H) * P +
                      (((x - H) * 2) % S) + H;
                dB  = 2;
                fin = 1;
            }
            const size_t o = (size_t)q * G::FLAT + off;
            if (fin)
            {
                double tA[NV], tB[NV];
                fine_mean5(a.cur, o, P, PP, tA);
                fine_mean5(a.cur, o + 2, P, PP, tB);
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
                    vA[f] = __ldg(a.cur.p[f] + o);
                    vB[f] = __ldg(a.cur.p[f] + o + dB);
                }
            }
        };

        // ---- task prologue: boundary fluxes of plane 0 and the ghost plane below
        double gzA[NV], gzB[NV];
        bnd_issue(0);
        cp_async_commit();
        zghost(0, gzA, gzB);
        cp_async_wait<0>();
        bnd_flux(0);
        __syncwarp();

        struct PlaneState
        {
            Cell3  A, B;               // records of the lane's two cells
            double accA[NV], accB[NV]; // U + hx dFx + hy dFy - hz Fz(low)
        };
        PlaneState s0, s1;
        // pair (padded x = x0 + 2 xq, +1) = (left cell | A) of the row being finished
        size_t go = pb + (size_t)H * PP + (size_t)(H + y0 + yy) * P + x0 + 2 * xq;
        int    sl = 0; // slot of the streamed plane inside its chunk

        // finish a plane: add the upper z-face flux, store, wave speeds.  Whole padded rows (and
        // the ghost rows of the plane) are stored; lane (y, xq) stores the 16-byte aligned pair
        // (B of the left lane | A).
        const bool edge_row = (yy == 0) || (yy == 7); // also writes the ghost row next to it
        const int  edge_off = (yy == 0) ? -P : P;
        auto finish = [&](const PlaneState& pv, const double (&GzA)[NV], const double (&GzB)[NV],
                          bool fin) {
            double rA[NV], rB[NV];
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                rA[f]        = fma(hz, GzA[f], pv.accA[f]);
                rB[f]        = fma(hz, GzB[f], pv.accB[f]);
                const double lft = __shfl_up_sync(0xffffffffu, rB[f], 1);
                double*      rowp = a.nxt.p[f] + go;
                if constexpr (C::NB == 1)
                {
                    // pair (left | A); the row's first lane writes the ghost column as a copy
                    const double2 v = make_double2(xq == 0 ? rA[f] : lft, rA[f]);
                    const double2 w = make_double2(rB[f], rB[f]);
                    if (fin)
                    {
                        *reinterpret_cast<double2*>(rowp) = v;
                        if (xq == 3) *reinterpret_cast<double2*>(rowp + 2) = w;
                        if (edge_row)
                        {
                            *reinterpret_cast<double2*>(rowp + edge_off) = v;
                            if (xq == 3) *reinterpret_cast<double2*>(rowp + edge_off + 2) = w;
                        }
                    }
                }
                else
                {
                    auto put = [&](double* q) {
                        if (xq == 0)
                        {
                            if (bx == 0)
                                *reinterpret_cast<double2*>(q) = make_double2(rA[f], rA[f]);
                            else
                                q[1] = rA[f];
                        }
                        else
                            *reinterpret_cast<double2*>(q) = make_double2(lft, rA[f]);
                        if (xq == 3)
                        {
                            if (bx == C::NBX - 1)
                                *reinterpret_cast<double2*>(q + 2) = make_double2(rB[f], rB[f]);
                            else
                                q[2] = rB[f];
                        }
                    };
                    if (fin)
                    {
                        put(rowp);
                        if (yy == 0 && by == 0) put(rowp - P);
                        if (yy == 7 && by == C::NBY - 1) put(rowp + P);
                    }
                }
            }
            if (fin) go += PP;
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
                sxm             = fin ? pos_max(sxm, fabs(n[1] * irho) + cs) : sxm;
                sym             = fin ? pos_max(sym, fabs(n[2] * irho) + cs) : sym;
                szm             = fin ? pos_max(szm, fabs(n[3] * irho) + cs) : szm;
            }
        };

        // one interior plane z: `pv` = state of plane z-1 (or of the ghost plane below), `nw` =
        // state of plane z.  fin: plane z-1 exists and is finished here; last: z == S-1 (the
        // ghost plane above is fetched instead of the next boundary faces).  ONE instance of this
        // body exists in the kernel (rolled loop): it has to stay inside the instruction cache.
        auto plane_step = [&](int z, const PlaneState& pv, PlaneState& nw, bool fin, bool last) {
            const int BUF = z & 1;
            // inputs of the next plane's boundary faces: requested now, used at the end of this
            // step.  (One plane of lead, not more: the neighbor patches are streamed by warps
            // running in lock-step with this one, a later request finds their planes in L2.)
            if (!last)
            {
                bnd_issue(z + 1);
                cp_async_commit();
            }
            else
                zghost(1, gzA, gzB); // in flight during the last plane
            if (sl == 0) mbar_wait(&bar[cst], cph);
            const double* src = ring + cst * C::STAGE + sl * PLD + (C::ROW0 + yy) * P + x0 + 2 * xq;
            double        Lu[NV];
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                const double2 v0 = *reinterpret_cast<const double2*>(src + f * FS);
                Lu[f]            = v0.x;
                nw.A.u[f]        = v0.y;
                nw.B.u[f]        = src[f * FS + 2];
            }
            prims3(nw.A, g, gm1);
            prims3(nw.B, g, gm1);
            double GzA[NV], GzB[NV];
            flux3<2>(pv.A, nw.A, GzA);
            flux3<2>(pv.B, nw.B, GzB);
            finish(pv, GzA, GzB, fin);
            const double* bfp = sBF + BUF * 32 * C::BFW;
            // ---- x faces: L|A from the left lane's B record, A|B local, B|R = right lane's
            // L|A; the row's first / last lane take the parked boundary fluxes
            {
                const double2* bf =
                    reinterpret_cast<const double2*>(bfp + ((xq >> 1) * 8 + yy) * C::BFW);
                const double2 b0 = bf[0], b1 = bf[1], b2 = bf[2];
                const double  bfl[NV] = { b0.x, b0.y, b1.x, b1.y, b2.x };
                Cell3         L;
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
                    nw.accA[f] = fma(hx, GM[f] - GL[f], nw.A.u[f]);
                    nw.accB[f] = fma(hx, GR - GM[f], nw.B.u[f]);
                }
            }
            // ---- y faces: lower face from the record of lane-4 (U from the staged plane),
            // upper face = lane+4's lower face; first / last row take the parked fluxes
            {
                const int yl = (C::ROW0 + yy > 0) ? -P : 0; // staged row below (clamped)
                Cell3     YA, YB;
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    const double2 v0 = *reinterpret_cast<const double2*>(src + f * FS + yl);
                    YA.u[f]          = v0.y;
                    YB.u[f]          = src[f * FS + yl + 2];
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
                const double2* bf = reinterpret_cast<const double2*>(
                    bfp + ((2 + (yy >> 2)) * 8 + 2 * xq) * C::BFW);
                const double2 c0 = bf[0], c1 = bf[1], c2 = bf[2], d0 = bf[3], d1 = bf[4], d2 = bf[5];
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
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                nw.accA[f] = fma(nhz, GzA[f], nw.accA[f]);
                nw.accB[f] = fma(nhz, GzB[f], nw.accB[f]);
            }
            if (!last)
            {
                cp_async_wait<0>();
                bnd_flux(z + 1);
            }
            __syncwarp();
            if (++sl == CR)
            {
                // every lane has consumed its values of this stage's last plane (see the 2D
                // kernel): the stage can be refilled
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
#pragma unroll 1
        for (int z = 0; z < S; ++z)
        {
            if (z == 2)
            {
                // the next task is known by now: request its halo tables
                resolve_next();
                if (tau_nxt < n_tasks) tab_nxt = tab_load(tau_nxt);
            }
            plane_step(z, s0, s1, z > 0, z == S - 1);
            s0 = s1;
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
        resolve_next();
        tau_cur = tau_nxt;
        tab     = tab_nxt;
        if (tau_cur < n_tasks) fetch_next();
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
        if (lane == 0 && kc > 0)
            atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
    }
}

} // namespace amrb
