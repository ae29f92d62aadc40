// Fused advection step for 2D patches (padded reference layout): warp-autonomous streaming.
//
// Scalar advection moves 16 algorithmic bytes per cell update: a pure streaming kernel, and the equation of
// the reference's literal example (examples/fvm_solver_advection.e.cpp: 10 x 10 patches, halo 2).
//   * task = a group of TP Morton-consecutive small patches (whole padded patches: 4 x 1 568 B for 10 x 10 /
//     halo 2) or a band of 8 rows (+ one row below and above) of one wide patch; staged by TMA 1-D bulk
//     copies (cp.async.bulk + mbarrier complete_tx; padded patches and row bands are contiguous) into a
//     warp-private double buffer.  A warp walks its tasks back to back and never meets a block barrier; the
//     copies of task k+1 are in flight while task k is computed.  The thread-per-cell step_kernel ran one
//     CTA per 100-cell patch and was launch / occupancy bound there (0.09 of the HBM roofline).
//   * the face ghosts the stencil reads (ONE layer, also for halo 2: amr_solver.hpp:317-321) are gathered
//     from the neighbor patch INTERIORS through the halo tables (halo_source: same / coarser injection /
//     finer restriction in the reference's summation order) for task k+1 while task k is computed, kept in
//     registers across the compute loop, parked in shared memory and dropped into the staged tile when it
//     has landed; ghost cells are never written to the pool by this kernel.
//   * stores cover whole padded rows of the interior rows (ghost columns receive a copy of the adjacent
//     cell): a patch's interior rows are one contiguous run of fully written sectors (tools/store_bench.cu).
// Arithmetic: AdvectionPhysics.hpp:45-66 (Rusanov), amr_solver.hpp:265-353 (update order), expression for
// expression the thread-per-cell step_kernel (amrb_kernels.cuh).
#pragma once
#include "amrb_march_euler3d.cuh"

namespace amrb
{

template <int S, int H, int WPC, int BRB = 8>
struct Adv2Cfg
{
    using G                    = Geo<2, S, H>;
    static constexpr int P     = G::P;
    static constexpr int FLAT  = G::FLAT;
    static constexpr bool WHOLE = (FLAT * 8 <= 4096);                       // whole padded patches per task
    static constexpr int TP    = WHOLE ? ((6400 / (FLAT * 8)) > 0 ? (6400 / (FLAT * 8)) : 1) : 1;
    static constexpr int BR    = WHOLE ? S : BRB;                           // interior rows per task
    static constexpr int NB    = S / BR;                                    // tasks per patch (band mode)
    static constexpr int NR    = WHOLE ? P : BR + 2;                        // staged rows per patch
    static constexpr int ROW0  = WHOLE ? H : 1;                             // staged row of the task's first interior row
    static constexpr int PST   = NR * P;                                    // staged doubles per patch
    static constexpr int STAGE = TP * PST;
    static constexpr int GP    = 2 * BR + 2 * S;                            // ghost cells per patch: x-, x+, y-, y+
    static constexpr int GH    = TP * GP;
    static constexpr int NG    = (GH + 31) / 32;                            // ghost cells per lane
    static constexpr int STAGE_PAD = (STAGE + 1) & ~1;
    static constexpr int GH_PAD    = (GH + 1) & ~1;
    static constexpr int WARP_DOUBLES = 2 * STAGE_PAD + 2 * GH_PAD;
    static constexpr size_t SMEM      = (size_t)WPC * WARP_DOUBLES * sizeof(double);
    static_assert(S % BR == 0, "band shape");
    static_assert((PST * 8) % 16 == 0 && (P * 8) % 16 == 0, "bulk copy size / alignment (padded extents are even)");
};

template <int S, int H, int WPC, int MINB, int BRB = 8>
__global__ void __launch_bounds__(WPC * 32, MINB)
advect2d_kernel(const __grid_constant__ StepArgs a, int n_items)
{
    using C          = Adv2Cfg<S, H, WPC, BRB>;
    using G          = Geo<2, S, H>;
    constexpr int P = C::P, FLAT = C::FLAT, TP = C::TP, BR = C::BR, PST = C::PST, GP = C::GP;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WPC * 2];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double*   ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::WARP_DOUBLES;
    double*   sG   = ring + 2 * C::STAGE_PAD;
    uint64_t* bar  = bars + warp * 2;
    const double* __restrict__ cur = a.cur.p[0];
    double* __restrict__       nxt = a.nxt.p[0];

    // tasks: whole mode = groups of TP consecutive items; band mode = (item, band)
    const int n_tasks = C::WHOLE ? (n_items + TP - 1) / TP : n_items * C::NB;
    const int gw = blockIdx.x * WPC + warp, nw_all = gridDim.x * WPC;
    const int nt = (n_tasks > gw) ? (n_tasks - gw + nw_all - 1) / nw_all : 0;

    // task k of this warp: first item, number of patches, first interior row of the band
    auto task_of = [&](int k, int& item0, int& np, int& r0) {
        const int tau = gw + k * nw_all;
        if constexpr (C::WHOLE)
        {
            item0 = tau * TP;
            np    = min(TP, n_items - item0);
            r0    = 0;
        }
        else
        {
            item0 = tau / C::NB;
            np    = 1;
            r0    = (tau % C::NB) * BR;
        }
    };
    auto patch_of = [&](int item) { return a.list ? a.list[item] : item; };

    if (lane == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncwarp();

    // bulk copies of task k into stage k & 1: per patch one contiguous run of NR padded rows
    auto issue = [&](int k) {
        if (k >= nt || lane != 0) return;
        int item0, np, r0;
        task_of(k, item0, np, r0);
        const int b = k & 1;
        mbar_expect_tx(&bar[b], (uint32_t)(np * PST * 8));
        for (int j = 0; j < np; ++j)
        {
            const size_t go = (size_t)patch_of(item0 + j) * FLAT + (size_t)(C::WHOLE ? 0 : (H + r0 - 1) * P);
            bulk_g2s(ring + b * C::STAGE_PAD + j * PST, cur + go, PST * 8, &bar[b]);
        }
    };
    // ghost cell g of a task: patch j, side (0 x-, 1 x+, 2 y-, 3 y+), position t
    auto ghost_decode = [&](int g, int& j, int& sd, int& t) {
        j           = g / GP;
        const int r = g % GP;
        if (r < 2 * BR)
        {
            sd = r / BR;
            t  = r % BR;
        }
        else
        {
            sd = 2 + (r - 2 * BR) / S;
            t  = (r - 2 * BR) % S;
        }
    };
    // halo tables of a task's patches, 10 ints per patch (8 neighbor indices = 4 directions x 2, the 4 relation
    // bytes as one word, the level), spread over the lanes (NT3 registers per lane), loaded TWO tasks ahead and
    // handed round by shuffles.  Looked up per ghost cell they were a chain of three dependent global loads
    // (relation, neighbor index, value) in front of every gather: ~10 k cycles per 512-cell task.
    constexpr int NT3 = (TP * 10 + 31) / 32;
    struct Tabs
    {
        int r[NT3];
    };
    auto tab_load = [&](int k) -> Tabs {
        Tabs t;
#pragma unroll
        for (int i = 0; i < NT3; ++i) t.r[i] = 0;
        if (k >= nt) return t;
        int item0, np, r0;
        task_of(k, item0, np, r0);
#pragma unroll
        for (int i = 0; i < NT3; ++i)
        {
            const int w = lane + 32 * i, j = w / 10, c = w % 10;
            const int p = patch_of(item0 + ((j < np) ? j : 0));
            t.r[i]      = tab_piece2(a.nbr, a.level, a.meta, p, c, j < np);
        }
        return t;
    };
    auto tab_get = [&](const Tabs& t, int w) -> int {
        int v = __shfl_sync(0xffffffffu, t.r[0], w & 31);
#pragma unroll
        for (int i = 1; i < NT3; ++i)
        {
            const int u = __shfl_sync(0xffffffffu, t.r[i], w & 31);
            v           = ((w >> 5) == i) ? u : v;
        }
        return v;
    };
    // gather the ghost cells of task k into registers (default 0, skipped when unused); tb = its tables
    auto gather = [&](int k, const Tabs& tb, double (&gv)[C::NG]) {
#pragma unroll
        for (int i = 0; i < C::NG; ++i) gv[i] = 0.0;
        if (k >= nt || !a.lazy_halo) return;
        int item0, np, r0;
        task_of(k, item0, np, r0);
#pragma unroll
        for (int i = 0; i < C::NG; ++i)
        {
            const int g = lane + 32 * i;
            int       j, sd, t;
            ghost_decode(g < C::GH ? g : 0, j, sd, t);
            const int d = (sd < 2) ? 2 + sd : sd - 2;  // tree direction: x- = 2, x+ = 3, y- = 0, y+ = 1
            // all lanes take part in the table shuffles; lanes without a ghost cell skip afterwards
            const int m = (tab_get(tb, j * 10 + 8) >> (8 * d)) & 0xff;
            NbRegs<2> nq;
            nq.v[0] = tab_get(tb, j * 10 + 2 * d);
            nq.v[1] = tab_get(tb, j * 10 + 2 * d + 1);
            if (g >= np * GP) continue;
            if (sd & 1) continue;                      // x+ / y+: never read (upwind form, see the compute loop)
            if (sd == 2 && r0 != 0) continue;          // the band does not touch the bottom face
            if ((m & 3) == 0) continue;
            int idx[2];
            if (sd < 2)
            {
                idx[0] = H + r0 + t;
                idx[1] = sd ? H + S : H - 1;
            }
            else
            {
                idx[0] = (sd == 3) ? H + S : H - 1;
                idx[1] = H + t;
            }
            gv[i] = halo_source<2, S, H, H, NbRegs<2>>(cur, nq, m, d, idx);
        }
    };

    double       rem_after;
    const double dt = resolve_step_dt(a.sc, rem_after);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.sc.dtmin_in != nullptr)
    {
        *a.sc.dt_taken      = dt;
        *a.sc.remaining_out = rem_after;
    }
    double cand = DBL_MAX;
    double gv[C::NG];

    Tabs tab_cur = tab_load(0), tab_nxt = tab_load(1);
    issue(0);
    gather(0, tab_cur, gv);
#pragma unroll
    for (int i = 0; i < C::NG; ++i)
        if (lane + 32 * i < C::GH) sG[lane + 32 * i] = gv[i];
    __syncwarp();

    for (int k = 0; k < nt; ++k)
    {
        issue(k + 1);   // stage (k+1)&1 was released at the end of task k-1
        gather(k + 1, tab_nxt, gv);
        const Tabs tab_nn = tab_load(k + 2); // in flight during this task
        int item0, np, r0;
        task_of(k, item0, np, r0);
        const int b  = k & 1;
        double*   st = ring + b * C::STAGE_PAD;
        mbar_wait(&bar[b], (k >> 1) & 1);
        // drop the parked ghost cells of this task into the staged tiles (relation "none" / trusted halos:
        // the stored ghost stays)
        if (a.lazy_halo)
        {
            const double* gh = sG + b * C::GH_PAD;
#pragma unroll
            for (int i = 0; i < C::NG; ++i)
            {
                const int g = lane + 32 * i;
                int       j, sd, t;
                ghost_decode(g < C::GH ? g : 0, j, sd, t);
                const int d = (sd < 2) ? 2 + sd : sd - 2;
                const int m = (tab_get(tab_cur, j * 10 + 8) >> (8 * d)) & 0xff;
                if (g >= np * GP) continue;
                if (sd & 1) continue;
                if (sd == 2 && r0 != 0) continue;
                if ((m & 3) == 0) continue;
                int o;
                if (sd < 2)
                    o = (C::ROW0 + t) * P + (sd ? H + S : H - 1);
                else
                    o = (sd == 3 ? C::ROW0 + BR : C::ROW0 - 1) * P + H + t;
                st[j * PST + o] = gh[g];
            }
        }
        __syncwarp();
        // ---- compute: every element of the padded interior rows (ghost columns = copy of the adjacent cell).
        // UPWIND form of the Rusanov flux: for the reference's constant velocity v = {1, 0.5} >= 0
        // (AdvectionPhysics.hpp:24, 45-66)  F = 1/2 (v uL + v uR) - 1/2 |v| (uR - uL) = v uL; the kernel evaluates
        // v uL directly (differs from the textbook grouping by the rounding of that cancellation, ~1e-16 of the
        // field; parity bound 1e-12) and never reads the x+ / y+ ghost cells: half of the gathers.
        for (int j = 0; j < np; ++j)
        {
            const int    p   = patch_of(item0 + j);
            const int    lvl = tab_get(tab_cur, j * 10 + 9);
            const double cx = dt / a.dx[lvl][0], cy = dt / a.dx[lvl][1]; // amr_solver.hpp:330
            cand = fmin(cand, fmin(a.dx[lvl][0] / 1.0, a.dx[lvl][1] / 0.5));
            const double* tile = st + j * PST;
            double*       out  = nxt + (size_t)p * FLAT + (size_t)(H + r0) * P;
            if constexpr (H == 1)
            {
                // a lane owns two x-adjacent cells: the 16-byte aligned pairs (2i, 2i+1) of the padded row.
                // First pair = (ghost | first cell), last pair = (last cell | ghost): the ghost gets a copy.
                constexpr int HP = P / 2;
#pragma unroll 2
                for (int q = lane; q < BR * HP; q += 32)
                {
                    const int  r = q / HP, pc = q % HP, c = 2 * pc;
                    const int  o = (C::ROW0 + r) * P + c;
                    const bool first = (pc == 0), last = (pc == HP - 1);
                    const double2 ct = *reinterpret_cast<const double2*>(tile + o);
                    const double2 dn = *reinterpret_cast<const double2*>(tile + o - P);
                    const double  lf = tile[first ? o : o - 1];
                    constexpr double vx = 1.0, vy = 0.5;
                    double nv[2];
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2)
                    {
                        const double u = s2 ? ct.y : ct.x, uL = s2 ? ct.x : lf, uD = s2 ? dn.y : dn.x;
                        double       upd = 0.0;
                        upd -= cx * (vx * u - vx * uL);
                        upd -= cy * (vy * u - vy * uD);
                        nv[s2] = u + upd;
                    }
                    *reinterpret_cast<double2*>(out + r * P + c) = make_double2(first ? nv[1] : nv[0], last ? nv[0] : nv[1]);
                }
            }
            else
            {
                int r = lane / P, c = lane % P; // element (row, padded column) walked incrementally
                for (int e = lane; e < BR * P; e += 32)
                {
                    const int    cc = min(max(c, H), H + S - 1);
                    const int    o  = (C::ROW0 + r) * P + cc;
                    constexpr double vx = 1.0, vy = 0.5;
                    const double     u = tile[o];
                    double           upd = 0.0;
                    upd -= cx * (vx * u - vx * tile[o - 1]);
                    upd -= cy * (vy * u - vy * tile[o - P]);
                    out[e] = u + upd;
                    r += 32 / P;
                    c += 32 % P;
                    if (c >= P)
                    {
                        c -= P;
                        ++r;
                    }
                }
            }
        }
        // park the ghost cells of the next task (their loads were in flight during the compute loop)
        {
            double* gn = sG + ((k + 1) & 1) * C::GH_PAD;
#pragma unroll
            for (int i = 0; i < C::NG; ++i)
                if (lane + 32 * i < C::GH) gn[lane + 32 * i] = gv[i];
        }
        __syncwarp(); // every lane is done with stage b: task k+2 may overwrite it
        tab_cur = tab_nxt;
        tab_nxt = tab_nn;
    }

    if (a.sc.dtmin_out != nullptr)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0 && nt > 0) atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
    }
}

} // namespace amrb
