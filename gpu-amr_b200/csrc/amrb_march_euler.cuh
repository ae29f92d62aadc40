// Fused Euler step, third generation (2D): warp-autonomous marching pipelines.
//
// Every WARP owns a private pipeline and never meets a block barrier:
//   * task = one band of BAND rows of one 64-wide patch; a warp walks its tasks back to back
//   * the rows of a task (band + one row below and above, all fields) stream through a warp-private
//     shared-memory ring of NS stages filled by TMA 1-D bulk copies (cp.async.bulk + mbarrier
//     complete_tx; rows of a field-patch are contiguous, so one copy per field per chunk of CR rows);
//     the copies of the following chunks -- also those of the warp's NEXT task -- are in flight
//     while a chunk is computed
//   * a lane owns two x-adjacent cells and marches them along y: the state record of a cell
//     (U, p, a, 1/rho) is derived ONCE, in registers; the y-face flux is carried in registers, the
//     x-face between the lane's two cells is computed locally, the left x-face from the left
//     lane's pressure / sound speed / 1/rho (warp shuffle) and the right x-face flux is the right
//     lane's left flux (warp shuffle): every face flux is computed exactly once
//   * ghost cells are never written anywhere: the two patch-boundary x-face fluxes of every row are
//     computed at task start by the 32 lanes (lane = side x row) straight from the neighbor patch
//     interiors (halo tables: same / coarser injection / finer restriction) and parked in 1 KB of
//     shared memory; ghost rows across y-faces are gathered into the registers of the lanes that
//     march those columns
//   * epilogue per cell: wave speed of the new state, running maximum through the integer pipe
//     (bit patterns of non-negative doubles order like integers); one atomicMin per warp at the end
//
// Arithmetic follows include/solver/EulerPhysics.hpp:74-129 and amr_solver.hpp:265-353 of the
// reference; 0.5 of the Rusanov flux is folded into dt/dx (exact), the update is accumulated as
// (U - cx (Fx+ - Fx-)) - cy (Fy+ - Fy-) with fused multiply-adds (differs from the reference's
// U + (0 - .. - ..) by rounding only; parity bound 1e-12 field-max-normalised).
#pragma once
#include "amrb_step_euler.cuh"

namespace amrb
{

// reciprocal / square root: MUFU seed (20 mantissa bits of the operand) + 2 Newton steps: the
// error after two steps is O(2^-80), the result is within 1 ulp of the IEEE value.
__device__ __forceinline__ double rcp_nr2(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r        = fma(r, e, r);
    e        = fma(-x, r, 1.0);
    r        = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double sqrt_nr2(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g        = fma(g, r, g);
    h        = fma(h, r, h);
    r        = fma(-g, g, x); // residual form of the second step: g += (x - g^2) * h
    g        = fma(r, h, g);
    return g;
}
// max of two non-negative doubles on the integer pipe
__device__ __forceinline__ double pos_max(double a, double b)
{
    const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    return __longlong_as_double(x > y ? x : y);
}

struct Cell2
{
    double u[4];
    double p, a, ir;
};

__device__ __forceinline__ void prims2(Cell2& c, double g, double gm1)
{
    c.ir     = rcp_nr2(c.u[0]);
    double K = c.u[1] * c.u[1];
    K        = fma(c.u[2], c.u[2], K);
    K *= 0.5 * c.ir;
    c.p = gm1 * (c.u[3] - K);
    c.a = sqrt_nr2(g * c.p * c.ir);
}

// G = F(L) + F(R) - smax (U_R - U_L): twice the Rusanov flux (EulerPhysics.hpp:74-129) across a face normal to
// solver direction DS (0 = x: momentum u[1]; 1 = y: momentum u[2]), in WAVE FORM: with F_k = U_k u (+ p for the
// normal momentum, + p u for the energy) the sum regroups to
//     G_k = U_k,L (u_L + smax) + U_k,R (u_R - smax)   [+ p_L + p_R]   [+ p_L u_L + p_R u_R]
// -- 2 DP instructions per component instead of 5 (no per-cell flux vectors, no state differences): 15 per face
// in 2D, 19 in 3D, where the textbook grouping needs 23 / 29.  Same value up to rounding (one more rounding of
// rho u against m; parity bound 1e-12); a face's flux is still computed once and used by both cells, so the
// update telescopes as before.
template <int DS>
__device__ __forceinline__ void flux2(const Cell2& L, const Cell2& R, double (&G)[4])
{
    const double uL = L.u[1 + DS] * L.ir, uR = R.u[1 + DS] * R.ir;
    const double sm = pos_max(fabs(uL) + L.a, fabs(uR) + R.a);
    const double wp = uL + sm, wm = uR - sm;
    G[0]            = fma(R.u[0], wm, L.u[0] * wp);
    G[1 + DS]       = fma(R.u[1 + DS], wm, fma(L.u[1 + DS], wp, L.p + R.p));
    G[2 - DS]       = fma(R.u[2 - DS], wm, L.u[2 - DS] * wp);
    G[3]            = fma(R.u[3], wm, fma(L.u[3], wp, fma(L.p, uL, R.p * uR)));
}

template <int S, int H, int BAND, int CR, int NS, int WPC>
struct March2Cfg
{
    using G                    = Geo<2, S, H>;
    static constexpr int NV    = 4;
    static constexpr int P     = G::P;
    static constexpr int NB    = S / BAND;   // tasks per patch
    static constexpr int NR    = BAND + 2;   // streamed rows per task
    static constexpr int NCH   = NR / CR;    // chunks per task
    static constexpr int CHUNK = CR * P;     // doubles per field per chunk
    static constexpr int STAGE = NV * CHUNK; // doubles per ring stage
    static constexpr int RING  = NS * STAGE;
    static constexpr int BF    = BAND * 2 * NV; // boundary x-face fluxes of one task
    static constexpr int WARP_DOUBLES = RING + BF;
    static constexpr size_t SMEM      = (size_t)WPC * WARP_DOUBLES * sizeof(double);
    static_assert(S == 64, "one warp spans a 64-cell row (two cells per lane)");
    static_assert((H & 1) == 1, "odd ghost width: (ghost|first) and (second|right) pairs are 16-byte aligned");
    static_assert(S % BAND == 0 && NR % CR == 0, "band / chunk shape");
    static_assert((CHUNK * 8) % 16 == 0 && (WARP_DOUBLES * 8) % 16 == 0, "bulk copy alignment");
};

template <int S, int H, int BAND, int CR, int NS, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
euler2d_march_kernel(const __grid_constant__ StepArgs a, int n_items)
{
    using C          = March2Cfg<S, H, BAND, CR, NS, WPC>;
    using G          = Geo<2, S, H>;
    constexpr int NV = 4;
    constexpr int P  = C::P;
    constexpr int NR = C::NR;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WPC * NS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double*   ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::WARP_DOUBLES;
    double*   sBF  = ring + C::RING;
    uint64_t* bar  = bars + warp * NS;

    // ---- this warp's tasks: the CTA owns a contiguous task range, its warps interleave
    const int n_tasks = n_items * C::NB;
    const int tb      = (int)((long long)n_tasks * blockIdx.x / gridDim.x);
    const int te      = (int)((long long)n_tasks * (blockIdx.x + 1) / gridDim.x);
    const int gw = blockIdx.x * WPC + warp, nw_all = gridDim.x * WPC; // warp of the grid
    const int nt = a.task_map ? ((n_tasks > gw) ? (n_tasks - gw + nw_all - 1) / nw_all : 0)
                              : ((te - tb > warp) ? (te - tb - warp + WPC - 1) / WPC : 0);

    auto task_of = [&](int k, int& p, int& r0) {
        const int tau  = a.task_map ? gw + k * nw_all : tb + warp + k * WPC;
        const int item = tau / C::NB;
        r0             = (tau % C::NB) * BAND;
        p              = a.list ? a.list[item] : item;
    };

    if (lane == 0)
    {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    }
    __syncwarp();

    // ---- producer side of the ring (lane 0 issues; all lanes track the counters)
    int  ik = 0, ic = 0, ist = 0; // next chunk to issue: task, chunk in task, stage
    auto issue_next = [&]() {
        if (ik >= nt) return;
        int p, r0;
        task_of(ik, p, r0);
        if (lane == 0)
        {
            const size_t go  = (size_t)p * G::FLAT + (size_t)(H + r0 - 1 + ic * CR) * P;
            double*      dst = ring + ist * C::STAGE;
            mbar_expect_tx(&bar[ist], C::STAGE * 8);
#pragma unroll
            for (int f = 0; f < NV; ++f)
                bulk_g2s(dst + f * C::CHUNK, a.cur.p[f] + go, C::CHUNK * 8, &bar[ist]);
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
    double       sxm = 0.0, sym = 0.0;
    int          lvl_prev = -1;

    int cst = 0, cph = 0; // consumer: stage and phase parity of the next chunk to wait for

    // halo tables of a task's patch, one 32-bit piece per lane (lanes 0-7: the 4 x 2 neighbor indices,
    // lane 8: the 4 relation bytes, lane 9: level): loaded for the NEXT task while the current one is
    // marched and handed round by shuffles, so that no task starts with a table lookup
    auto tab_load = [&](int k) -> int {
        int q, r;
        task_of(k, q, r);
        return tab_piece2(a.nbr, a.level, a.meta, q, lane, true);
    };
    int tab = (nt > 0) ? tab_load(0) : 0;

    for (int k = 0; k < nt; ++k)
    {
        int p, r0;
        task_of(k, p, r0);
        const int      lvl = __shfl_sync(0xffffffffu, tab, 9);
        const uint32_t mb  = (uint32_t)__shfl_sync(0xffffffffu, tab, 8);
        int4           nA, nB;
        nA.x = __shfl_sync(0xffffffffu, tab, 0);
        nA.y = __shfl_sync(0xffffffffu, tab, 1);
        nA.z = __shfl_sync(0xffffffffu, tab, 2);
        nA.w = __shfl_sync(0xffffffffu, tab, 3);
        nB.x = __shfl_sync(0xffffffffu, tab, 4);
        nB.y = __shfl_sync(0xffffffffu, tab, 5);
        nB.z = __shfl_sync(0xffffffffu, tab, 6);
        nB.w = __shfl_sync(0xffffffffu, tab, 7);
        if (k + 1 < nt) tab = tab_load(k + 1);
        if (lvl != lvl_prev && lvl_prev >= 0)
        {
            if (sxm > 1e-12) cand = fmin(cand, a.dx[lvl_prev][0] / sxm);
            if (sym > 1e-12) cand = fmin(cand, a.dx[lvl_prev][1] / sym);
            sxm = 0.0;
            sym = 0.0;
        }
        lvl_prev          = lvl;
        const double hx   = -0.5 * (dt / a.dx[lvl][0]); // -0.5 dt/dx, x (fastest layout dim)
        const double hy   = -0.5 * (dt / a.dx[lvl][1]);
        const size_t pb   = (size_t)p * G::FLAT;
        const bool   bot  = (r0 == 0), top = (r0 + BAND == S);

        // (halo tables mb / nA / nB of the patch: resolved once per task above.  Looked up per
        // gather they made every ghost a chain of three dependent global loads -- relation, index,
        // value -- and 20 % of all stall samples.)

        // ---- ghost value of padded cell (row, col) across face d, or the stored ghost when the
        // tables say "none" / the caller asked to trust materialised halos.  Same index arithmetic
        // as halo_source<2, S, H> (amrb_kernels.cuh), resolved once for all fields.
        auto ghost = [&](int d, int row, int col, double (&v)[NV]) {
            const int m   = (int)((mb >> (8 * d)) & 0xffu);
            const int rel = a.lazy_halo ? (m & 3) : 0;
            const int n0  = (d == 0) ? nA.x : (d == 1) ? nA.z : (d == 2) ? nB.x : nB.z;
            const int n1  = (d == 0) ? nA.y : (d == 1) ? nA.w : (d == 2) ? nB.y : nB.w;
            int       fr = row, fc = col; // mirrored into the neighbor's frame
            if ((d >> 1) == 0)
                fr += (d & 1) ? -S : S;
            else
                fc += (d & 1) ? -S : S;
            size_t o   = pb + (size_t)(row * P + col);
            bool   fin = false;
            if (rel == 1)
                o = (size_t)n0 * G::FLAT + (size_t)(fr * P + fc);
            else if (rel == 3)
                o = (size_t)n0 * G::FLAT +
                    (size_t)((H + ((m >> 2) & 1) * (S / 2) + (fr - H) / 2) * P +
                             (H + ((m >> 3) & 1) * (S / 2) + (fc - H) / 2));
            else if (rel == 2)
            {
                const int t = (((d >> 1) == 0) ? (col - H) : (row - H)) / (S / 2);
                o   = (size_t)(t ? n1 : n0) * G::FLAT +
                    (size_t)(((((fr - H) * 2) % S) + H) * P + (((fc - H) * 2) % S) + H);
                fin = true;
            }
            if (!fin)
            {
#pragma unroll
                for (int f = 0; f < NV; ++f) v[f] = __ldg(a.cur.p[f] + o);
            }
            else
            {
                // restriction: mean of the 2 x 2 fine cells, summed last-dim-fastest
                // (patch_utils.hpp:203-234, intergrid_operator.hpp:92-106)
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    const double* sp  = a.cur.p[f] + o;
                    double        sum = 0.0;
                    sum += __ldg(sp);
                    sum += __ldg(sp + 1);
                    sum += __ldg(sp + P);
                    sum += __ldg(sp + P + 1);
                    v[f] = sum / 4.0;
                }
            }
        };

        // ---- patch-boundary x-face fluxes of the band's rows -> sBF[row][side][field]
        // All gathers of the task (boundary columns and the ghost row below) are issued before the
        // first value is used, so the task start pays one load latency, not one per item.
        constexpr int NIT = (2 * BAND + 31) / 32;
        double        gU[NIT][NV], iU[NIT][NV];
#pragma unroll
        for (int n = 0; n < NIT; ++n)
        {
            const int it = lane + 32 * n;
            if (it < 2 * BAND)
            {
                const int side = it / BAND, row = H + r0 + it % BAND;
                ghost(2 + side, row, side ? H + S : H - 1, gU[n]);
                const int coli = side ? H + S - 1 : H;
#pragma unroll
                for (int f = 0; f < NV; ++f) iU[n][f] = __ldg(a.cur.p[f] + pb + row * P + coli);
            }
        }
        // ---- ghost row below the patch, for the lanes' own columns
        double gy[2][NV];
        if (bot)
        {
            ghost(0, H - 1, H + 2 * lane, gy[0]);
            ghost(0, H - 1, H + 2 * lane + 1, gy[1]);
        }
#pragma unroll
        for (int n = 0; n < NIT; ++n)
        {
            const int it = lane + 32 * n;
            if (it < 2 * BAND)
            {
                const int side = it / BAND, i = it % BAND;
                Cell2     gc, ic_;
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    gc.u[f]  = gU[n][f];
                    ic_.u[f] = iU[n][f];
                }
                prims2(gc, g, gm1);
                prims2(ic_, g, gm1);
                double F[NV];
                if (side)
                    flux2<0>(ic_, gc, F);
                else
                    flux2<0>(gc, ic_, F);
                double2* o = reinterpret_cast<double2*>(sBF + (i * 2 + side) * NV);
                o[0]       = make_double2(F[0], F[1]);
                o[1]       = make_double2(F[2], F[3]);
            }
        }
        __syncwarp();

        // state carried from one streamed row to the next (two copies, used alternately, so that
        // the row loop needs no register moves)
        struct RowState
        {
            Cell2  A, B;           // records of the lane's two cells
            double nxA[NV], nxB[NV]; // U - cx (Fx+ - Fx-)
            double GmA[NV], GmB[NV]; // y-face flux below the row
        };
        RowState s0, s1;
        size_t   go = pb + (size_t)(H + r0) * P + 2 * lane; // (col 2l, 2l+1) of the row being finished
        int      sl = 0;                                        // slot of the streamed row inside its chunk

        // one streamed row j = 0 .. NR-1: `pv` = state of row j-1, `nw` = state of row j
        auto row_step = [&](int j, const RowState& pv, RowState& nw, auto DO_Y, auto DO_FIN, auto DO_X,
                            auto GHOST) {
            if (sl == 0) mbar_wait(&bar[cst], cph);
            const double* src = ring + cst * C::STAGE + sl * P + 2 * lane;
            double        Lu[NV];
#pragma unroll
            for (int f = 0; f < NV; ++f)
            {
                const double2 v0 = *reinterpret_cast<const double2*>(src + f * C::CHUNK);
                const double2 v1 = *reinterpret_cast<const double2*>(src + f * C::CHUNK + 2);
                Lu[f]            = v0.x;
                nw.A.u[f]        = v0.y;
                nw.B.u[f]        = v1.x;
            }
            if constexpr (decltype(GHOST)::value)
            {
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    nw.A.u[f] = gy[0][f];
                    nw.B.u[f] = gy[1][f];
                }
            }
            prims2(nw.A, g, gm1);
            prims2(nw.B, g, gm1);
            if constexpr (decltype(DO_Y)::value)
            {
                flux2<1>(pv.A, nw.A, nw.GmA);
                flux2<1>(pv.B, nw.B, nw.GmB);
            }
            if constexpr (decltype(DO_FIN)::value)
            {
                // finish the previous row: add the y-flux difference, store, wave speeds
                // The WHOLE padded row is stored (ghost columns get a copy of the adjacent cell):
                // rows are then contiguous and no 32-byte sector is left partially written --
                // partial sectors cost an L2 read-modify-write and halve the store bandwidth
                // (tools/store_bench.cu: 2.9 TB/s vs 5.6 TB/s).  The ghost columns of interior
                // rows are face ghosts: the next step gathers them from the neighbor interiors and
                // the halo kernel rewrites them before anything can observe them.
                // Lane l stores the 16-byte aligned pair (col 2l, col 2l+1) = (B of lane l-1, A).
                double rA[NV], rB[NV];
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    rA[f] = fma(hy, nw.GmA[f] - pv.GmA[f], pv.nxA[f]);
                    rB[f] = fma(hy, nw.GmB[f] - pv.GmB[f], pv.nxB[f]);
                    double lft = __shfl_up_sync(0xffffffffu, rB[f], 1);
                    if (lane == 0) lft = rA[f];
                    double* rowp = a.nxt.p[f] + go;
                    *reinterpret_cast<double2*>(rowp) = make_double2(lft, rA[f]);
                    if (lane == 31) *reinterpret_cast<double2*>(rowp + 2) = make_double2(rB[f], rB[f]);
                }
                go += P;
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2)
                {
                    const double* n    = c2 ? rB : rA;
                    const double  irho = rcp_nr2(n[0]);
                    double        K    = n[1] * n[1];
                    K                  = fma(n[2], n[2], K);
                    K *= 0.5 * irho;
                    const double pr = gm1 * (n[3] - K);
                    const double cs = sqrt_nr2(g * pr * irho);
                    sxm             = pos_max(sxm, fabs(n[1] * irho) + cs);
                    sym             = pos_max(sym, fabs(n[2] * irho) + cs);
                }
            }
            if constexpr (decltype(DO_X)::value)
            {
                // x-faces of this row: L|A from the left lane's B record, A|B local, B|R = right
                // lane's L|A; lanes 0 / 31 take the patch-boundary fluxes computed at task start
                // (every lane reads its half-warp's entry: a two-address broadcast, no branch)
                const double2* bf =
                    reinterpret_cast<const double2*>(sBF + ((j - 1) * 2 + (lane >> 4)) * NV);
                const double2 b0 = bf[0], b1 = bf[1];
                const double  bfl[NV] = { b0.x, b0.y, b1.x, b1.y };
                Cell2         L;
#pragma unroll
                for (int f = 0; f < NV; ++f) L.u[f] = Lu[f];
                L.p  = __shfl_up_sync(0xffffffffu, nw.B.p, 1);
                L.a  = __shfl_up_sync(0xffffffffu, nw.B.a, 1);
                L.ir = __shfl_up_sync(0xffffffffu, nw.B.ir, 1);
                double GL[NV], GM[NV];
                flux2<0>(L, nw.A, GL);
                flux2<0>(nw.A, nw.B, GM);
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    if (lane == 0) GL[f] = bfl[f];
                    double GR = __shfl_down_sync(0xffffffffu, GL[f], 1);
                    if (lane == 31) GR = bfl[f];
                    nw.nxA[f] = fma(hx, GM[f] - GL[f], nw.A.u[f]);
                    nw.nxB[f] = fma(hx, GR - GM[f], nw.B.u[f]);
                }
            }
            if (++sl == CR)
            {
                // every lane has consumed (not merely requested) its values of this stage's last
                // row: the stage can be refilled.  (A __syncwarp alone does not drain LDS in
                // flight; releasing before the values were used corrupted rows in testing.)
                sl = 0;
                __syncwarp();
                issue_next();
                if (++cst == NS)
                {
                    cst = 0;
                    cph ^= 1;
                }
            }
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
        static_assert((NR - 4) % 2 == 0, "interior rows are walked in pairs");
        if (bot)
            row_step(0, s1, s0, F_{}, F_{}, F_{}, T_{});
        else
            row_step(0, s1, s0, F_{}, F_{}, F_{}, F_{});
        row_step(1, s0, s1, T_{}, F_{}, T_{}, F_{});
#pragma unroll 1
        for (int j = 2; j < NR - 2; j += 2)
        {
            row_step(j, s1, s0, T_{}, T_{}, T_{}, F_{});
            row_step(j + 1, s0, s1, T_{}, T_{}, T_{}, F_{});
        }
        if (top)
        {
            // ghost row above the patch: issue the gather before the last interior row is computed
            ghost(1, H + S, H + 2 * lane, gy[0]);
            ghost(1, H + S, H + 2 * lane + 1, gy[1]);
        }
        row_step(NR - 2, s1, s0, T_{}, T_{}, T_{}, F_{});
        if (top)
            row_step(NR - 1, s0, s1, T_{}, T_{}, F_{}, T_{});
        else
            row_step(NR - 1, s0, s1, T_{}, T_{}, F_{}, F_{});
        __syncwarp(); // sBF is rewritten by the next task
    }

    if (a.sc.dtmin_out != nullptr)
    {
        if (lvl_prev >= 0)
        {
            if (sxm > 1e-12) cand = fmin(cand, a.dx[lvl_prev][0] / sxm);
            if (sym > 1e-12) cand = fmin(cand, a.dx[lvl_prev][1] / sym);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0 && nt > 0)
            atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
    }
}

} // namespace amrb
