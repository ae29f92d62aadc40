// Fused advection step over the interior-only 3D pool layout (amrb_layout.storage = AMRB_STORAGE_INTERIOR).
//
// Scalar advection moves 16 algorithmic bytes per cell update and ~20 flops: a pure streaming kernel.
//   * AdvectionPhysics<3> has velocity {1, 0.5, 0} (include/solver/AdvectionPhysics.hpp:24): no motion
//     along z, the z-face fluxes are exactly zero -> planes are independent.  task = 512 cells = ZT whole
//     planes of one patch (an 8^3 patch is one task, a 16^3 patch eight), one 4 KB TMA bulk copy
//     (cp.async.bulk + mbarrier complete_tx) into a warp-private double buffer; a warp walks its tasks
//     back to back and never meets a block barrier; the copy of task k+1 is in flight while task k is
//     computed.
//   * the 4 x S x ZT lateral ghost cells of task k+1 are gathered from the neighbor patch INTERIORS through
//     the halo tables (same / coarser injection: per-lane 8-byte cp.async, no register staging; finer:
//     the 8-cell mean in the reference's summation order) into a second double buffer while task k is
//     computed; ghost cells are never written to the pool.
//   * a lane owns two x-adjacent cells; the x-neighbors are the adjacent lanes' cells (warp shuffles), the
//     y-neighbors come from the staged plane; stores are 16-byte pairs, 512 contiguous bytes per warp
//     instruction, every sector fully written.
//   * UPWIND form of the Rusanov flux: for the reference's constant velocity v = {1, 0.5, 0} >= 0
//     (AdvectionPhysics.hpp:24, 45-66)  F = 1/2 (v uL + v uR) - 1/2 |v| (uR - uL) = v uL: the downwind cell
//     cancels.  The kernel evaluates v uL directly -- the result differs from the textbook grouping by the rounding
//     of that cancellation (~1e-16 of the field, parity bound 1e-12) -- and therefore never reads the x+ / y+
//     ghost cells: half of the lateral gathers, each of which is one 8-byte element per cache line on the x
//     sides (profiles/r02_summary.md: those gathers, not the 4 KB stream, are what the kernel waits for).
// Arithmetic: AdvectionPhysics.hpp:45-66 (Rusanov), amr_solver.hpp:265-353 (update order).
#pragma once
#include "amrb_march_euler3d.cuh"

namespace amrb
{

template <int S, int WPC>
struct Adv3DenseCfg
{
    static constexpr int SS    = S * S;
    static constexpr int N     = S * S * S;
    static constexpr int TASK  = 512;            // cells per task
    static constexpr int ZT    = TASK / SS;      // planes per task
    static constexpr int NB    = S / ZT;         // tasks per patch
    static constexpr int HP    = S / 2;          // pairs per row
    static constexpr int GH    = 2 * S * ZT;     // upwind ghost cells per task: [side x- / y-][plane][tangential]
    static constexpr int WARP_DOUBLES = 2 * TASK + 2 * GH;
    static constexpr size_t SMEM      = (size_t)WPC * WARP_DOUBLES * sizeof(double);
    static_assert(S == 8 || S == 16, "512-cell tasks of whole planes");
};

template <int S, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
advect3d_dense_kernel(const __grid_constant__ StepArgs a, int n_items)
{
    using C          = Adv3DenseCfg<S, WPC>;
    constexpr int SS = C::SS, N = C::N, ZT = C::ZT, HP = C::HP, GH = C::GH, HF = S / 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WPC * 2];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double*   ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * C::WARP_DOUBLES;
    double*   sG   = ring + 2 * C::TASK;
    uint64_t* bar  = bars + warp * 2;
    const double* __restrict__ cur = a.cur.p[0];
    double* __restrict__       nxt = a.nxt.p[0];

    const int n_tasks = n_items * C::NB;
    const int gw = blockIdx.x * WPC + warp, nw_all = gridDim.x * WPC;
    const int nt = (n_tasks > gw) ? (n_tasks - gw + nw_all - 1) / nw_all : 0;

    auto task_of = [&](int k, int& p, int& z0) {
        const int tau  = gw + k * nw_all;
        const int item = tau / C::NB;
        z0             = (tau % C::NB) * ZT;
        p              = a.list ? a.list[item] : item;
    };

    if (lane == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncwarp();

    // halo tables of a task's patch, one 32-bit piece per lane (lanes 0-23: the 6 x 4 neighbor indices, lane 24:
    // level, lanes 25-30: the 6 relation bytes), loaded TWO tasks ahead and handed round by shuffles: looked up
    // per ghost cell they were a chain of three dependent global loads (relation, neighbor index, value) in
    // front of every gather -- ~10 k cycles per 512-cell task, the whole kernel (profiles/r02_summary.md)
    auto tab_load = [&](int k) -> int {
        if (k >= nt) return 0;
        int p, z0;
        task_of(k, p, z0);
        return tab_piece3(a.nbr, a.level, a.meta, p, lane);
    };
    // request task k: the bulk copy of its planes and the gather of its lateral ghost cells (tab = its tables)
    auto request = [&](int k, int tab) {
        if (k >= nt) return;
        int p, z0;
        task_of(k, p, z0);
        const int b = k & 1;
        if (lane == 0)
        {
            mbar_expect_tx(&bar[b], C::TASK * 8);
            bulk_g2s(ring + b * C::TASK, cur + (size_t)p * N + (size_t)z0 * SS, C::TASK * 8, &bar[b]);
        }
        double* gdst = sG + b * GH;
#pragma unroll
        for (int e = lane; e < GH; e += 32)
        {
            const int sd = e / (S * ZT), r = e % (S * ZT), zl = r / S, t = r % S; // sd: 0 = x-, 1 = y-
            const int d  = sd ? 2 : 4;                  // tree direction of the side
            const int m  = tab_meta3(tab, p, d);
            const int rel = m & 3;
            const int z   = z0 + zl;
            // neighbor index this ghost cell reads: the first one, or for a finer neighbor the one covering it
            const int nsel = (rel == 2) ? (z / HF) + 2 * (t / HF) : 0;
            const int nq   = __shfl_sync(0xffffffffu, tab, d * 4 + nsel);
            // interior coordinates of the ghost cell mirrored into the neighbor's frame
            // (patch_utils.hpp:322-327): the normal coordinate -1 -> S-1, S -> 0
            const int fy = sd ? S - 1 : t, fx = sd ? t : S - 1;
            if (rel == 2)
            {
                // finer_t: mean of the 2^3 fine cells of one of the 4 finer neighbors, summed
                // last-dim-fastest (patch_utils.hpp:334-386, 203-234); finer index = z half + 2 x half of
                // the other tangential dim (neighbor.hpp:316-337)
                const double* s = cur + (size_t)nq * N + (size_t)((z * 2) % S) * SS + ((fy * 2) % S) * S +
                                  ((fx * 2) % S);
                double sum = 0.0;
                sum += __ldg(s);
                sum += __ldg(s + 1);
                sum += __ldg(s + S);
                sum += __ldg(s + S + 1);
                sum += __ldg(s + SS);
                sum += __ldg(s + SS + 1);
                sum += __ldg(s + SS + S);
                sum += __ldg(s + SS + S + 1);
                gdst[e] = sum / 8.0;
            }
            else
            {
                // same_t: the mirrored cell; coarser_t: injection of the covering coarse cell
                // (patch_utils.hpp:315-332, 388-441); relation "none" (never in a periodic balanced tree): the
                // own boundary cell.  One select per coordinate instead of a branch per relation.
                const int  qz = (m >> 2) & 1, qy = (m >> 3) & 1, qx = (m >> 4) & 1;
                const bool co = (rel == 3), none = (rel == 0);
                const int  iy = sd ? 0 : t;
                const int  ix = sd ? t : 0;
                const int  sz = co ? qz * HF + z / 2 : z;
                const int  sy = co ? qy * HF + fy / 2 : (none ? iy : fy);
                const int  sx = co ? qx * HF + fx / 2 : (none ? ix : fx);
                const size_t o = (size_t)(none ? p : nq) * N + (size_t)sz * SS + sy * S + sx;
                // a y- ghost row of a same-level neighbor is 8 contiguous cells: two per copy (the load / store
                // unit works per REQUEST: scattered 8-byte copies, not bytes, are what these gathers cost)
                if (sd && !co)
                {
                    if ((t & 1) == 0) cp_async16(gdst + e, cur + o);
                }
                else
                    cp_async8(gdst + e, cur + o);
            }
        }
        cp_async_commit();
    };

    double       rem_after;
    const double dt = resolve_step_dt(a.sc, rem_after);
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.sc.dtmin_in != nullptr)
    {
        *a.sc.dt_taken      = dt;
        *a.sc.remaining_out = rem_after;
    }
    double cand = DBL_MAX;

    int tab_cur = tab_load(0), tab_nxt = tab_load(1);
    request(0, tab_cur);
    for (int k = 0; k < nt; ++k)
    {
        request(k + 1, tab_nxt); // stage (k+1)&1 was released at the end of task k-1
        if (k + 1 >= nt) cp_async_commit(); // keep one group per iteration: wait_group<1> below
        const int tab_nn = tab_load(k + 2); // in flight during this task
        int p, z0;
        task_of(k, p, z0);
        const int    b   = k & 1;
        const int    lvl = __shfl_sync(0xffffffffu, tab_cur, 24);
        const double cx = dt / a.dx[lvl][0], cy = dt / a.dx[lvl][1]; // amr_solver.hpp:330
        // CFL of the next step: speeds are state-independent (AdvectionPhysics.hpp:74-85), z has speed 0
        cand = fmin(cand, fmin(a.dx[lvl][0] / 1.0, a.dx[lvl][1] / 0.5));
        mbar_wait(&bar[b], (k >> 1) & 1);
        cp_async_wait<1>(); // this task's ghosts have landed (the younger group is task k+1's)
        __syncwarp();
        const double* pl = ring + b * C::TASK;
        const double* gh = sG + b * GH;
        double*       out = nxt + (size_t)p * N + (size_t)z0 * SS;
        // branch-free loop body: the ghost cells are ALWAYS loaded (valid addresses for every lane) and merged
        // with selects, the y-neighbors come through a selected pointer -- the divergent if / else regions of the
        // first version were 23 % of all issued instructions (profiles/r02i_advect3d_dense_ncu_summary.txt)
#pragma unroll
        for (int it = 0; it < C::TASK / 64; ++it)
        {
            const int q  = it * 32 + lane; // pair index inside the task
            const int x2 = q % HP, y = (q / HP) % S, zl = q / (HP * S);
            const int o  = zl * SS + y * S + 2 * x2;
            const double2 c = *reinterpret_cast<const double2*>(pl + o);
            const double  gl  = gh[(0 * ZT + zl) * S + y];
            const double  sl  = __shfl_up_sync(0xffffffffu, c.y, 1);
            const double  lft = (x2 == 0) ? gl : sl;
            const double* pdn = (y > 0) ? pl + o - S : gh + (1 * ZT + zl) * S + 2 * x2;
            const double2 dn  = *reinterpret_cast<const double2*>(pdn);
            // U - cx (F(U) - F(left)) - cy (G(U) - G(below)),  F = 1.0 u, G = 0.5 u  (update order of amr_solver.hpp:330)
            constexpr double vx = 1.0, vy = 0.5;
            double2          r;
            {
                double upd = 0.0;
                upd -= cx * (vx * c.x - vx * lft);
                upd -= cy * (vy * c.x - vy * dn.x);
                r.x = c.x + upd;
            }
            {
                double upd = 0.0;
                upd -= cx * (vx * c.y - vx * c.x);
                upd -= cy * (vy * c.y - vy * dn.y);
                r.y = c.y + upd;
            }
            *reinterpret_cast<double2*>(out + o) = r;
        }
        __syncwarp(); // every lane is done with stage b and ghost buffer b: task k+2 may overwrite them
        tab_cur = tab_nxt;
        tab_nxt = tab_nn;
    }
    cp_async_wait<0>();

    if (a.sc.dtmin_out != nullptr)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0 && nt > 0) atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
    }
}

} // namespace amrb
