// Fused Euler step, second generation: software-pipelined multi-tile CTAs.
//
// One CTA processes TPC consecutive tiles (tile = BAND slowest-dim rows of one patch + one ghost
// row each side, all fields) through a 2-stage shared-memory ring filled by TMA 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx): the copy of tile t+1 / t+2 is in flight while tile t is
// computed.  The table-driven ghost gather of tile t+1 is issued into registers before tile t is
// computed and committed to shared memory when the tile's bulk copy has landed, so its three
// dependent global loads (relation, neighbor index, value) are off the critical path as well.
//
// Per tile:  pass A  primitives once per cell (p, a, 1/rho) into shared memory
//            pass B  each thread marches a 2-cell-wide column strip along the slowest dim:
//                    the marching-direction face flux is computed once per face and carried in
//                    registers, the three x-faces of the pair are computed once each from the
//                    four x-consecutive records L|A B|R fetched with two aligned LDS.128 per plane
//            epilogue  wave speed of the new state -> CFL minimum of the next step
//
// Arithmetic follows include/solver/EulerPhysics.hpp:74-129 / amr_solver.hpp:265-353 of the
// reference operation by operation; the common factor 0.5 of the Rusanov flux is folded into
// dt/dx, which is exact in binary floating point.
#pragma once
#include "amrb_kernels.cuh"

#include <type_traits>

namespace amrb
{

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Branch-free double-precision reciprocal and square root: hardware seed (MUFU.RCP64H /
// MUFU.RSQ64H, ~20 bits) + Newton-Raphson in FMA arithmetic.  Results are within 1 ulp of the
// correctly rounded value (the reference's IEEE `/` and std::sqrt), far inside the 1e-12 parity
// tolerance, and avoid the exponent checks / slow-path calls of div.rn.f64 and sqrt.rn.f64.
// Valid for positive normal arguments (densities, gamma p / rho).
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r        = fma(r, e, r);
    e        = fma(-x, r, 1.0);
    r        = fma(r, e, r);
    e        = fma(-x, r, 1.0);
    r        = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double fast_sqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g        = fma(g, r, g);
    h        = fma(h, r, h);
    r        = fma(-g, h, 0.5);
    g        = fma(g, r, g);
    h        = fma(h, r, h);
    r        = fma(-g, g, x); // final residual correction: g += (x - g^2) * h
    g        = fma(r, h, g);
    return g;
}

template <int R>
struct Rec
{
    double u[R + 2];
    double p, a, ir;
};

template <int R, int TILE>
__device__ __forceinline__ Rec<R> ld_rec(const double* __restrict__ sU,
                                         const double* __restrict__ sW, int o)
{
    Rec<R> r;
#pragma unroll
    for (int f = 0; f < R + 2; ++f) r.u[f] = sU[f * TILE + o];
    r.p  = sW[o];
    r.a  = sW[TILE + o];
    r.ir = sW[2 * TILE + o];
    return r;
}

// Records of the four x-consecutive cells L | A B | R around the pair (A, B); `o` = tile offset of A.
// With an odd halo width A sits at an odd double index, so (L, A) and (B, R) are two 16-byte
// aligned pairs: 2 conflict-free LDS.128 per plane, lanes reading consecutive 16-byte chunks.
template <int R>
struct Quad
{
    Rec<R> L, A, B, Rr;
};

template <int H>
__device__ __forceinline__ void ld4(const double* __restrict__ plane, int o, double& l, double& a,
                                    double& b, double& r)
{
    if constexpr (H & 1)
    {
        const double2 v0 = *reinterpret_cast<const double2*>(plane + o - 1);
        const double2 v1 = *reinterpret_cast<const double2*>(plane + o + 1);
        l = v0.x;
        a = v0.y;
        b = v1.x;
        r = v1.y;
    }
    else
    {
        const double2 v = *reinterpret_cast<const double2*>(plane + o);
        a = v.x;
        b = v.y;
        l = plane[o - 1];
        r = plane[o + 2];
    }
}

template <int R, int H, int TILE>
__device__ __forceinline__ Quad<R> ld_quad(const double* __restrict__ sU,
                                           const double* __restrict__ sW, int o)
{
    Quad<R> q;
#pragma unroll
    for (int f = 0; f < R + 2; ++f) ld4<H>(sU + f * TILE, o, q.L.u[f], q.A.u[f], q.B.u[f], q.Rr.u[f]);
    ld4<H>(sW, o, q.L.p, q.A.p, q.B.p, q.Rr.p);
    ld4<H>(sW + TILE, o, q.L.a, q.A.a, q.B.a, q.Rr.a);
    ld4<H>(sW + 2 * TILE, o, q.L.ir, q.A.ir, q.B.ir, q.Rr.ir);
    return q;
}

// G = F(L) + F(R) - smax (U_R - U_L) across a face normal to solver direction DS
// (2 x the Rusanov flux of EulerPhysics.hpp:100-128)
template <int R, int DS>
__device__ __forceinline__ void face_flux(const Rec<R>& L, const Rec<R>& Rr, double (&G)[R + 2])
{
    const double uL = L.u[1 + DS] * L.ir, uR = Rr.u[1 + DS] * Rr.ir;
    const double sm = fmax(fabs(uL) + L.a, fabs(uR) + Rr.a);
    G[0]            = (L.u[1 + DS] + Rr.u[1 + DS]) - sm * (Rr.u[0] - L.u[0]);
#pragma unroll
    for (int k = 0; k < R; ++k)
    {
        double fl = L.u[1 + k] * uL, fr = Rr.u[1 + k] * uR;
        if (k == DS)
        {
            fl += L.p;
            fr += Rr.p;
        }
        G[1 + k] = (fl + fr) - sm * (Rr.u[1 + k] - L.u[1 + k]);
    }
    const double eL = uL * (L.u[R + 1] + L.p), eR = uR * (Rr.u[R + 1] + Rr.p);
    G[R + 1]        = (eL + eR) - sm * (Rr.u[R + 1] - L.u[R + 1]);
}

template <int R, int S, int H, int BAND, int RG, int TPC, int NT>
struct EulerStepCfg
{
    using G                     = Geo<R, S, H>;
    static constexpr int NV     = R + 2;
    static constexpr int NW     = 3;
    static constexpr int P0     = G::pitch(0);
    static constexpr int ROWS   = BAND + 2;
    static constexpr int TILE   = ROWS * P0;
    static constexpr int NBANDS = S / BAND;
    static constexpr size_t SMEM = (size_t)(2 * NV + NW) * TILE * sizeof(double);
    // ghost items per tile (inner layer only) and per thread
    static constexpr int SIDE   = (R == 2) ? BAND : BAND * S;
    static constexpr int GITEMS = 2 * G::FACE + (G::NDIR - 2) * SIDE;
    static constexpr int IPT    = (GITEMS + NT - 1) / NT;
    // marching items
    static constexpr int NQ     = (R == 2) ? S / 2 : S * (S / 2);
    static constexpr int NG     = BAND / RG;
    static constexpr int MITEMS = NG * NQ;
    static_assert(S % BAND == 0 && BAND % RG == 0 && S % 2 == 0, "tile shape");
};

template <int R, int S, int H, int BAND, int RG, int TPC, int NT, int PB = 0, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB)
euler_step_kernel(const __grid_constant__ StepArgs a, int n_items)
{
    using C            = EulerStepCfg<R, S, H, BAND, RG, TPC, NT>;
    using G            = Geo<R, S, H>;
    constexpr int NV   = C::NV;
    constexpr int P0   = C::P0;
    constexpr int TILE = C::TILE;
    constexpr int DM   = R - 1; // solver direction of the marching (slowest layout) dim

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sRing = reinterpret_cast<double*>(smem_raw);  // [2][NV][TILE]
    double* sW    = sRing + 2 * NV * TILE;                // [3][TILE]  p, a, 1/rho
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ double red[NT / 32];

    const int tid     = threadIdx.x;
    const int lane    = tid & 31;
    const int n_tiles = n_items * C::NBANDS;
    const int tile0   = blockIdx.x * TPC;
    const int ntile   = min(TPC, n_tiles - tile0);

    auto tile_patch = [&](int t, int& t0) -> int {
        const int tau  = tile0 + t;
        const int item = tau / C::NBANDS;
        t0             = (tau % C::NBANDS) * BAND;
        return a.list ? a.list[item] : item;
    };
    auto issue = [&](int t) {
        int       t0;
        const int p     = tile_patch(t, t0);
        double*   dst   = sRing + (t & 1) * NV * TILE;
        const size_t go = (size_t)p * G::FLAT + (size_t)(H + t0 - 1) * P0;
        mbar_expect_tx(&bar[t & 1], NV * TILE * 8);
#pragma unroll
        for (int f = 0; f < NV; ++f) bulk_g2s(dst + f * TILE, a.cur.p[f] + go, TILE * 8, &bar[t & 1]);
    };

    if (tid == 0)
    {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncthreads();
    if (tid == 0)
    {
        issue(0);
        if (ntile > 1) issue(1);
    }

    // ---- ghost prefetch registers
    double gv[C::IPT][NV];
    int    glo[C::IPT];
    auto   gather_issue = [&](int t) {
        int       t0;
        const int p    = tile_patch(t, t0);
        const int row0 = H + t0 - 1;
#pragma unroll
        for (int k = 0; k < C::IPT; ++k)
        {
            const int it = tid + k * NT;
            glo[k]       = -1;
            if (!a.lazy_halo || it >= C::GITEMS) continue;
            int d, idx[R];
            if (it < 2 * G::FACE)
            {
                d = it / G::FACE;
                if (d == 0 ? (t0 != 0) : (t0 + BAND != S)) continue;
                slab_index<R, S, H>(d, 0, it % G::FACE, idx);
            }
            else
            {
                const int r   = it - 2 * G::FACE;
                d             = 2 + r / C::SIDE;
                int       q   = r % C::SIDE;
                const int dim = d >> 1, pos = d & 1;
#pragma unroll
                for (int kk = R - 1; kk >= 1; --kk)
                {
                    if (kk == dim)
                        idx[kk] = pos ? (H + S) : (H - 1);
                    else
                    {
                        idx[kk] = H + (q % S);
                        q /= S;
                    }
                }
                idx[0] = H + t0 + q;
            }
            const int m = a.meta[(size_t)p * G::NDIR + d];
            if ((m & 3) == 0) continue;
            const int32_t* nb = a.nbr + ((size_t)p * G::NDIR + d) * G::KF;
            int            lo = (idx[0] - row0) * P0;
#pragma unroll
            for (int kk = 1; kk < R; ++kk) lo += idx[kk] * G::pitch(kk);
            glo[k] = lo;
#pragma unroll
            for (int f = 0; f < NV; ++f) gv[k][f] = halo_source<R, S, H>(a.cur.p[f], nb, m, d, idx);
        }
    };
    gather_issue(0);

    // ---- step scalars
    double       rem_after;
    const double dt = resolve_step_dt(a.sc, rem_after);
    if (blockIdx.x == 0 && tid == 0 && a.sc.dtmin_in != nullptr)
    {
        *a.sc.dt_taken      = dt;
        *a.sc.remaining_out = rem_after;
    }
    const double g    = a.gamma;
    double       cand = DBL_MAX; // per-thread min of dx/speed over finished patches
    double       smax[R];
#pragma unroll
    for (int ds = 0; ds < R; ++ds) smax[ds] = 0.0;
    int lvl_prev = -1;

    for (int t = 0; t < ntile; ++t)
    {
        double* sU = sRing + (t & 1) * NV * TILE;
        int     t0;
        const int p   = tile_patch(t, t0);
        const int lvl = a.level[p];
        if (lvl != lvl_prev && lvl_prev >= 0)
        {
            // fold the finished level's speeds: min dx/speed == dx / max speed (monotone)
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
            {
                if (smax[ds] > 1e-12) cand = fmin(cand, a.dx[lvl_prev][ds] / smax[ds]);
                smax[ds] = 0.0;
            }
        }
        lvl_prev = lvl;
        double hd[R]; // 0.5 * dt / dx per solver direction
#pragma unroll
        for (int ds = 0; ds < R; ++ds) hd[ds] = 0.5 * (dt / a.dx[lvl][ds]);

        mbar_wait(&bar[t & 1], (t >> 1) & 1);
#pragma unroll
        for (int k = 0; k < C::IPT; ++k)
            if (glo[k] >= 0)
            {
#pragma unroll
                for (int f = 0; f < NV; ++f) sU[f * TILE + glo[k]] = gv[k][f];
            }
        if (t + 1 < ntile) gather_issue(t + 1);
        __syncthreads();

        // ---- pass A: primitives once per tile cell (EulerPhysics.hpp:83-99)
        for (int c = tid; c < TILE; c += NT)
        {
            const double irho = fast_rcp(sU[c]);
            double       K    = 0.0;
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
            {
                const double m = sU[(1 + ds) * TILE + c];
                K += m * m;
            }
            K *= 0.5 * irho;
            const double pr  = (g - 1.0) * (sU[(R + 1) * TILE + c] - K);
            sW[c]            = pr;
            sW[TILE + c]     = fast_sqrt(g * pr * irho);
            sW[2 * TILE + c] = irho;
        }
        __syncthreads();

        if constexpr (PB == 1)
        {
            // ---- pass B (variant): one thread per cell, faces not shared between threads
            const size_t gb = (size_t)p * G::FLAT;
            constexpr int ROWCELL = G::DATA / S;
            for (int ci = tid; ci < BAND * ROWCELL; ci += NT)
            {
                int r = ci, lt = 0;
#pragma unroll
                for (int k = R - 1; k >= 1; --k)
                {
                    lt += (H + (r % S)) * G::pitch(k);
                    r /= S;
                }
                const size_t gl = gb + (size_t)(H + t0 + r) * P0 + lt;
                lt += (r + 1) * P0;
                const Rec<R> Cc = ld_rec<R, TILE>(sU, sW, lt);
                double       up[NV];
#pragma unroll
                for (int f = 0; f < NV; ++f) up[f] = 0.0;
                auto dir = [&](auto DSc) {
                    constexpr int DS = decltype(DSc)::value;
                    constexpr int st = G::pitch(R - 1 - DS);
                    const Rec<R>  Lr = ld_rec<R, TILE>(sU, sW, lt - st), Rr = ld_rec<R, TILE>(sU, sW, lt + st);
                    double        Gl[NV], Gh[NV];
                    face_flux<R, DS>(Lr, Cc, Gl);
                    face_flux<R, DS>(Cc, Rr, Gh);
#pragma unroll
                    for (int f = 0; f < NV; ++f) up[f] -= hd[DS] * (Gh[f] - Gl[f]);
                };
                dir(std::integral_constant<int, 0>{});
                dir(std::integral_constant<int, 1>{});
                if constexpr (R == 3) dir(std::integral_constant<int, 2>{});
                double n[NV];
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    n[f]            = Cc.u[f] + up[f];
                    a.nxt.p[f][gl] = n[f];
                }
                {
                    const int xi = ci % S;
                    if (xi == 0 || xi == S - 1)
                    {
                        const int dir = (xi == 0) ? -1 : 1;
#pragma unroll
                        for (int f = 0; f < NV; ++f)
#pragma unroll
                            for (int h = 1; h <= H; ++h) a.nxt.p[f][gl + dir * h] = n[f];
                    }
                }
                const double irho = fast_rcp(n[0]);
                double       K    = 0.0;
#pragma unroll
                for (int ds = 0; ds < R; ++ds) K += n[1 + ds] * n[1 + ds];
                K *= 0.5 * irho;
                const double pr = (g - 1.0) * (n[R + 1] - K);
                const double cs = fast_sqrt(g * pr * irho);
#pragma unroll
                for (int ds = 0; ds < R; ++ds) smax[ds] = fmax(smax[ds], fabs(n[1 + ds] * irho) + cs);
            }
        }
        else
        {
        // ---- pass B: march 2-wide column strips along the slowest dim
        const size_t gbase = (size_t)p * G::FLAT;
        for (int it = tid; it < C::MITEMS; it += NT)
        {
            const int grp = it / C::NQ;
            const int q   = it % C::NQ;
            int       cross; // offset of cell A inside a slowest-dim row
            const int xq = (R == 2) ? q : q % (S / 2); // index of the cell pair along x
            if constexpr (R == 2)
                cross = H + 2 * q;
            else
                cross = (H + q / (S / 2)) * G::P + H + 2 * (q % (S / 2));
            int     o   = (grp * RG) * P0 + cross; // ghost/previous row below the group's first row
            Quad<R> cur = ld_quad<R, H, TILE>(sU, sW, o);
            o += P0;
            Quad<R> nxt = ld_quad<R, H, TILE>(sU, sW, o);
            double  GmA[NV], GmB[NV];
            face_flux<R, DM>(cur.A, nxt.A, GmA);
            face_flux<R, DM>(cur.B, nxt.B, GmB);
            size_t go = gbase + (size_t)(H + t0 + grp * RG) * P0 + cross;
#pragma unroll
            for (int j = 0; j < RG; ++j)
            {
                cur = nxt;
                nxt = ld_quad<R, H, TILE>(sU, sW, o + P0);
                double GL[NV], GM[NV], GR[NV], uA[NV], uB[NV];
                face_flux<R, 0>(cur.L, cur.A, GL);
                face_flux<R, 0>(cur.A, cur.B, GM);
                face_flux<R, 0>(cur.B, cur.Rr, GR);
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    uA[f] = 0.0 - hd[0] * (GM[f] - GL[f]);
                    uB[f] = 0.0 - hd[0] * (GR[f] - GM[f]);
                }
                if constexpr (R == 3)
                {
                    double Gl[NV], Gh[NV];
                    {
                        const Rec<R> Yl = ld_rec<R, TILE>(sU, sW, o - G::P),
                                     Yh = ld_rec<R, TILE>(sU, sW, o + G::P);
                        face_flux<R, 1>(Yl, cur.A, Gl);
                        face_flux<R, 1>(cur.A, Yh, Gh);
#pragma unroll
                        for (int f = 0; f < NV; ++f) uA[f] -= hd[1] * (Gh[f] - Gl[f]);
                    }
                    {
                        const Rec<R> Yl = ld_rec<R, TILE>(sU, sW, o + 1 - G::P),
                                     Yh = ld_rec<R, TILE>(sU, sW, o + 1 + G::P);
                        face_flux<R, 1>(Yl, cur.B, Gl);
                        face_flux<R, 1>(cur.B, Yh, Gh);
#pragma unroll
                        for (int f = 0; f < NV; ++f) uB[f] -= hd[1] * (Gh[f] - Gl[f]);
                    }
                }
                double GuA[NV], GuB[NV];
                face_flux<R, DM>(cur.A, nxt.A, GuA);
                face_flux<R, DM>(cur.B, nxt.B, GuB);
                double nA[NV], nB[NV];
#pragma unroll
                for (int f = 0; f < NV; ++f)
                {
                    uA[f] -= hd[DM] * (GuA[f] - GmA[f]);
                    uB[f] -= hd[DM] * (GuB[f] - GmB[f]);
                    nA[f]  = cur.A.u[f] + uA[f];
                    nB[f]  = cur.B.u[f] + uB[f];
                    GmA[f] = GuA[f];
                    GmB[f] = GuB[f];
                    a.nxt.p[f][go]     = nA[f];
                    a.nxt.p[f][go + 1] = nB[f];
                }
                // whole padded x-rows are stored (ghost columns = copy of the adjacent cell): full
                // 32-byte sectors, see step_kernel
                if (xq == 0)
                {
#pragma unroll
                    for (int f = 0; f < NV; ++f)
#pragma unroll
                        for (int h = 1; h <= H; ++h) a.nxt.p[f][go - h] = nA[f];
                }
                if (xq == S / 2 - 1)
                {
#pragma unroll
                    for (int f = 0; f < NV; ++f)
#pragma unroll
                        for (int h = 1; h <= H; ++h) a.nxt.p[f][go + 1 + h] = nB[f];
                }
                // wave speed of the new state (EulerPhysics.hpp:137-161)
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2)
                {
                    const double* n    = c2 ? nB : nA;
                    const double  irho = fast_rcp(n[0]);
                    double        K    = 0.0;
#pragma unroll
                    for (int ds = 0; ds < R; ++ds) K += n[1 + ds] * n[1 + ds];
                    K *= 0.5 * irho;
                    const double pr = (g - 1.0) * (n[R + 1] - K);
                    const double cs = fast_sqrt(g * pr * irho);
#pragma unroll
                    for (int ds = 0; ds < R; ++ds)
                        smax[ds] = fmax(smax[ds], fabs(n[1 + ds] * irho) + cs);
                }
                o += P0;
                go += P0;
            }
        }
        }
        fence_proxy_async(); // order this thread's generic smem writes before the next bulk copy
        __syncthreads();
        if (tid == 0 && t + 2 < ntile) issue(t + 2);
    }

    if (a.sc.dtmin_out != nullptr)
    {
        if (lvl_prev >= 0)
        {
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
                if (smax[ds] > 1e-12) cand = fmin(cand, a.dx[lvl_prev][ds] / smax[ds]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fmin(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        if (lane == 0) red[tid >> 5] = cand;
        __syncthreads();
        if (tid == 0)
        {
            for (int i = 1; i < NT / 32; ++i) cand = fmin(cand, red[i]);
            atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
        }
    }
}

} // namespace amrb
