// Device code of the B200-native gpu-amr hot path (sm_100a only).
//
//   halo_kernel        standalone ghost-cell fill, all fields in one launch
//   step_kernel        fused ghost gather (in shared memory) + Rusanov face flux + conservative
//                      update + CFL reduction for the NEXT step, one CTA per (patch, band)
//   compute_dt_kernel  standalone CFL reduction (first step of a batch)
//   plan_kernel        refine/coarsen data motion fused with the Morton re-sort
//   pack/unpack        inter-GPU ghost-face slabs
//
// Reference behaviour being reproduced (paths relative to the reference repository):
//   halo operators      include/ndtree/patch_utils.hpp:303-441  (same / finer / coarser)
//   restriction order   include/ndtree/patch_utils.hpp:203-234, intergrid_operator.hpp:92-106
//   update              include/solver/amr_solver.hpp:265-353
//   CFL                 include/solver/amr_solver.hpp:355-413
//   fluxes              include/solver/EulerPhysics.hpp:74-129, AdvectionPhysics.hpp:45-66
//   batch scalars       src/cuda/fvm_time_step.cu:204-233 (finalize_step_dt_kernel)
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace amrb
{

constexpr int kMaxVar   = 5;
constexpr int kMaxLevel = 24;
constexpr int kEqAdvection = 0;
constexpr int kEqEuler     = 1;

struct FieldPtrs
{
    double* p[kMaxVar];
};

// compile-time patch geometry: cubic patches of S^R interior cells with H ghost layers per side.
// HS = ghost layers that are STORED: HS == H is the reference's padded device layout; HS == 0 the
// interior-only layout of rank-3 pools (amrb_layout.storage = AMRB_STORAGE_INTERIOR), where a
// field-patch is S^R contiguous doubles and ghosts only ever exist inside the kernels.  Multi-indices
// are always LOGICAL padded coordinates (interior = [H, H+S)); storage offset = (i - H + HS) * pitch.
template <int R, int S, int H, int HS = H>
struct Geo
{
    static constexpr int rank = R;
    static constexpr int O    = HS;                           // storage coordinate of logical index H
    static constexpr int SH   = H - HS;                       // logical -> storage index shift
    static constexpr int P    = S + 2 * HS;                   // stored extent per dim
    static constexpr int FLAT = (R == 2) ? P * P : P * P * P; // doubles per field-patch
    static constexpr int DATA = (R == 2) ? S * S : S * S * S;
    static constexpr int NDIR = 2 * R;
    static constexpr int KF   = 1 << (R - 1); // finer neighbors per face
    static constexpr int FAN  = 1 << R;
    static constexpr int FACE = (R == 2) ? S : S * S; // cells per face layer
    // stride (in doubles) of layout dim k; last dim fastest
    __host__ __device__ static constexpr int pitch(int k)
    {
        return (k == R - 1) ? 1 : (k == R - 2) ? P : P * P;
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- TMA 1-D bulk copy (cp.async.bulk) + mbarrier: SASS UBLKCP / SYNCS ------------------------
// Predicated read-only loads for the per-lane pieces of a patch's halo tables (neighbor indices, level,
// relation bytes: three arrays, one piece per lane).  Written as `if (lane < 24) return ldg(..); if (lane == 24)
// return ldg(..); ...` the divergent paths run one after the other AND share their destination register, so
// each path waits for the previous path's load to land: a full memory latency per path (10 % of the warp time
// of advect3d_dense_kernel, profiles/r02i_advect3d_dense_ncu_summary.txt).  These issue back to back.
__device__ __forceinline__ int ldg_s32_if(const void* p, bool take, int otherwise)
{
    int v = otherwise;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.nc.s32 %0, [%1];\n\t}"
        : "+r"(v)
        : "l"(p), "r"((int)take));
    return v;
}
__device__ __forceinline__ int ldg_u8_if(const void* p, bool take, int otherwise)
{
    int v = otherwise;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.nc.u8 %0, [%1];\n\t}"
        : "+r"(v)
        : "l"(p), "r"((int)take));
    return v;
}
// 3D tables of patch q: lanes 0-23 the 6 x 4 neighbor indices, lane 24 the level, lanes 25-30 the aligned 32-bit
// word that holds the relation byte of direction lane-25 (ONE load instruction for all lanes; tab_meta3 picks the
// byte at the point of use -- the byte arrays are allocated with 4 spare bytes), lane 31 q itself
__device__ __forceinline__ int tab_piece3(const int32_t* nbr, const int32_t* level, const uint8_t* meta, int q,
                                          int lane)
{
    const size_t   mb  = ((size_t)q * 6 + (size_t)(lane >= 25 ? lane - 25 : 0)) & ~(size_t)3;
    const int32_t* p32 = (lane < 24)    ? nbr + (size_t)q * 24 + lane
                         : (lane == 24) ? level + q
                                        : reinterpret_cast<const int32_t*>(meta + mb);
    return ldg_s32_if(p32, lane < 31, q);
}
// relation byte (rel | quadrant << 2) of direction d of patch q out of the warp's table pieces
__device__ __forceinline__ int tab_meta3(int tab, int q, int d)
{
    const int w = __shfl_sync(0xffffffffu, tab, 25 + d);
    return (w >> ((((q & 1) * 2 + d) & 3) * 8)) & 0xff;
}
// 2D tables of patch q, piece c: 0-7 the 4 x 2 neighbor indices, 8 the 4 relation bytes, 9 the level
__device__ __forceinline__ int tab_piece2(const int32_t* nbr, const int32_t* level, const uint8_t* meta, int q,
                                          int c, bool take)
{
    const int32_t* p32 = (c < 8) ? nbr + (size_t)q * 8 + c
                                 : (c == 8 ? reinterpret_cast<const int32_t*>(meta) + q : level + q);
    return ldg_s32_if(p32, take && c < 10, 0);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- halo source evaluation ---------------------------------------------------------------------
// Value of ghost cell `idx` (padded multi-index, idx[dim] in the ghost range of direction d)
// gathered from the neighbor patch(es) across face d.   field = base pointer of one field's
// patch array.  nb = KF neighbor indices, meta = rel | quadrant bits << 2.
// neighbor indices of one (patch, direction) held in registers (kernels that prefetch their halo tables):
// indexable like the table row, without a local-memory array
template <int KF>
struct NbRegs
{
    int32_t v[KF];
    __device__ __forceinline__ int32_t operator[](int i) const
    {
        int32_t q = v[0];
#pragma unroll
        for (int k = 1; k < KF; ++k) q = (i == k) ? v[k] : q;
        return q;
    }
};

template <int R, int S, int H, int HS = H, typename NB = const int32_t*>
__device__ __forceinline__ double
halo_source(const double* __restrict__ field, NB nb, int meta, int d, const int (&idx)[R])
{
    using G        = Geo<R, S, H, HS>;
    const int dim  = d >> 1;
    const int pos  = d & 1;
    const int rel  = meta & 3;
    int       from[R];
#pragma unroll
    for (int k = 0; k < R; ++k) from[k] = idx[k];
    from[dim] += pos ? -S : S; // mirror into the neighbor's frame (patch_utils.hpp:322-327)
    if (rel == 1)
    {
        // same_t (patch_utils.hpp:315-332)
        int off = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) off += (from[k] - G::SH) * G::pitch(k);
        return __ldg(field + (size_t)nb[0] * G::FLAT + off);
    }
    if (rel == 3)
    {
        // coarser_t -> linear_interpolator::interpolation = injection of the covering coarse
        // cell (patch_utils.hpp:388-441, intergrid_operator.hpp:41-50)
        int off = 0;
#pragma unroll
        for (int k = 0; k < R; ++k)
        {
            const int q = (meta >> (2 + k)) & 1;
            off += (G::O + q * (S / 2) + (from[k] - H) / 2) * G::pitch(k);
        }
        return __ldg(field + (size_t)nb[0] * G::FLAT + off);
    }
    if (rel == 2)
    {
        // finer_t: mean of the 2^R covering fine cells of one of the 2^(R-1) finer neighbors
        // (patch_utils.hpp:334-386); summation order last-dim-fastest (hypercube_offset :203-234)
        int fine = 0, mul = 1, base = 0;
#pragma unroll
        for (int k = 0; k < R; ++k)
        {
            base += ((((from[k] - H) * 2) % S) + G::O) * G::pitch(k);
            if (k != dim)
            {
                fine += ((idx[k] - H) / (S / 2)) * mul;
                mul *= 2;
            }
        }
        const double* src = field + (size_t)nb[fine] * G::FLAT + base;
        double        sum = 0.0;
#pragma unroll
        for (int n = 0; n < G::FAN; ++n)
        {
            int off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) off += ((n >> (R - 1 - k)) & 1) * G::pitch(k);
            sum += __ldg(src + off);
        }
        return sum / (double)G::FAN;
    }
    return 0.0; // none: never read (boundary_t is a no-op, patch_utils.hpp:303-313)
}

// decode item t of the (layer l) ghost slab of direction d into a padded multi-index.
// tangential dims enumerate with the last layout dim fastest (coalesced for dim != R-1).
template <int R, int S, int H>
__device__ __forceinline__ void slab_index(int d, int layer, int t, int (&idx)[R])
{
    const int dim = d >> 1;
    const int pos = d & 1;
#pragma unroll
    for (int k = R - 1; k >= 0; --k)
    {
        if (k == dim)
            idx[k] = pos ? (H + S + layer) : (H - 1 - layer);
        else
        {
            idx[k] = H + (t % S);
            t /= S;
        }
    }
}

// ---- standalone halo fill: one CTA per patch, all directions, all layers, all fields --------------
template <int R, int S, int H, int NV>
__global__ void __launch_bounds__(256)
halo_kernel(FieldPtrs cur, const int32_t* __restrict__ nbr, const uint8_t* __restrict__ meta,
            int n_patches)
{
    using G     = Geo<R, S, H>;
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    constexpr int PER_DIR = H * G::FACE;
    constexpr int ITEMS   = G::NDIR * PER_DIR;
    for (int it = threadIdx.x; it < ITEMS; it += blockDim.x)
    {
        const int d     = it / PER_DIR;
        const int r     = it % PER_DIR;
        const int layer = r / G::FACE;
        const int t     = r % G::FACE;
        const int m     = meta[(size_t)p * G::NDIR + d];
        if ((m & 3) == 0) continue;
        const int32_t* nb = nbr + ((size_t)p * G::NDIR + d) * G::KF;
        int            idx[R];
        slab_index<R, S, H>(d, layer, t, idx);
        int to = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) to += idx[k] * G::pitch(k);
#pragma unroll
        for (int f = 0; f < NV; ++f)
        {
            const double v = halo_source<R, S, H>(cur.p[f], nb, m, d, idx);
            cur.p[f][(size_t)p * G::FLAT + to] = v;
        }
    }
}


// ---- neighbor / halo tables on the device ------------------------------------------
// One thread per (leaf, direction): the relation, neighbor indices and contact quadrant are derived
// from the ASCENDING array of leaf ids by key lookup (binary search), the same set-based rule as the
// host builder (amrb_topology.cpp: amrb_tree::neighbor; reference: ndtree/neighbor.hpp:291-572,
// ndtree.hpp:1606-1660).  id = morton(x, y[, z]) << 6 | level, x in bit 0 of every interleaved group.
__device__ __forceinline__ uint64_t topo_encode(int rank, const uint32_t (&c)[3], int level, int depth)
{
    uint64_t m = 0;
    for (int b = 0; b <= depth; ++b)
        for (int a = 0; a < rank; ++a) m |= (uint64_t)((c[a] >> b) & 1u) << (rank * b + a);
    return (m << 6) | (uint64_t)level;
}
__device__ __forceinline__ int topo_find(const uint64_t* __restrict__ ids, int n, uint64_t key)
{
    int lo = 0, hi = n;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (ids[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < n && ids[lo] == key) ? lo : -1;
}
static __global__ void __launch_bounds__(256)
topology_kernel(const uint64_t* __restrict__ ids, int n, int rank, int depth, int32_t* __restrict__ level,
                uint8_t* __restrict__ meta, int32_t* __restrict__ nbr)
{
    const int ND = 2 * rank, KF = 1 << (rank - 1);
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * ND) return;
    const int      i = (int)(t / ND), d = (int)(t % ND);
    const uint64_t id  = ids[i];
    const int      lvl = (int)(id & 63u);
    if (d == 0) level[i] = lvl;
    uint32_t c[3] = { 0u, 0u, 0u };
    {
        const uint64_t m = id >> 6;
        for (int b = 0; b <= depth; ++b)
            for (int a = 0; a < rank; ++a) c[a] |= (uint32_t)((m >> (rank * b + a)) & 1ull) << b;
    }
    const int      dim = d >> 1, pos = d & 1;
    const int      ax  = rank - 1 - dim; // layout dim k <-> morton axis rank-1-k
    const uint32_t span = 1u << depth, e = 1u << (depth - lvl);
    uint32_t       nn[3] = { c[0], c[1], c[2] };
    nn[ax] = (pos ? c[ax] + e : c[ax] + span - e) & (span - 1); // periodic domain
    int32_t out[4] = { -1, -1, -1, -1 };
    int     m      = 0; // relation | quadrant bits << 2
    int     li     = topo_find(ids, n, topo_encode(rank, nn, lvl, depth));
    if (li >= 0)
    {
        m      = 1; // same
        out[0] = li;
    }
    else
    {
        bool done = false;
        if (lvl > 0)
        {
            const uint32_t E = e << 1;
            uint32_t       cn[3];
            for (int a = 0; a < 3; ++a) cn[a] = nn[a] & ~(E - 1);
            li = topo_find(ids, n, topo_encode(rank, cn, lvl - 1, depth));
            if (li >= 0)
            {
                m      = 3; // coarser; contact quadrant: normal 0 for a positive direction, 1 for negative
                out[0] = li;
                for (int k = 0; k < rank; ++k)
                {
                    const int q = (k == dim) ? (pos ? 0 : 1) : (int)((c[rank - 1 - k] / e) & 1u);
                    m |= q << (2 + k);
                }
                done = true;
            }
        }
        if (!done && lvl < depth)
        {
            const uint32_t hh = e >> 1;
            bool           ok = true;
            for (int j = 0; j < KF && ok; ++j)
            {
                // finer-neighbor order: non-normal layout dims ascending, lowest = bit 0
                uint32_t cc[3] = { nn[0], nn[1], nn[2] };
                cc[ax] += pos ? 0u : hh;
                int bit = 0;
                for (int k = 0; k < rank; ++k)
                {
                    if (k == dim) continue;
                    if ((j >> bit) & 1) cc[rank - 1 - k] += hh;
                    ++bit;
                }
                li = topo_find(ids, n, topo_encode(rank, cc, lvl + 1, depth));
                if (li < 0)
                    ok = false;
                else
                    out[j] = li;
            }
            if (ok)
                m = 2; // finer
            else
                for (int j = 0; j < 4; ++j) out[j] = -1;
        }
    }
    meta[(size_t)i * ND + d] = (uint8_t)m;
    for (int k = 0; k < KF; ++k) nbr[((size_t)i * ND + d) * KF + k] = out[k];
}

// ---- batch scalars (device-resident dt bookkeeping) -----------------------------------------------
struct StepScalars
{
    // slot k of a batch: dtmin[k] = min dx/speed over the state entering step k (bits of a
    // positive double, atomicMin-able), remaining[k] = time left before step k,
    // dts[k] = step size actually taken.
    const unsigned long long* dtmin_in;
    unsigned long long*       dtmin_out;
    const double*             remaining_in;
    double*                   remaining_out;
    double*                   dt_taken;
    double                    fixed_dt; // used when dtmin_in == nullptr
    double                    cfl;
};

// finalize_step_dt_kernel semantics (src/cuda/fvm_time_step.cu:204-233), evaluated redundantly
// by every CTA of the step (all read the same two scalars)
__device__ __forceinline__ double resolve_step_dt(const StepScalars& sc, double& rem_after)
{
    if (sc.dtmin_in == nullptr)
    {
        rem_after = 0.0;
        return sc.fixed_dt;
    }
    const double raw = __longlong_as_double((long long)*sc.dtmin_in);
    const double rem = *sc.remaining_in;
    double       dt  = raw * sc.cfl;
    if (rem <= 0.0)
        dt = 0.0;
    else if (dt > rem)
        dt = rem;
    rem_after = (dt > 0.0) ? rem - dt : rem;
    return dt;
}

struct StepArgs
{
    FieldPtrs      cur;
    FieldPtrs      nxt;
    const int32_t* nbr;
    const uint8_t* meta;
    const int32_t* level;
    const int32_t* list; // optional patch sub-list
    int            n_patches;
    int            lazy_halo; // 1: gather ghosts from neighbor interiors; 0: trust global halos
    int            task_map;  // marching kernels: 1 = tasks interleaved over all warps of the grid
                              // (the chip sweeps one Morton window at a time: ghost gathers hit L2),
                              // 0 = contiguous task range per CTA
    unsigned int*  queue;       // optional dynamic task counter of THIS launch (3D marching kernel):
                                // a warp takes its next task when it finishes one, so warps slowed
                                // down by coarse/fine faces do not drift out of the common window;
                                // zeroed by the host in stream order before the launch
    int            variant;     // host side only: which instantiation the dispatcher launches
    double         gamma;
    double         dx[kMaxLevel + 1][3]; // per level, per solver direction (x,y,z)
    StepScalars    sc;
};

template <int EQ, int R>
struct EqTraits
{
    static constexpr int NV = (EQ == kEqAdvection) ? 1 : R + 2;
    static constexpr int NW = (EQ == kEqAdvection) ? 0 : R + 2; // derived arrays: p, a, u_d
};

// block-wide max of R values; result valid in thread 0
template <int R, int NT>
__device__ __forceinline__ void block_max(double (&v)[R], double* scratch /* [R][NT/32] */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < R; ++k)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] = fmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
        if (lane == 0) scratch[k * (NT / 32) + w] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < R; ++k)
        {
            double m = scratch[k * (NT / 32)];
            for (int i = 1; i < NT / 32; ++i) m = fmax(m, scratch[k * (NT / 32) + i]);
            v[k] = m;
        }
    }
}

// advection velocity (solver/AdvectionPhysics.hpp:24,54,82)
__device__ __forceinline__ double adv_vel(int ds) { return ds == 0 ? 1.0 : (ds == 1 ? 0.5 : 0.0); }

// ---- fused step ---------------------------------------------------------------------------------------
// One CTA per (patch, band of BAND slowest-dim rows).  Stage the band + one ghost row each side of every
// field with one cp.async.bulk per field (contiguous in the padded layout), overwrite the face
// ghosts in shared memory from the neighbor INTERIORS (halo sources are always interior cells, so
// the current buffer is read-only during the step), derive primitive quantities once per cell,
// then update.  The epilogue reduces max wave speed of the NEW state -> dt of the next step.
template <int R, int S, int H, int EQ, int BAND, int NT>
__global__ void __launch_bounds__(NT) step_kernel(const __grid_constant__ StepArgs a)
{
    using G               = Geo<R, S, H>;
    using E               = EqTraits<EQ, R>;
    constexpr int NV      = E::NV;
    [[maybe_unused]] constexpr int NW = E::NW;
    constexpr int P0      = G::pitch(0);      // doubles per slowest-dim row
    constexpr int ROWS    = BAND + 2;
    constexpr int TILE    = ROWS * P0;        // doubles per field tile
    constexpr int NBANDS  = S / BAND;
    constexpr int ROWCELL = G::DATA / S;      // interior cells per slowest-dim row
    static_assert(S % BAND == 0, "band must divide the patch");
    static_assert((TILE * 8) % 16 == 0, "bulk copy size must be a multiple of 16 bytes");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double*   sU = reinterpret_cast<double*>(smem_raw); // [NV][TILE]
    double*   sW = sU + NV * TILE;                      // [NW][TILE]  p, a, u_x, u_y(, u_z)
    __shared__ __align__(8) uint64_t bar;
    __shared__ double red[R * (NT / 32)];
    __shared__ int32_t sNbr[G::NDIR * G::KF]; // halo tables of the patch, fetched while the tile lands
    __shared__ int     sMeta[G::NDIR];

    const int item = blockIdx.x / NBANDS;
    const int band = blockIdx.x % NBANDS;
    const int p    = a.list ? a.list[item] : item;
    const int t0   = band * BAND;      // first interior row of the band (interior coordinates)
    const int row0 = H + t0 - 1;       // padded slowest-dim index of tile row 0
    const int tid  = threadIdx.x;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0)
    {
        if constexpr (EQ == kEqAdvection && R == 3)
        {
            // no z coupling (see the gather below): the ghost planes below / above the band are not
            // staged at all
            mbar_expect_tx(&bar, NV * BAND * P0 * 8);
            const size_t goff = (size_t)p * G::FLAT + (size_t)(row0 + 1) * P0;
#pragma unroll
            for (int f = 0; f < NV; ++f)
                bulk_g2s(sU + f * TILE + P0, a.cur.p[f] + goff, BAND * P0 * 8, &bar);
        }
        else
        {
            mbar_expect_tx(&bar, NV * TILE * 8);
            const size_t goff = (size_t)p * G::FLAT + (size_t)row0 * P0;
#pragma unroll
            for (int f = 0; f < NV; ++f) bulk_g2s(sU + f * TILE, a.cur.p[f] + goff, TILE * 8, &bar);
        }
    }

    // halo tables of the patch: independent loads issued with the copy in flight (looked up per ghost
    // item they were a chain of three dependent global loads: relation, neighbor index, value)
    if (tid < G::NDIR * G::KF) sNbr[tid] = a.nbr[(size_t)p * (G::NDIR * G::KF) + tid];
    if (tid >= 32 && tid < 32 + G::NDIR) sMeta[tid - 32] = a.meta[(size_t)p * G::NDIR + (tid - 32)];

    // scalar work overlapped with the copy
    double       rem_after;
    const double dt  = resolve_step_dt(a.sc, rem_after);
    const int    lvl = a.level[p];
    double       dtdx[R], dxs[R];
#pragma unroll
    for (int ds = 0; ds < R; ++ds)
    {
        dxs[ds]  = a.dx[lvl][ds];
        dtdx[ds] = dt / dxs[ds]; // amr_solver.hpp:330
    }
    if (blockIdx.x == 0 && tid == 0 && a.sc.dtmin_in != nullptr)
    {
        *a.sc.dt_taken      = dt;
        *a.sc.remaining_out = rem_after;
    }

    // ghost gather: prefetch sources into registers while the bulk copy is in flight? the gather
    // writes shared memory the copy also writes, so it must be ordered after the wait.
    mbar_wait(&bar, 0);
    __syncthreads(); // sNbr / sMeta visible

    if (a.lazy_halo)
    {
        // inner ghost layer only: the stencil never reads beyond one cell (amr_solver.hpp:317-321)
        constexpr int SIDE  = (R == 2) ? BAND : BAND * S; // ghost cells per non-slowest face in band
        constexpr int ITEMS = 2 * G::FACE + (G::NDIR - 2) * SIDE;
        // 3D advection has no motion along z (velocity {1, 0.5, 0}, AdvectionPhysics.hpp:24): its z-face
        // fluxes are exactly zero, so neither the z ghosts nor the z fluxes are evaluated
        constexpr int FIRST = (EQ == kEqAdvection && R == 3) ? 2 * G::FACE : 0;
        for (int it = FIRST + tid; it < ITEMS; it += NT)
        {
            int d, idx[R];
            if (it < 2 * G::FACE)
            {
                d = it / G::FACE; // directions 0/1: across the slowest dim
                if (d == 0 ? (t0 != 0) : (t0 + BAND != S)) continue;
                slab_index<R, S, H>(d, 0, it % G::FACE, idx);
            }
            else
            {
                const int r = it - 2 * G::FACE;
                d           = 2 + r / SIDE;
                int t       = r % SIDE;
                // tangential enumeration restricted to the band rows
                const int dim = d >> 1;
                const int pos = d & 1;
#pragma unroll
                for (int k = R - 1; k >= 1; --k)
                {
                    if (k == dim)
                        idx[k] = pos ? (H + S) : (H - 1);
                    else
                    {
                        idx[k] = H + (t % S);
                        t /= S;
                    }
                }
                idx[0] = H + t0 + t;
            }
            const int m = sMeta[d];
            if ((m & 3) == 0) continue;
            const int32_t* nb = sNbr + d * G::KF;
            int            lo = (idx[0] - row0) * P0;
#pragma unroll
            for (int k = 1; k < R; ++k) lo += idx[k] * G::pitch(k);
#pragma unroll
            for (int f = 0; f < NV; ++f)
                sU[f * TILE + lo] = halo_source<R, S, H>(a.cur.p[f], nb, m, d, idx);
        }
        __syncthreads();
    }

    if constexpr (EQ == kEqEuler)
    {
        // primitives once per tile cell (EulerPhysics.hpp:83-99): p, a, u_d
        const double g = a.gamma;
        for (int c = tid; c < TILE; c += NT)
        {
            const double rho  = sU[c];
            const double irho = 1.0 / rho;
            double       K    = 0.0;
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
            {
                const double m = sU[(1 + ds) * TILE + c];
                K += m * m;
            }
            K *= 0.5 * irho;
            const double pr = (g - 1.0) * (sU[(R + 1) * TILE + c] - K);
            sW[c]           = pr;
            sW[TILE + c]    = sqrt(g * pr * irho);
#pragma unroll
            for (int ds = 0; ds < R; ++ds) sW[(2 + ds) * TILE + c] = sU[(1 + ds) * TILE + c] * irho;
        }
        __syncthreads();
    }

    double       smax_new[R];
#pragma unroll
    for (int ds = 0; ds < R; ++ds) smax_new[ds] = 0.0;
    const size_t gbase = (size_t)p * G::FLAT;

    for (int ci = tid; ci < BAND * ROWCELL; ci += NT)
    {
        // interior cell -> tile-local / patch-global linear index
        int r = ci, lt = 0, gl = 0;
#pragma unroll
        for (int k = R - 1; k >= 1; --k)
        {
            const int i = H + (r % S);
            r /= S;
            lt += i * G::pitch(k);
            gl += i * G::pitch(k);
        }
        lt += (r + 1) * P0;
        gl += (H + t0 + r) * P0;

        double Uc[NV], upd[NV];
#pragma unroll
        for (int f = 0; f < NV; ++f)
        {
            Uc[f]  = sU[f * TILE + lt];
            upd[f] = 0.0;
        }

        if constexpr (EQ == kEqAdvection)
        {
#pragma unroll
            for (int ds = 0; ds < ((R == 3) ? 2 : R); ++ds) // ds = 2: zero velocity, zero flux
            {
                const int    st = G::pitch(R - 1 - ds);
                const double v  = adv_vel(ds);
                const double uL = sU[lt - st], uR = sU[lt + st];
                // AdvectionPhysics.hpp:45-66
                const double fL = 0.5 * (uL * v + Uc[0] * v) - 0.5 * fabs(v) * (Uc[0] - uL);
                const double fR = 0.5 * (Uc[0] * v + uR * v) - 0.5 * fabs(v) * (uR - Uc[0]);
                upd[0] -= dtdx[ds] * (fR - fL);
            }
        }
        else
        {
            const double pc = sW[lt], ac = sW[TILE + lt];
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
            {
                const int    st  = G::pitch(R - 1 - ds);
                const double uc  = sW[(2 + ds) * TILE + lt];
                const double sc_ = fabs(uc) + ac;
                double       fl[2][NV];
#pragma unroll
                for (int side = 0; side < 2; ++side)
                {
                    const int    n  = side ? lt + st : lt - st;
                    const double pn = sW[n], an = sW[TILE + n], un = sW[(2 + ds) * TILE + n];
                    const double sm = fmax(fabs(un) + an, sc_);
                    double       Un[NV];
#pragma unroll
                    for (int f = 0; f < NV; ++f) Un[f] = sU[f * TILE + n];
                    // left/right roles: side 0 -> (L = n, R = c); side 1 -> (L = c, R = n)
                    const double sg = side ? 1.0 : -1.0; // UR - UL = sg * (Un - Uc)
                    fl[side][0] = 0.5 * (Un[1 + ds] + Uc[1 + ds] - sm * (sg * (Un[0] - Uc[0])));
#pragma unroll
                    for (int k = 0; k < R; ++k)
                    {
                        double fn = Un[1 + k] * un, fc = Uc[1 + k] * uc;
                        if (k == ds)
                        {
                            fn += pn;
                            fc += pc;
                        }
                        fl[side][1 + k] =
                            0.5 * (fn + fc - sm * (sg * (Un[1 + k] - Uc[1 + k])));
                    }
                    const double en = un * (Un[R + 1] + pn), ec = uc * (Uc[R + 1] + pc);
                    fl[side][R + 1] = 0.5 * (en + ec - sm * (sg * (Un[R + 1] - Uc[R + 1])));
                }
#pragma unroll
                for (int f = 0; f < NV; ++f) upd[f] -= dtdx[ds] * (fl[1][f] - fl[0][f]);
            }
        }

        double Un_[NV];
#pragma unroll
        for (int f = 0; f < NV; ++f)
        {
            Un_[f]                  = Uc[f] + upd[f];
            a.nxt.p[f][gbase + gl] = Un_[f];
        }
        // Store the whole padded x-row: the first / last interior cell of a row also fills the
        // ghost columns next to it, so that rows are written as full 32-byte sectors (a partially
        // written sector costs an L2 read-modify-write; tools/store_bench.cu: 2.9 vs 5.6 TB/s).
        // Those cells are x-face ghosts: gathered from the neighbor interiors by the next step and
        // rewritten by halo_kernel before anything observes them.
        {
            const int xi = ci % S;
            if (xi == 0 || xi == S - 1)
            {
                const int dir = (xi == 0) ? -1 : 1;
#pragma unroll
                for (int f = 0; f < NV; ++f)
#pragma unroll
                    for (int h = 1; h <= H; ++h) a.nxt.p[f][gbase + gl + dir * h] = Un_[f];
            }
        }

        if constexpr (EQ == kEqEuler)
        {
            // wave speed of the new state (EulerPhysics.hpp:137-161) for the next step's CFL
            const double irho = 1.0 / Un_[0];
            double       K    = 0.0;
#pragma unroll
            for (int ds = 0; ds < R; ++ds) K += Un_[1 + ds] * Un_[1 + ds];
            K *= 0.5 * irho;
            const double pr = (a.gamma - 1.0) * (Un_[R + 1] - K);
            const double cs = sqrt(a.gamma * pr * irho);
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
                smax_new[ds] = fmax(smax_new[ds], fabs(Un_[1 + ds] * irho) + cs);
        }
    }

    if (a.sc.dtmin_out != nullptr)
    {
        if constexpr (EQ == kEqAdvection)
        {
#pragma unroll
            for (int ds = 0; ds < R; ++ds) smax_new[ds] = fabs(adv_vel(ds));
        }
        else
        {
            block_max<R, NT>(smax_new, red);
        }
        if (tid == 0)
        {
            // min over cells of dx/speed == dx / max speed (division is monotone), guard as in
            // amr_solver.hpp:399 / fvm_time_step.cu:169
            double cand = DBL_MAX;
#pragma unroll
            for (int ds = 0; ds < R; ++ds)
                if (smax_new[ds] > 1e-12) cand = fmin(cand, dxs[ds] / smax_new[ds]);
            atomicMin(a.sc.dtmin_out, (unsigned long long)__double_as_longlong(cand));
        }
    }
}

// ---- standalone CFL reduction over the current buffer (first step of a batch) ----------------------
template <int R, int S, int H, int EQ, int NT, int HS = H>
__global__ void __launch_bounds__(NT)
compute_dt_kernel(FieldPtrs cur, const int32_t* __restrict__ level, int n_patches, double gamma,
                  const __grid_constant__ StepArgs a, unsigned long long* dtmin_out)
{
    using G          = Geo<R, S, H, HS>;
    constexpr int NV = EqTraits<EQ, R>::NV;
    __shared__ double red[R * (NT / 32)];
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    double smax[R];
#pragma unroll
    for (int ds = 0; ds < R; ++ds) smax[ds] = 0.0;
    if constexpr (EQ == kEqEuler)
    {
        for (int ci = threadIdx.x; ci < G::DATA; ci += NT)
        {
            int r = ci, gl = 0;
#pragma unroll
            for (int k = R - 1; k >= 0; --k)
            {
                gl += (G::O + (r % S)) * G::pitch(k);
                r /= S;
            }
            double U[NV];
#pragma unroll
            for (int f = 0; f < NV; ++f) U[f] = cur.p[f][(size_t)p * G::FLAT + gl];
            const double irho = 1.0 / U[0];
            double       K    = 0.0;
#pragma unroll
            for (int ds = 0; ds < R; ++ds) K += U[1 + ds] * U[1 + ds];
            K *= 0.5 * irho;
            const double pr = (gamma - 1.0) * (U[R + 1] - K);
            const double cs = sqrt(gamma * pr * irho);
#pragma unroll
            for (int ds = 0; ds < R; ++ds) smax[ds] = fmax(smax[ds], fabs(U[1 + ds] * irho) + cs);
        }
        block_max<R, NT>(smax, red);
    }
    else
    {
#pragma unroll
        for (int ds = 0; ds < R; ++ds) smax[ds] = fabs(adv_vel(ds));
    }
    if (threadIdx.x == 0)
    {
        const int lvl  = level[p];
        double    cand = DBL_MAX;
#pragma unroll
        for (int ds = 0; ds < R; ++ds)
            if (smax[ds] > 1e-12) cand = fmin(cand, a.dx[lvl][ds] / smax[ds]);
        atomicMin(dtmin_out, (unsigned long long)__double_as_longlong(cand));
    }
}

// ---- batch scalar init: dtmin[1..n] = DBL_MAX, remaining[0] = remaining -----------------------------
static __global__ void init_scalars_kernel(unsigned long long* dtmin, int first, int count,
                                    double* remaining, double remaining_value, double* dts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count)
    {
        dtmin[first + i] = (unsigned long long)__double_as_longlong(DBL_MAX);
        if (dts) dts[i] = 0.0;
    }
    if (i == 0 && remaining) remaining[0] = remaining_value;
}

// ---- refine/coarsen data motion fused with the Morton re-sort ---------------------------------------
// new patch q <- kind 0: whole padded copy of old patch src (permutation.cu semantics)
//               kind 1: prolongation of child `child` from old patch src (ndtree.hpp:1528-1555)
//               kind 2: restriction of old patches src .. src+2^R-1       (ndtree.hpp:1499-1526)
// child number c: row-major over (2,..,2), last layout dim = bit 0 (ndtree.hpp:1463-1497)
template <int R, int S, int H, int NV, int HS = H>
__global__ void __launch_bounds__(256)
plan_kernel(FieldPtrs old_, FieldPtrs new_, const int8_t* __restrict__ kind,
            const int32_t* __restrict__ src, const int8_t* __restrict__ child, int n_new)
{
    using G     = Geo<R, S, H, HS>;
    const int q = blockIdx.x;
    if (q >= n_new) return;
    const int    kd = kind[q];
    const size_t s0 = (size_t)src[q];
    if (kd == 0)
    {
        for (int f = 0; f < NV; ++f)
        {
            const double2* s = reinterpret_cast<const double2*>(old_.p[f] + s0 * G::FLAT);
            double2*       d = reinterpret_cast<double2*>(new_.p[f] + (size_t)q * G::FLAT);
            for (int i = threadIdx.x; i < G::FLAT / 2; i += blockDim.x) d[i] = s[i];
        }
        return;
    }
    // ghosts of freshly created patches are zero until the next halo exchange
    if constexpr (HS > 0)
    {
        for (int f = 0; f < NV; ++f)
            for (int i = threadIdx.x; i < G::FLAT; i += blockDim.x)
                new_.p[f][(size_t)q * G::FLAT + i] = 0.0;
        __syncthreads();
    }
    const int cn = child[q];
    for (int ci = threadIdx.x; ci < G::DATA; ci += blockDim.x)
    {
        int r = ci, gl = 0, i[R];
#pragma unroll
        for (int k = R - 1; k >= 0; --k)
        {
            i[k] = r % S;
            r /= S;
            gl += (G::O + i[k]) * G::pitch(k);
        }
        if (kd == 1)
        {
            // fine cell i of child cn <- coarse cell cbit*(S/2) + i/2
            int co = 0;
#pragma unroll
            for (int k = 0; k < R; ++k)
            {
                const int cb = (cn >> (R - 1 - k)) & 1;
                co += (G::O + cb * (S / 2) + i[k] / 2) * G::pitch(k);
            }
            for (int f = 0; f < NV; ++f)
                new_.p[f][(size_t)q * G::FLAT + gl] = old_.p[f][s0 * G::FLAT + co];
        }
        else
        {
            // coarse cell i <- mean of the 2^R cells of child number from i / (S/2)
            int ch = 0, base = 0;
#pragma unroll
            for (int k = 0; k < R; ++k)
            {
                ch |= (i[k] / (S / 2)) << (R - 1 - k);
                base += (G::O + (i[k] % (S / 2)) * 2) * G::pitch(k);
            }
            for (int f = 0; f < NV; ++f)
            {
                const double* sp  = old_.p[f] + (s0 + ch) * G::FLAT + base;
                double        sum = 0.0;
#pragma unroll
                for (int n = 0; n < G::FAN; ++n)
                {
                    int off = 0;
#pragma unroll
                    for (int k = 0; k < R; ++k) off += ((n >> (R - 1 - k)) & 1) * G::pitch(k);
                    sum += sp[off];
                }
                new_.p[f][(size_t)q * G::FLAT + gl] = sum / (double)G::FAN;
            }
        }
    }
}

// ---- per-patch max over ALL flat cells -> refine flags (fvm_refinement_criterion.cu:27-67) ----------
template <int FLAT>
__global__ void __launch_bounds__(128)
patch_max_flags_kernel(const double* __restrict__ field, const int32_t* __restrict__ level,
                       int n_patches, double refine_thr, double coarsen_thr, int min_level,
                       int max_level, int8_t* __restrict__ flags)
{
    __shared__ double red[4];
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    double m = -DBL_MAX;
    for (int i = threadIdx.x; i < FLAT; i += 128)
    {
        const double v = field[(size_t)p * FLAT + i];
        m              = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double t = __shfl_xor_sync(0xffffffffu, m, o);
        m              = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 4; ++i) m = red[i] > m ? red[i] : m;
        const int lv = level[p];
        int8_t    fl = 0;
        if (lv < max_level && m > refine_thr)
            fl = 1;
        else if (lv > min_level && m < coarsen_thr)
            fl = 2;
        flags[p] = fl;
    }
}

// runtime-sized variant behind the raw-pointer criterion entry point
static __global__ void __launch_bounds__(128)
patch_max_flags_rt_kernel(const double* __restrict__ field, const int32_t* __restrict__ level,
                          int n_patches, int flat, double refine_thr, double coarsen_thr,
                          int min_level, int max_level, int8_t* __restrict__ flags)
{
    __shared__ double red[4];
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    double m = -DBL_MAX;
    for (int i = threadIdx.x; i < flat; i += 128)
    {
        const double v = field[(size_t)p * flat + i];
        m              = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double t = __shfl_xor_sync(0xffffffffu, m, o);
        m              = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 4; ++i) m = red[i] > m ? red[i] : m;
        const int lv = level[p];
        int8_t    fl = 0;
        if (lv < max_level && m > refine_thr)
            fl = 1;
        else if (lv > min_level && m < coarsen_thr)
            fl = 2;
        flags[p] = fl;
    }
}

// ---- interior <-> padded (host staging helpers) -------------------------------------------------------
template <int R, int S, int H>
__global__ void __launch_bounds__(256)
interior_copy_kernel(double* __restrict__ padded, double* __restrict__ dense, int n_patches,
                     int to_padded)
{
    using G     = Geo<R, S, H>;
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    for (int ci = threadIdx.x; ci < G::DATA; ci += blockDim.x)
    {
        int r = ci, gl = 0;
#pragma unroll
        for (int k = R - 1; k >= 0; --k)
        {
            gl += (H + (r % S)) * G::pitch(k);
            r /= S;
        }
        if (to_padded)
            padded[(size_t)p * G::FLAT + gl] = dense[(size_t)p * G::DATA + ci];
        else
            dense[(size_t)p * G::DATA + ci] = padded[(size_t)p * G::FLAT + gl];
    }
}

// ---- inter-GPU ghost faces -----------------------------------------------------------------------------
// entry e = {patch, direction | layers << 4}; slab = the T = min(2H, S) interior layers next to face `direction`,
// every field (a finer neighbor restricts 2 fine layers per coarse ghost layer, patch_utils.hpp:
// 334-386, so 2H layers cover all three halo operators); buffer layout [entry][field][layer][face cell].
// `layers` (0 = all T) is how many of them are moved: H are enough when no COARSER patch reads this face
// (same-level copy and injection into a finer patch look at the first H layers only) -- on a mostly uniform
// mesh that halves the bytes on the wire; the slab keeps its fixed size and layout.
template <int R, int S, int H>
struct SlabGeo
{
    static constexpr int T    = (2 * H < S) ? 2 * H : S;
    static constexpr int SLAB = T * Geo<R, S, H>::FACE;
};

template <int R, int S, int H, int NV, int HS = H>
__global__ void __launch_bounds__(128)
face_pack_kernel(FieldPtrs cur, const int32_t* __restrict__ entries, int count,
                 double* __restrict__ buffer, int unpack)
{
    using G     = Geo<R, S, H, HS>;
    const int e = blockIdx.x;
    if (e >= count) return;
    const int     p = entries[2 * e], dw = entries[2 * e + 1], d = dw & 15;
    const int     dim = d >> 1, pos = d & 1;
    constexpr int SLAB = SlabGeo<R, S, H>::SLAB;
    const int     part = ((dw >> 4) > 0 && (dw >> 4) < SlabGeo<R, S, H>::T) ? (dw >> 4) * G::FACE : SLAB;
    for (int i2 = threadIdx.x; i2 < NV * part; i2 += blockDim.x)
    {
        const int f = i2 / part, r = i2 % part, layer = r / G::FACE, it = f * SLAB + r;
        int       t = r % G::FACE, gl = 0;
#pragma unroll
        for (int k = R - 1; k >= 0; --k)
        {
            int i;
            if (k == dim)
                i = pos ? (G::O + S - 1 - layer) : (G::O + layer);
            else
            {
                i = G::O + (t % S);
                t /= S;
            }
            gl += i * G::pitch(k);
        }
        double* cell = cur.p[f] + (size_t)p * G::FLAT + gl;
        double* slot = buffer + (size_t)e * NV * SLAB + it;
        if (unpack)
            *cell = *slot;
        else
            *slot = *cell;
    }
}

} // namespace amrb
