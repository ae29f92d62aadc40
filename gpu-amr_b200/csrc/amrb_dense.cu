// Interior-only 3D pools (amrb_layout.storage = AMRB_STORAGE_INTERIOR): kernel instantiations and their
// dispatch entries.  Separate translation unit: it compiles in parallel with amrb_api.cu.
#include "amrb_advect3d_dense.cuh"
#include "amrb_dense.cuh"
#include "amrb_march_euler3d_dense.cuh"
#include "amrb_ops.cuh"

#include <algorithm>

namespace amrb
{

template <int S, int EQ>
struct DenseInst
{
    static constexpr int R  = 3;
    static constexpr int H  = 1;
    static constexpr int NV = EqTraits<EQ, R>::NV;
    static constexpr int NT = 256;

    static cudaError_t prepare() { return cudaSuccess; } // marching kernels opt in at first launch, per device

    static void halo_fill(cudaStream_t, const FieldPtrs&, const int32_t*, const uint8_t*, int) {} // no stored ghosts

    template <int CR, int NS, int WPC, int MINB>
    static void march(cudaStream_t st, const StepArgs& a, int n_items)
    {
        using MC = March3DenseCfg<S, CR, NS, WPC>;
        auto k   = euler3d_dense_kernel<S, CR, NS, WPC, MINB>;
        static DevicePrepared prepared;
        if (!prepared.ensure((const void*)k, (int)MC::SMEM)) return;
        const int tasks = n_items * MC::NB;
        const int grid  = std::max(1, std::min(device_sm_count() * MINB, (tasks + WPC - 1) / WPC));
        k<<<grid, WPC * 32, MC::SMEM, st>>>(a, n_items);
    }
    template <int CR, int NS, int WPC, int MINB>
    static void march_pp(cudaStream_t st, const StepArgs& a, int n_items)
    {
        using MC = March3DenseCfg<S, CR, NS, WPC>;
        auto k   = euler3d_dense_kernel_pp<S, CR, NS, WPC, MINB>;
        static DevicePrepared prepared;
        if (!prepared.ensure((const void*)k, (int)MC::SMEM)) return;
        const int tasks = n_items * MC::NB;
        const int grid  = std::max(1, std::min(device_sm_count() * MINB, (tasks + WPC - 1) / WPC));
        k<<<grid, WPC * 32, MC::SMEM, st>>>(a, n_items);
    }
    template <int CR, int NS, int WPC, int MINB, int OPT>
    static void march_o(cudaStream_t st, const StepArgs& a, int n_items)
    {
        using MC = March3DenseCfg<S, CR, NS, WPC>;
        auto k   = euler3d_dense_kernel_o<S, CR, NS, WPC, MINB, OPT>;
        static DevicePrepared prepared;
        if (!prepared.ensure((const void*)k, (int)MC::SMEM)) return;
        const int tasks = n_items * MC::NB;
        const int grid  = std::max(1, std::min(device_sm_count() * MINB, (tasks + WPC - 1) / WPC));
        k<<<grid, WPC * 32, MC::SMEM, st>>>(a, n_items);
    }
    template <int WPC, int MINB>
    static void advect(cudaStream_t st, const StepArgs& a, int n_items)
    {
        using AC = Adv3DenseCfg<S, WPC>;
        auto k   = advect3d_dense_kernel<S, WPC, MINB>;
        static DevicePrepared prepared;
        if (!prepared.ensure((const void*)k, (int)AC::SMEM)) return;
        const int tasks = n_items * AC::NB;
        const int grid  = std::max(1, std::min(device_sm_count() * MINB, (tasks + WPC - 1) / WPC));
        k<<<grid, WPC * 32, AC::SMEM, st>>>(a, n_items);
    }
    static void step(cudaStream_t st, const StepArgs& a, int n_items)
    {
        if constexpr (EQ == kEqEuler)
        {
            // variant (amrb_pool_set_variant): ring shape A/B.  0 = chunks of 4 planes (2 KB bulk copies),
            // 2 stages (8^3) / 1-plane cp.async chunks, 4 stages (16^3); 21 = 2-plane chunks, 3 stages;
            // 28 = two planes per loop trip; 41 / 44 = body options (amrb_march_euler3d_dense.cuh: kOpt*); all within
            // 1 % of variant 0 (profiles/r02_summary.md).  Earlier A/B runs, removed again: 3 CTAs per SM with 2-plane
            // chunks, 9 / 10 / 12 warps per SM under __maxnreg__ (slower: spills), one CTA per SM (5.29 vs 3.53 ms), parked
            // fluxes loaded by the boundary lanes only (equal), lateral ghost gathers of ONE field instead of five (a wrong-
            // result probe of what the per-lane cp.async gathers cost: 3.47 -> 3.08 ms)
            if constexpr (S == 8)
            {
                if (a.variant == 21)
                    march<2, 3, 4, 2>(st, a, n_items);
                else if (a.variant == 28)
                    march_pp<4, 2, 4, 2>(st, a, n_items); // two planes per loop trip, no state copies
                else if (a.variant == 41)
                    march_o<4, 2, 4, 2, kOptPrefetch>(st, a, n_items); // task prologue gathered one task ahead
                else if (a.variant == 44)
                    march_o<4, 2, 4, 2, kOptPrefetch | kOptOneBlock | kOptEarlyZ>(st, a, n_items);
                else
                    march<4, 2, 4, 2>(st, a, n_items);
            }
            else if (a.variant == 21)
                march<2, 4, 4, 2>(st, a, n_items); // 2-plane chunks: 4.49 ms per 1.34e8 cells
            else
                march<1, 4, 4, 2>(st, a, n_items); // 16^3: 1-plane chunks of the block's 8 x 8 tile, 4 stages: 3.69 ms
        }
        else
        {
            // variant 31: 4 warps per CTA, 4 CTAs per SM instead of 8 x 2
            if (a.variant == 31)
                advect<4, 4>(st, a, n_items);
            else
                advect<8, 2>(st, a, n_items);
        }
    }
    static void compute_dt(cudaStream_t st, const StepArgs& a, unsigned long long* out)
    {
        compute_dt_kernel<R, S, H, EQ, NT, 0><<<a.n_patches, NT, 0, st>>>(a.cur, a.level, a.n_patches, a.gamma, a, out);
    }
    static void plan(cudaStream_t st, const FieldPtrs& o, const FieldPtrs& n, const int8_t* kind,
                     const int32_t* src, const int8_t* child, int count)
    {
        plan_kernel<R, S, H, NV, 0><<<count, 256, 0, st>>>(o, n, kind, src, child, count);
    }
    static void interior(cudaStream_t st, double* padded, double* dense, int n, int to_padded)
    {
        interior_copy_kernel<R, S, H><<<n, 256, 0, st>>>(padded, dense, n, to_padded);
    }
    static void faces(cudaStream_t st, const FieldPtrs& cur, const int32_t* entries, int count, double* buffer,
                      int unpack)
    {
        face_pack_kernel<R, S, H, NV, 0><<<count, 128, 0, st>>>(cur, entries, count, buffer, unpack);
    }
    static void export_padded(cudaStream_t st, const double* field, const int32_t* nbr, const uint8_t* meta,
                              int first, int n, int n_tabled, double* staging)
    {
        dense_export_kernel<R, S, H><<<n, 256, 0, st>>>(field, nbr, meta, first, n, n_tabled, staging);
    }
    static void flags_dense(cudaStream_t st, const double* field, const int32_t* nbr, const uint8_t* meta,
                            const int32_t* level, int n, double rt, double ct, int minl, int maxl, int8_t* out)
    {
        dense_flags_kernel<R, S, H><<<n, 128, 0, st>>>(field, nbr, meta, level, n, rt, ct, minl, maxl, out);
    }
    static constexpr Ops ops()
    {
        return Ops{ R,     S,        H,           EQ,    1,       1,         0,      &prepare, &halo_fill,
                    &step, &step,    &compute_dt, &plan, nullptr, &interior, &faces, &export_padded,
                    &flags_dense };
    }
};

static const Ops g_dense_ops[] = {
    DenseInst<8, kEqAdvection>::ops(),
    DenseInst<8, kEqEuler>::ops(),
    DenseInst<16, kEqAdvection>::ops(),
    DenseInst<16, kEqEuler>::ops(),
};

const Ops* dense_ops(int* count)
{
    *count = (int)(sizeof(g_dense_ops) / sizeof(g_dense_ops[0]));
    return g_dense_ops;
}

} // namespace amrb
