// C-ABI of the device side: memory/sync wrappers, the device patch pool and all kernel launches.
// See include/gpuamr_b200.h for the contract and the reference interfaces each group replaces.
#include "amrb_kernels.cuh"
#include "amrb_ops.cuh"
#include "amrb_pool.h"
#include "amrb_step_euler.cuh"
#include "amrb_march_euler.cuh"
#include "amrb_march_euler3d.cuh"
#include "amrb_advect2d.cuh"

#include "../../include/gpuamr_b200.h"

#include <algorithm>
#include <cmath>
#include <cuda_profiler_api.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace amrb
{

thread_local std::string g_error;

amrb_status fail(amrb_status code, const std::string& what)
{
    g_error = what;
    return code;
}

// Function attributes (opt-in shared-memory size) and the SM count belong to a DEVICE: a process
// may open pools on several GPUs, so both are cached per device ordinal (the caller has made the
// pool's device current).  A failed attribute call is kept for check_launch to report.
constexpr int kMaxDevices = 64;
thread_local cudaError_t g_prepare_error = cudaSuccess;

static int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
int device_sm_count()
{
    static int sms[kMaxDevices] = {};
    const int  dev = current_device();
    if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev];
}
bool DevicePrepared::ensure(const void* kernel, int smem_bytes)
{
    const int dev = current_device();
    if (done[dev]) return true;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess)
    {
        g_prepare_error = e;
        return false;
    }
    done[dev] = true;
    return true;
}

// ---------------------------------------------------------------------------------- kernel dispatch
// struct Ops: amrb_ops.cuh

template <int R, int S, int H, int EQ, int BAND, int EB, int RG, int TPC, int ENT>
struct Inst
{
    using EC = EulerStepCfg<R, S, H, EB, RG, TPC, ENT>;
    using G                 = Geo<R, S, H>;
    static constexpr int NV = EqTraits<EQ, R>::NV;
    static constexpr int NW = EqTraits<EQ, R>::NW;
    static constexpr int NT = 256;
    static constexpr size_t SMEM =
        (size_t)(NV + NW) * (BAND + 2) * G::pitch(0) * sizeof(double);

    static cudaError_t prepare()
    {
        if constexpr (EQ == kEqEuler)
        {
            cudaError_t e = cudaFuncSetAttribute(euler_step_kernel<R, S, H, EB, RG, TPC, ENT>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)EC::SMEM);
            if (e != cudaSuccess) return e;
        }
        return cudaFuncSetAttribute(step_kernel<R, S, H, EQ, BAND, NT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    }
    static void halo_fill(cudaStream_t st, const FieldPtrs& cur, const int32_t* nbr,
                          const uint8_t* meta, int n)
    {
        halo_kernel<R, S, H, NV><<<n, 256, 0, st>>>(cur, nbr, meta, n);
    }
    static void step_v1(cudaStream_t st, const StepArgs& a, int n_items)
    {
        step_kernel<R, S, H, EQ, BAND, NT><<<n_items * (S / BAND), NT, SMEM, st>>>(a);
    }
    // warp-autonomous marching kernel (amrb_march_euler.cuh): persistent grid, MINB CTAs per SM
    static int sm_count() { return device_sm_count(); }
    template <int BANDM, int CR, int NS, int WPC, int MINB>
    static void march(cudaStream_t st, const StepArgs& a, int n_items)
    {
        using MC = March2Cfg<S, H, BANDM, CR, NS, WPC>;
        auto k   = euler2d_march_kernel<S, H, BANDM, CR, NS, WPC, MINB>;
        static DevicePrepared prepared;
        if (!prepared.ensure((const void*)k, (int)MC::SMEM)) return;
        const int tasks = n_items * MC::NB;
        const int grid  = std::max(1, std::min(sm_count() * MINB, (tasks + WPC - 1) / WPC));
        k<<<grid, WPC * 32, MC::SMEM, st>>>(a, n_items);
    }
    // Task height of the marching kernel: taller bands stream fewer halo rows (2 per band) and
    // store longer contiguous runs; shorter bands balance small meshes over the 8 warps x SMs.
    // score = wave efficiency / (1 + 2 / band)
    static int pick_band(int n_items)
    {
        const double warps = 8.0 * sm_count();
        int          best = 16;
        double       best_score = 0.0;
        for (int band : { 64, 32, 16 })
        {
            const double rounds = n_items * (double)(S / band) / warps;
            const double eff    = rounds / std::ceil(rounds);
            const double score  = eff / (1.0 + 2.0 / band);
            if (score > best_score * 1.0001)
            {
                best_score = score;
                best       = band;
            }
        }
        return best;
    }
    // 3D warp-autonomous plane-marching kernel (amrb_march_euler3d.cuh)
    template <int CR, int NS, int WPC, int MINB, int CPRING = 0>
    static void march3(cudaStream_t st, const StepArgs& a, int n_items)
    {
        using MC = March3Cfg<S, H, CR, NS, WPC>;
        auto k   = euler3d_march_kernel<S, H, CR, NS, WPC, MINB, CPRING>;
        static DevicePrepared prepared;
        if (!prepared.ensure((const void*)k, (int)MC::SMEM)) return;
        const int tasks = n_items * MC::NB;
        const int grid  = std::max(1, std::min(sm_count() * MINB, (tasks + WPC - 1) / WPC));
        k<<<grid, WPC * 32, MC::SMEM, st>>>(a, n_items);
    }
    static void step(cudaStream_t st, const StepArgs& a, int n_items)
    {
        if constexpr (EQ == kEqEuler && R == 3 && H == 1 && S % 8 == 0)
        {
            // AMRB_VARIANT: 0 = plane-marching kernel (default; ring of 2-plane TMA copies for 8^3 patches,
            // of per-lane cp.async row blocks for 16^3); 11 = 1-plane chunks, 4 stages; 14/15 = cp.async ring for
            // 8^3 as well (2-plane / 1-plane chunks); 10 = block-cooperative
            // pipeline (second generation)
            const int v = a.variant;
            if (v != 10)
            {
                if (v == 11)
                    march3<1, 4, 4, 2>(st, a, n_items);
                else if (v == 14)
                    march3<(S == 8 ? 2 : 1), (S == 8 ? 2 : 4), 4, 2, 1>(st, a, n_items);
                else if (v == 15)
                    march3<1, 4, 4, 2, 1>(st, a, n_items);
                else
                    march3<(S == 8 ? 2 : 1), (S == 8 ? 2 : 4), 4, 2>(st, a, n_items);
                return;
            }
        }
        if constexpr (EQ == kEqEuler && R == 2 && S == 64 && H == 1)
        {
            // AMRB_VARIANT: 0 = marching kernel, band chosen per launch (default); 1/2/3 = marching
            // kernel with 16/32/64-row bands; 10 = block-cooperative pipeline (second generation)
            const int v    = a.variant;
            const int band = (v == 1) ? 16 : (v == 2) ? 32 : (v == 3) ? 64 : pick_band(n_items);
            if (v != 10)
            {
                // AMRB_RING: ring shape (rows per TMA copy / stages): a warp's bulk copies complete
                // one at a time (~0.5 us each, tools/tma_bench.cu), so taller chunks = fewer waits
                static const int ring = getenv("AMRB_RING") ? atoi(getenv("AMRB_RING")) : 0;
                if (band == 64 && ring == 1)
                    march<64, 3, 3, 4, 2>(st, a, n_items);
                else if (band == 64 && ring == 2)
                    march<64, 6, 2, 4, 2>(st, a, n_items);
                else if (band == 64 && ring == 3)
                    march<64, 3, 2, 4, 2>(st, a, n_items);
                else if (band == 64)
                    march<64, 2, 3, 4, 2>(st, a, n_items);
                else if (band == 32)
                    march<32, 2, 3, 4, 2>(st, a, n_items);
                else
                    march<16, 2, 3, 4, 2>(st, a, n_items);
                return;
            }
        }
        if constexpr (EQ == kEqEuler)
        {
            const int tiles = n_items * EC::NBANDS;
            euler_step_kernel<R, S, H, EB, RG, TPC, ENT>
                <<<(tiles + TPC - 1) / TPC, ENT, EC::SMEM, st>>>(a, n_items);
        }
        else
        {
            if constexpr (R == 2)
            {
                // warp-autonomous streaming kernel (amrb_advect2d.cuh), upwind form: narrow patches stream in groups
                // of whole patches, patches of 32 cells and wider in 16-row bands.  variant 10 = the thread-per-cell
                // kernel, 41 = 8-row bands.  Measured (profiles/r02z_dev_bench.jsonl, fraction of the HBM roofline):
                // 64-wide 0.75 (16-row bands) / 0.68 (8-row) / 0.60 (thread per cell); 32-wide 0.58; 16-wide 0.44;
                // 10 x 10 / halo 2 (the C1 shape) 0.27; before the upwind form: 0.57 / - / 0.59, 0.43, 0.39, 0.22
                if (a.variant != 10)
                {
                    auto launch = [&](auto kern, size_t smem, int tasks, int wpc, int ctas) {
                        static DevicePrepared prepared;
                        if (!prepared.ensure((const void*)kern, (int)smem)) return;
                        const int grid = std::max(1, std::min(sm_count() * ctas, (tasks + wpc - 1) / wpc));
                        kern<<<grid, wpc * 32, smem, st>>>(a, n_items);
                    };
                    if (S >= 32 && a.variant != 41)
                    {
                        using AC = Adv2Cfg<S, H, 4, (S % 16 == 0 ? 16 : 8)>;
                        launch(advect2d_kernel<S, H, 4, 2, (S % 16 == 0 ? 16 : 8)>, AC::SMEM, n_items * AC::NB, 4, 2);
                    }
                    else
                    {
                        using AC = Adv2Cfg<S, H, 4>;
                        launch(advect2d_kernel<S, H, 4, 3>, AC::SMEM,
                               AC::WHOLE ? (n_items + AC::TP - 1) / AC::TP : n_items * AC::NB, 4, 3);
                    }
                    return;
                }
            }
            step_v1(st, a, n_items);
        }
    }
    static void compute_dt(cudaStream_t st, const StepArgs& a, unsigned long long* out)
    {
        compute_dt_kernel<R, S, H, EQ, NT>
            <<<a.n_patches, NT, 0, st>>>(a.cur, a.level, a.n_patches, a.gamma, a, out);
    }
    static void plan(cudaStream_t st, const FieldPtrs& o, const FieldPtrs& n, const int8_t* kind,
                     const int32_t* src, const int8_t* child, int count)
    {
        plan_kernel<R, S, H, NV><<<count, 256, 0, st>>>(o, n, kind, src, child, count);
    }
    static void flags(cudaStream_t st, const double* field, const int32_t* level, int n,
                      double rt, double ct, int minl, int maxl, int8_t* out)
    {
        patch_max_flags_kernel<G::FLAT><<<n, 128, 0, st>>>(field, level, n, rt, ct, minl, maxl, out);
    }
    static void interior(cudaStream_t st, double* padded, double* dense, int n, int to_padded)
    {
        interior_copy_kernel<R, S, H><<<n, 256, 0, st>>>(padded, dense, n, to_padded);
    }
    static void faces(cudaStream_t st, const FieldPtrs& cur, const int32_t* entries, int count,
                      double* buffer, int unpack)
    {
        face_pack_kernel<R, S, H, NV><<<count, 128, 0, st>>>(cur, entries, count, buffer, unpack);
    }
    static constexpr Ops ops()
    {
        return Ops{ R,     S,        H,           EQ,    0,      S / BAND,  SMEM,   &prepare, &halo_fill,
                    &step, &step_v1, &compute_dt, &plan, &flags, &interior, &faces, nullptr,  nullptr };
    }
};

// (rank, size, halo, band) for both equations.  Patch shapes of the reference's drivers:
// 10x10/h2 (examples/fvm_solver_advection.e.cpp:24-53), 64x64/h1 (benchmark/
// bench_fvm_solver_integration.b.cpp:34-63), 8^3/h1 (bench_fvm_solver_integration3D.b.cpp:31-64),
// 16x16/h1 (KA-2D), plus the small shapes used by the parity fixtures.
// X(rank, size, halo, band of the thread-per-cell kernel,
//   Euler pipeline: band, rows per marching group, tiles per CTA, threads)
#define AMRB_SHAPES(X)                                                                           \
    X(2, 8, 1, 8, 8, 4, 8, 32)                                                                   \
    X(2, 10, 2, 10, 10, 5, 8, 32)                                                                \
    X(2, 16, 1, 16, 16, 4, 8, 64)                                                                \
    X(2, 32, 1, 16, 32, 4, 4, 128)                                                               \
    X(2, 64, 1, 16, 16, 4, 4, 128)                                                               \
    X(3, 4, 1, 4, 4, 2, 8, 32)                                                                   \
    X(3, 4, 2, 4, 4, 2, 8, 32)                                                                   \
    X(3, 8, 1, 8, 8, 2, 4, 128)                                                                  \
    X(3, 16, 1, 4, 4, 4, 4, 128)

static const Ops g_ops[] = {
#define X(R, S, H, B, EB, RG, TPC, ENT)                                                          \
    Inst<R, S, H, kEqAdvection, B, EB, RG, TPC, ENT>::ops(),                                     \
        Inst<R, S, H, kEqEuler, B, EB, RG, TPC, ENT>::ops(),
    AMRB_SHAPES(X)
#undef X
};

static const Ops* find_ops(const amrb_layout& l)
{
    if (l.rank != 2 && l.rank != 3) return nullptr;
    for (int k = 1; k < l.rank; ++k)
        if (l.size[k] != l.size[0]) return nullptr; // cubic patches only
    const int nv = (l.equation == AMRB_EQ_ADVECTION) ? 1 : l.rank + 2;
    if (l.nvar != nv) return nullptr;
    if (l.storage == AMRB_STORAGE_INTERIOR)
    {
        int        n = 0;
        const Ops* d = dense_ops(&n);
        for (int i = 0; i < n; ++i)
            if (d[i].rank == l.rank && d[i].size == l.size[0] && d[i].halo == l.halo && d[i].eq == l.equation)
                return &d[i];
        return nullptr;
    }
    if (l.storage != AMRB_STORAGE_PADDED) return nullptr;
    for (const Ops& o : g_ops)
        if (o.rank == l.rank && o.size == l.size[0] && o.halo == l.halo && o.eq == l.equation)
            return &o;
    return nullptr;
}

} // namespace amrb

using namespace amrb;

// ---------------------------------------------------------------------------------- pool
// struct amrb_pool: amrb_pool.h

namespace
{

amrb_status set_device(const amrb_pool* p)
{
    AMRB_CUDA(cudaSetDevice(p->device));
    return AMRB_OK;
}

void compute_dx(amrb_pool* p)
{
    // solver/physics_system.hpp:58-85: dx_i = L_i * 2^(Depth-level) / 2^Depth / cells_i with
    // cells_i = data_sizes[rank-1-i]
    const int R = p->lay.rank;
    for (int lvl = 0; lvl <= kMaxLevel; ++lvl)
        for (int i = 0; i < 3; ++i)
        {
            if (i >= R || lvl > p->lay.depth)
            {
                p->dx[lvl][i] = 1.0;
                continue;
            }
            const double pm    = (double)(1u << (p->lay.depth - lvl));
            const double patch = p->lengths[i] * pm / (double)(1u << p->lay.depth);
            p->dx[lvl][i]      = patch / (double)p->lay.size[R - 1 - i];
        }
}

amrb_status ensure_scalars(amrb_pool* p, size_t steps)
{
    if (p->scal_cap >= steps + 2) return AMRB_OK;
    // the carried dt-min of the previous batch lives in the old array: drop it
    p->carry_valid = false;
    if (p->d_dtmin) cudaFree(p->d_dtmin);
    if (p->d_remaining) cudaFree(p->d_remaining);
    if (p->d_dts) cudaFree(p->d_dts);
    if (p->h_dts) cudaFreeHost(p->h_dts);
    const size_t cap = std::max<size_t>(steps + 2, 64);
    AMRB_CUDA(cudaMalloc(&p->d_dtmin, cap * sizeof(unsigned long long)));
    AMRB_CUDA(cudaMalloc(&p->d_remaining, cap * sizeof(double)));
    AMRB_CUDA(cudaMalloc(&p->d_dts, cap * sizeof(double)));
    AMRB_CUDA(cudaHostAlloc(&p->h_dts, cap * sizeof(double), cudaHostAllocDefault));
    p->scal_cap = cap;
    return AMRB_OK;
}

amrb_status ensure_stage(amrb_pool* p, size_t doubles)
{
    if (p->stage_cap >= doubles) return AMRB_OK;
    if (p->d_stage) cudaFree(p->d_stage);
    p->stage_cap = 0;
    AMRB_CUDA(cudaMalloc(&p->d_stage, doubles * sizeof(double)));
    p->stage_cap = doubles;
    return AMRB_OK;
}

void fill_step_args(const amrb_pool* p, StepArgs& a)
{
    a.cur       = p->cur;
    a.nxt       = p->nxt;
    a.nbr       = p->d_nbr;
    a.meta      = p->d_meta;
    a.level     = p->d_level;
    a.list      = nullptr;
    a.n_patches = (int)p->n_owned;
    a.lazy_halo = (p->mode != 1) ? 1 : 0;
    {
        static const int tm = getenv("AMRB_TASKMAP") ? atoi(getenv("AMRB_TASKMAP")) : 1;
        a.task_map = tm;
    }
    a.queue     = nullptr;
    a.variant   = p->variant;
    a.gamma     = p->gamma;
    std::memcpy(a.dx, p->dx, sizeof(a.dx));
    a.sc = StepScalars{ nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, p->cfl };
}

amrb_status check_launch(amrb_pool* p, const char* what)
{
    if (g_prepare_error != cudaSuccess)
    {
        const cudaError_t pe = g_prepare_error;
        g_prepare_error      = cudaSuccess;
        cudaGetLastError();
        return fail(AMRB_ERR_CUDA, std::string(what) + ": cudaFuncSetAttribute(max dynamic shared memory): " +
                                       cudaGetErrorString(pe));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    ++p->launches;
    return AMRB_OK;
}

// wait for the batch enqueued last (amr_solver::wait_for_pending_dt_copy, amr_solver.hpp:416-466)
amrb_status wait_batch(amrb_pool* p)
{
    if (p->pending_in_graph)
        AMRB_CUDA(cudaStreamSynchronize(p->stream));
    else
        AMRB_CUDA(cudaEventSynchronize(p->batch_done));
    p->batch_pending    = false;
    p->pending_in_graph = false;
    return AMRB_OK;
}

amrb_status need_topology(const amrb_pool* p)
{
    if (!p->d_nbr || p->n_owned == 0) return fail(AMRB_ERR_STATE, "pool has no topology yet");
    return AMRB_OK;
}

void swap_buffers(amrb_pool* p) { std::swap(p->cur, p->nxt); } // ndtree.hpp:1558-1579

amrb_status pool_alloc_common(const amrb_layout* layout, size_t capacity, int device,
                              amrb_pool** out)
{
    if (!layout || !out || capacity == 0) return fail(AMRB_ERR_ARGUMENT, "null layout/out or zero capacity");
    const Ops* ops = find_ops(*layout);
    if (!ops) return fail(AMRB_ERR_UNSUPPORTED, "no kernels instantiated for this patch shape / equation");
    if (layout->depth < 1 || layout->depth > kMaxLevel) return fail(AMRB_ERR_ARGUMENT, "depth out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return fail(AMRB_ERR_CUDA, "no CUDA device: the hot path has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(AMRB_ERR_ARGUMENT, "bad device ordinal");
    AMRB_CUDA(cudaSetDevice(device));
    AMRB_CUDA(ops->prepare());
    amrb_pool* p = new amrb_pool();
    p->lay       = *layout;
    p->ops       = ops;
    p->device    = device;
    p->capacity  = capacity;
    p->pflat     = amrb_layout_flat_size(layout);
    p->data      = amrb_layout_data_size(layout);
    p->dense     = layout->storage == AMRB_STORAGE_INTERIOR;
    p->flat      = p->dense ? p->data : p->pflat;
    p->variant   = getenv("AMRB_VARIANT") ? atoi(getenv("AMRB_VARIANT")) : 0;
    compute_dx(p);
    *out = p;
    return AMRB_OK;
}

} // namespace

extern "C" {

const char* amrb_last_error(void) { return g_error.c_str(); }
const char* amrb_version(void) { return "gpuamr_b200 0.1 (sm_100a)"; }
int         amrb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------ memory & sync
amrb_status amrb_device_malloc(void** out, size_t bytes)
{
    if (!out) return fail(AMRB_ERR_ARGUMENT, "null out");
    AMRB_CUDA(cudaMalloc(out, bytes));
    return AMRB_OK;
}
amrb_status amrb_device_free(void* ptr)
{
    AMRB_CUDA(cudaFree(ptr));
    return AMRB_OK;
}
amrb_status amrb_host_pinned_malloc(void** out, size_t bytes)
{
    if (!out) return fail(AMRB_ERR_ARGUMENT, "null out");
    AMRB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return AMRB_OK;
}
amrb_status amrb_host_pinned_free(void* ptr)
{
    AMRB_CUDA(cudaFreeHost(ptr));
    return AMRB_OK;
}
amrb_status amrb_copy_host_to_device(void* dst, const void* src, size_t bytes)
{
    AMRB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return AMRB_OK;
}
amrb_status amrb_copy_host_to_device_async(void* dst, const void* src, size_t bytes, void* stream)
{
    AMRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return AMRB_OK;
}
amrb_status amrb_copy_device_to_host(void* dst, const void* src, size_t bytes)
{
    AMRB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return AMRB_OK;
}
amrb_status amrb_copy_device_to_host_async(void* dst, const void* src, size_t bytes, void* stream)
{
    AMRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return AMRB_OK;
}
amrb_status amrb_copy_device_to_device(void* dst, const void* src, size_t bytes)
{
    AMRB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    return AMRB_OK;
}
amrb_status amrb_stream_create(void** out)
{
    if (!out) return fail(AMRB_ERR_ARGUMENT, "null out");
    cudaStream_t s;
    AMRB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = s;
    return AMRB_OK;
}
amrb_status amrb_stream_destroy(void* stream)
{
    AMRB_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return AMRB_OK;
}
amrb_status amrb_stream_synchronize(void* stream)
{
    AMRB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return AMRB_OK;
}
amrb_status amrb_stream_wait_fence(void* stream, void* fence)
{
    AMRB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)fence, 0));
    return AMRB_OK;
}
amrb_status amrb_fence_create(void** out)
{
    if (!out) return fail(AMRB_ERR_ARGUMENT, "null out");
    cudaEvent_t e;
    AMRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = e;
    return AMRB_OK;
}
amrb_status amrb_fence_destroy(void* fence)
{
    AMRB_CUDA(cudaEventDestroy((cudaEvent_t)fence));
    return AMRB_OK;
}
amrb_status amrb_fence_record(void* fence, void* stream)
{
    AMRB_CUDA(cudaEventRecord((cudaEvent_t)fence, (cudaStream_t)stream));
    return AMRB_OK;
}
amrb_status amrb_fence_wait(void* fence)
{
    AMRB_CUDA(cudaEventSynchronize((cudaEvent_t)fence));
    return AMRB_OK;
}
amrb_status amrb_device_synchronize(void)
{
    AMRB_CUDA(cudaDeviceSynchronize());
    return AMRB_OK;
}

// ------------------------------------------------------------------------------ layout
int amrb_layout_supported(const amrb_layout* layout) { return layout && find_ops(*layout) ? 1 : 0; }
size_t amrb_layout_flat_size(const amrb_layout* l)
{
    size_t n = 1;
    for (int k = 0; k < l->rank; ++k) n *= (size_t)(l->size[k] + 2 * l->halo);
    return n;
}
size_t amrb_layout_storage_size(const amrb_layout* l)
{
    return l->storage == AMRB_STORAGE_INTERIOR ? amrb_layout_data_size(l) : amrb_layout_flat_size(l);
}
size_t amrb_layout_data_size(const amrb_layout* l)
{
    size_t n = 1;
    for (int k = 0; k < l->rank; ++k) n *= (size_t)l->size[k];
    return n;
}

// ------------------------------------------------------------------------------ pool lifetime
amrb_status amrb_pool_create(const amrb_layout* layout, size_t capacity, int device,
                             amrb_pool** out)
{
    AMRB_TRY(pool_alloc_common(layout, capacity, device, out));
    amrb_pool* p = *out;
    p->own_mem   = true;
    const size_t bytes = capacity * p->flat * sizeof(double);
    for (int f = 0; f < layout->nvar; ++f)
    {
        // zero-initialised: corner ghosts are observable
        if (cudaMalloc(&p->cur.p[f], bytes) != cudaSuccess ||
            cudaMalloc(&p->nxt.p[f], bytes) != cudaSuccess)
        {
            cudaGetLastError();
            amrb_pool_destroy(p);
            *out = nullptr;
            return fail(AMRB_ERR_CUDA, "cudaMalloc of the patch pool failed");
        }
        if (cudaMemset(p->cur.p[f], 0, bytes) != cudaSuccess ||
            cudaMemset(p->nxt.p[f], 0, bytes) != cudaSuccess)
        {
            cudaGetLastError();
            amrb_pool_destroy(p);
            *out = nullptr;
            return fail(AMRB_ERR_CUDA, "cudaMemset of the patch pool failed");
        }
    }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        amrb_pool_destroy(p);
        *out = nullptr;
        return fail(AMRB_ERR_CUDA, "cudaStreamCreate failed");
    }
    p->own_stream = true;
    AMRB_CUDA(cudaEventCreateWithFlags(&p->batch_done, cudaEventDisableTiming));
    AMRB_CUDA(cudaDeviceSynchronize());
    return AMRB_OK;
}

amrb_status amrb_pool_create_external(const amrb_layout* layout, size_t capacity, int device,
                                      double* const* cur, double* const* nxt, void* stream,
                                      amrb_pool** out)
{
    if (!cur || !nxt) return fail(AMRB_ERR_ARGUMENT, "null field pointer arrays");
    for (int f = 0; layout && f < layout->nvar && f < kMaxVar; ++f)
        if (!cur[f] || !nxt[f] || ((uintptr_t)cur[f] & 15) || ((uintptr_t)nxt[f] & 15))
            return fail(AMRB_ERR_ARGUMENT, "field pointers must be non-null and 16-byte aligned");
    AMRB_TRY(pool_alloc_common(layout, capacity, device, out));
    amrb_pool* p = *out;
    for (int f = 0; f < layout->nvar; ++f)
    {
        p->cur.p[f] = cur[f];
        p->nxt.p[f] = nxt[f];
    }
    p->stream = (cudaStream_t)stream;
    AMRB_CUDA(cudaEventCreateWithFlags(&p->batch_done, cudaEventDisableTiming));
    return AMRB_OK;
}

amrb_status amrb_pool_destroy(amrb_pool* p)
{
    if (!p) return AMRB_OK;
    cudaSetDevice(p->device);
    if (p->stream || !p->own_stream) cudaStreamSynchronize(p->stream);
    regrid_release(p);
    if (p->own_mem)
        for (int f = 0; f < kMaxVar; ++f)
        {
            if (p->cur.p[f]) cudaFree(p->cur.p[f]);
            if (p->nxt.p[f]) cudaFree(p->nxt.p[f]);
        }
    cudaFree(p->d_nbr);
    cudaFree(p->d_meta);
    cudaFree(p->d_level);
    cudaFree(p->d_dtmin);
    if (p->d_queue) cudaFree(p->d_queue);
    if (p->d_ids) cudaFree(p->d_ids);
    cudaFree(p->d_remaining);
    cudaFree(p->d_dts);
    cudaFree(p->d_stage);
    cudaFree(p->d_flags);
    cudaFree(p->d_plan);
    if (p->h_dts) cudaFreeHost(p->h_dts);
    if (p->batch_done) cudaEventDestroy(p->batch_done);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    cudaGetLastError();
    delete p;
    return AMRB_OK;
}

size_t  amrb_pool_capacity(const amrb_pool* p) { return p ? p->capacity : 0; }
size_t  amrb_pool_size(const amrb_pool* p) { return p ? p->n_owned : 0; }
void*   amrb_pool_stream(const amrb_pool* p) { return p ? (void*)p->stream : nullptr; }
double* amrb_pool_field(const amrb_pool* p, int f)
{
    return (p && f >= 0 && f < p->lay.nvar) ? p->cur.p[f] : nullptr;
}
double* amrb_pool_next_field(const amrb_pool* p, int f)
{
    return (p && f >= 0 && f < p->lay.nvar) ? p->nxt.p[f] : nullptr;
}
uint64_t amrb_pool_launch_count(const amrb_pool* p) { return p ? p->launches : 0; }

amrb_status amrb_pool_set_mode(amrb_pool* p, int mode)
{
    if (!p || mode < 0 || mode > 2) return fail(AMRB_ERR_ARGUMENT, "mode must be 0, 1 or 2");
    if (p->dense && mode != 0)
        return fail(AMRB_ERR_UNSUPPORTED, "interior-only pools run the fused step only (modes 1 / 2 need stored ghosts)");
    p->mode = mode;
    return AMRB_OK;
}

amrb_status amrb_pool_set_variant(amrb_pool* p, int variant)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    p->variant = variant;
    return AMRB_OK;
}

// The caller wrote the current buffer behind the pool's back (through amrb_pool_field /
// amrb_pool_next_field / get_device_buffer pointers): the dt-min carried over from the last batch
// describes a state that no longer exists.
amrb_status amrb_pool_mark_dirty(amrb_pool* p)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is open");
    p->carry_valid = false;
    return AMRB_OK;
}

amrb_status amrb_pool_set_physics(amrb_pool* p, const double* lengths, double gamma, double cfl)
{
    if (!p || !lengths) return fail(AMRB_ERR_ARGUMENT, "null pool/lengths");
    for (int i = 0; i < p->lay.rank; ++i) p->lengths[i] = lengths[i];
    p->gamma = gamma;
    p->cfl   = cfl;
    compute_dx(p);
    p->carry_valid = false;
    return AMRB_OK;
}

// ------------------------------------------------------------------------------ topology upload
amrb_status amrb_pool_set_topology(amrb_pool* p, size_t n_owned, size_t n_total,
                                   const int32_t* levels, const int8_t* rel, const int32_t* nbr,
                                   const int8_t* quad)
{
    if (!p || !levels || !rel || !nbr || !quad) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (n_owned == 0 || n_total < n_owned) return fail(AMRB_ERR_ARGUMENT, "bad patch counts");
    if (n_total > p->capacity) return fail(AMRB_ERR_CAPACITY, "patch count exceeds pool capacity");
    AMRB_TRY(set_device(p));
    const int R = p->lay.rank, ND = 2 * R, KF = 1 << (R - 1);
    // compact device form: int32 nbr[P][ND][KF]; uint8 meta[P][ND] = rel | quadrant bits << 2
    std::vector<uint8_t> meta(n_owned * ND);
    for (size_t i = 0; i < n_owned * ND; ++i)
    {
        const int r = rel[i];
        if (r < 0 || r > 3) return fail(AMRB_ERR_ARGUMENT, "relation out of range");
        int m = r;
        if (r == AMRB_REL_COARSER)
            for (int k = 0; k < R; ++k) m |= (quad[i * R + k] & 1) << (2 + k);
        meta[i] = (uint8_t)m;
        const int need = (r == AMRB_REL_FINER) ? KF : (r == AMRB_REL_NONE ? 0 : 1);
        for (int k = 0; k < need; ++k)
        {
            const int32_t n = nbr[i * KF + k];
            if (n < 0 || (size_t)n >= n_total)
                return fail(AMRB_ERR_ARGUMENT, "neighbor index out of range");
        }
    }
    for (size_t i = 0; i < n_owned; ++i)
        if (levels[i] < 0 || levels[i] > p->lay.depth)
            return fail(AMRB_ERR_ARGUMENT, "level out of range");
    if (p->table_cap < n_owned)
    {
        cudaFree(p->d_nbr);
        cudaFree(p->d_meta);
        cudaFree(p->d_level);
        p->d_nbr = nullptr;
        p->d_meta = nullptr;
        p->d_level = nullptr;
        p->table_cap = 0;
        const size_t cap = std::max(n_owned, std::min(p->capacity, n_owned * 2));
        AMRB_CUDA(cudaMalloc(&p->d_nbr, cap * ND * KF * sizeof(int32_t)));
        AMRB_CUDA(cudaMalloc(&p->d_meta, cap * ND + 4)); // + 4: the kernels read the bytes as aligned words
        AMRB_CUDA(cudaMalloc(&p->d_level, cap * sizeof(int32_t)));
        p->table_cap = cap;
    }
    // the previous tables may still be in use by launches in flight
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    AMRB_CUDA(cudaMemcpy(p->d_nbr, nbr, n_owned * ND * KF * sizeof(int32_t), cudaMemcpyHostToDevice));
    AMRB_CUDA(cudaMemcpy(p->d_meta, meta.data(), n_owned * ND, cudaMemcpyHostToDevice));
    AMRB_CUDA(cudaMemcpy(p->d_level, levels, n_owned * sizeof(int32_t), cudaMemcpyHostToDevice));
    p->n_owned     = n_owned;
    p->n_total     = n_total;
    p->carry_valid = false;
    return AMRB_OK;
}


// Tables of the whole leaf set built ON THE DEVICE from the ascending leaf ids: one
// H2D copy of 8 bytes per leaf instead of the host table build and 10 / 28 bytes per
// patch-direction.  Single-GPU form: no ghost slots (n_total == n_owned == n).
amrb_status amrb_pool_set_topology_from_ids(amrb_pool* p, const uint64_t* ids, size_t n)
{
    if (!p || !ids) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (n == 0) return fail(AMRB_ERR_ARGUMENT, "bad patch counts");
    if (n > p->capacity) return fail(AMRB_ERR_CAPACITY, "patch count exceeds pool capacity");
    for (size_t i = 0; i < n; ++i)
    {
        if ((int)(ids[i] & 63u) > p->lay.depth) return fail(AMRB_ERR_ARGUMENT, "level out of range");
        if (i && !(ids[i - 1] < ids[i])) return fail(AMRB_ERR_ARGUMENT, "leaf ids must be strictly ascending");
    }
    AMRB_TRY(set_device(p));
    const int R = p->lay.rank, ND = 2 * R, KF = 1 << (R - 1);
    if (p->table_cap < n)
    {
        cudaFree(p->d_nbr);
        cudaFree(p->d_meta);
        cudaFree(p->d_level);
        p->d_nbr = nullptr;
        p->d_meta = nullptr;
        p->d_level = nullptr;
        p->table_cap = 0;
        const size_t cap = std::max(n, std::min(p->capacity, n * 2));
        AMRB_CUDA(cudaMalloc(&p->d_nbr, cap * ND * KF * sizeof(int32_t)));
        AMRB_CUDA(cudaMalloc(&p->d_meta, cap * ND + 4)); // + 4: the kernels read the bytes as aligned words
        AMRB_CUDA(cudaMalloc(&p->d_level, cap * sizeof(int32_t)));
        p->table_cap = cap;
    }
    if (p->ids_cap < n)
    {
        cudaFree(p->d_ids);
        p->d_ids   = nullptr;
        p->ids_cap = 0;
        const size_t cap = std::max(n, std::min(p->capacity, n * 2));
        AMRB_CUDA(cudaMalloc(&p->d_ids, cap * sizeof(uint64_t)));
        p->ids_cap = cap;
    }
    // the previous tables may still be in use by launches in flight
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    AMRB_CUDA(cudaMemcpyAsync(p->d_ids, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, p->stream));
    const long threads = (long)n * ND;
    topology_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, p->stream>>>(
        p->d_ids, (int)n, R, p->lay.depth, p->d_level, p->d_meta, p->d_nbr);
    AMRB_TRY(check_launch(p, "topology_kernel"));
    AMRB_CUDA(cudaStreamSynchronize(p->stream)); // `ids` may be released by the caller
    p->n_owned     = n;
    p->n_total     = n;
    p->carry_valid = false;
    return AMRB_OK;
}

// device tables back to the host in the compact device form (tests, debugging):
// levels[n], meta[n][2R] = relation | quadrant bits << 2, nbr[n][2R][2^(R-1)]
amrb_status amrb_pool_get_tables(amrb_pool* p, int32_t* levels, uint8_t* meta, int32_t* nbr)
{
    if (!p || !levels || !meta || !nbr) return fail(AMRB_ERR_ARGUMENT, "null argument");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    const int R = p->lay.rank, ND = 2 * R, KF = 1 << (R - 1);
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    AMRB_CUDA(cudaMemcpy(levels, p->d_level, p->n_owned * sizeof(int32_t), cudaMemcpyDeviceToHost));
    AMRB_CUDA(cudaMemcpy(meta, p->d_meta, p->n_owned * ND, cudaMemcpyDeviceToHost));
    AMRB_CUDA(cudaMemcpy(nbr, p->d_nbr, p->n_owned * ND * KF * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return AMRB_OK;
}

// ------------------------------------------------------------------------------ host <-> device
static amrb_status check_range(const amrb_pool* p, int field, size_t first, size_t n, const void* host)
{
    if (!p || !host) return fail(AMRB_ERR_ARGUMENT, "null pool/host");
    if (field < 0 || field >= p->lay.nvar) return fail(AMRB_ERR_ARGUMENT, "field out of range");
    if (first + n > p->capacity) return fail(AMRB_ERR_CAPACITY, "patch range exceeds capacity");
    return AMRB_OK;
}

// patches per staging chunk of the padded <-> interior-only conversions (bounded staging memory for
// pools of millions of patches)
static size_t stage_chunk(const amrb_pool* p)
{
    const size_t target = (size_t)256 << 20; // bytes
    return std::max<size_t>(1, target / (p->pflat * sizeof(double)));
}

amrb_status amrb_pool_upload(amrb_pool* p, int field, size_t first, size_t n, const double* host)
{
    AMRB_TRY(check_range(p, field, first, n, host));
    AMRB_TRY(set_device(p));
    if (p->dense)
    {
        // host patches are padded, the pool keeps interiors only: stage chunks of padded patches and
        // scatter their interiors (ghost values of the host image are dropped: the device gathers them)
        const size_t chunk = stage_chunk(p);
        for (size_t s = 0; s < n; s += chunk)
        {
            const size_t m = std::min(chunk, n - s);
            AMRB_TRY(ensure_stage(p, m * p->pflat));
            AMRB_CUDA(cudaMemcpyAsync(p->d_stage, host + s * p->pflat, m * p->pflat * sizeof(double),
                                      cudaMemcpyHostToDevice, p->stream));
            p->ops->interior(p->stream, p->d_stage, p->cur.p[field] + (first + s) * p->flat, (int)m, 0);
            AMRB_TRY(check_launch(p, "interior_copy_kernel"));
            AMRB_CUDA(cudaStreamSynchronize(p->stream));
        }
        p->carry_valid = false;
        return AMRB_OK;
    }
    AMRB_CUDA(cudaMemcpyAsync(p->cur.p[field] + first * p->flat, host, n * p->flat * sizeof(double),
                              cudaMemcpyHostToDevice, p->stream));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    p->carry_valid = false;
    return AMRB_OK;
}
amrb_status amrb_pool_upload_next(amrb_pool* p, int field, size_t first, size_t n, const double* host)
{
    AMRB_TRY(check_range(p, field, first, n, host));
    AMRB_TRY(set_device(p));
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is open");
    if (p->dense)
    {
        const size_t chunk = stage_chunk(p);
        for (size_t s = 0; s < n; s += chunk)
        {
            const size_t m = std::min(chunk, n - s);
            AMRB_TRY(ensure_stage(p, m * p->pflat));
            AMRB_CUDA(cudaMemcpyAsync(p->d_stage, host + s * p->pflat, m * p->pflat * sizeof(double),
                                      cudaMemcpyHostToDevice, p->stream));
            p->ops->interior(p->stream, p->d_stage, p->nxt.p[field] + (first + s) * p->flat, (int)m, 0);
            AMRB_TRY(check_launch(p, "interior_copy_kernel"));
            AMRB_CUDA(cudaStreamSynchronize(p->stream));
        }
        return AMRB_OK;
    }
    AMRB_CUDA(cudaMemcpyAsync(p->nxt.p[field] + first * p->flat, host, n * p->flat * sizeof(double),
                              cudaMemcpyHostToDevice, p->stream));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    return AMRB_OK;
}
amrb_status amrb_pool_download(amrb_pool* p, int field, size_t first, size_t n, double* host)
{
    AMRB_TRY(check_range(p, field, first, n, host));
    AMRB_TRY(set_device(p));
    if (p->dense)
    {
        // padded image = interior + face ghosts gathered on the fly (what halo_kernel would have
        // materialised), zeros in the edge / corner ghosts
        const size_t chunk = stage_chunk(p);
        const int    tabled = p->d_nbr ? (int)p->n_owned : 0;
        for (size_t s = 0; s < n; s += chunk)
        {
            const size_t m = std::min(chunk, n - s);
            AMRB_TRY(ensure_stage(p, m * p->pflat));
            p->ops->export_padded(p->stream, p->cur.p[field], p->d_nbr, p->d_meta, (int)(first + s), (int)m,
                                  tabled, p->d_stage);
            AMRB_TRY(check_launch(p, "dense_export_kernel"));
            AMRB_CUDA(cudaMemcpyAsync(host + s * p->pflat, p->d_stage, m * p->pflat * sizeof(double),
                                      cudaMemcpyDeviceToHost, p->stream));
            AMRB_CUDA(cudaStreamSynchronize(p->stream));
        }
        return AMRB_OK;
    }
    AMRB_TRY(amrb_pool_ensure_halos(p));
    AMRB_CUDA(cudaMemcpyAsync(host, p->cur.p[field] + first * p->flat, n * p->flat * sizeof(double),
                              cudaMemcpyDeviceToHost, p->stream));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    return AMRB_OK;
}
amrb_status amrb_pool_upload_interior(amrb_pool* p, int field, size_t first, size_t n,
                                      const double* host)
{
    AMRB_TRY(check_range(p, field, first, n, host));
    AMRB_TRY(set_device(p));
    if (n == 0) return AMRB_OK;
    if (p->dense)
    {
        // the host array already has the pool's layout: one bulk copy
        AMRB_CUDA(cudaMemcpyAsync(p->cur.p[field] + first * p->flat, host, n * p->flat * sizeof(double),
                                  cudaMemcpyHostToDevice, p->stream));
        AMRB_CUDA(cudaStreamSynchronize(p->stream));
        p->carry_valid = false;
        return AMRB_OK;
    }
    const size_t chunk = stage_chunk(p);
    for (size_t s = 0; s < n; s += chunk)
    {
        const size_t m = std::min(chunk, n - s);
        AMRB_TRY(ensure_stage(p, m * p->data));
        AMRB_CUDA(cudaMemcpyAsync(p->d_stage, host + s * p->data, m * p->data * sizeof(double),
                                  cudaMemcpyHostToDevice, p->stream));
        p->ops->interior(p->stream, p->cur.p[field] + (first + s) * p->flat, p->d_stage, (int)m, 1);
        AMRB_TRY(check_launch(p, "interior_copy_kernel"));
        AMRB_CUDA(cudaStreamSynchronize(p->stream));
    }
    p->carry_valid = false;
    return AMRB_OK;
}
amrb_status amrb_pool_download_interior(amrb_pool* p, int field, size_t first, size_t n,
                                        double* host)
{
    AMRB_TRY(check_range(p, field, first, n, host));
    AMRB_TRY(set_device(p));
    if (n == 0) return AMRB_OK;
    if (p->dense)
    {
        AMRB_CUDA(cudaMemcpyAsync(host, p->cur.p[field] + first * p->flat, n * p->flat * sizeof(double),
                                  cudaMemcpyDeviceToHost, p->stream));
        AMRB_CUDA(cudaStreamSynchronize(p->stream));
        return AMRB_OK;
    }
    const size_t chunk = stage_chunk(p);
    for (size_t s = 0; s < n; s += chunk)
    {
        const size_t m = std::min(chunk, n - s);
        AMRB_TRY(ensure_stage(p, m * p->data));
        p->ops->interior(p->stream, p->cur.p[field] + (first + s) * p->flat, p->d_stage, (int)m, 0);
        AMRB_TRY(check_launch(p, "interior_copy_kernel"));
        AMRB_CUDA(cudaMemcpyAsync(host + s * p->data, p->d_stage, m * p->data * sizeof(double),
                                  cudaMemcpyDeviceToHost, p->stream));
        AMRB_CUDA(cudaStreamSynchronize(p->stream));
    }
    return AMRB_OK;
}

// ------------------------------------------------------------------------------ halo
amrb_status amrb_pool_halo_exchange(amrb_pool* p)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    p->halos_stale = false;
    if (p->dense) return AMRB_OK; // no stored ghosts: every reader gathers them from the interiors
    p->ops->halo_fill(p->stream, p->cur, p->d_nbr, p->d_meta, (int)p->n_owned);
    return check_launch(p, "halo_kernel");
}

// current <-> next, for callers that fill the next buffer themselves (patch migration between the
// Morton ranges of a sharded mesh writes the incoming patches there); the tables stay as they are
amrb_status amrb_pool_swap_buffers(amrb_pool* p)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is open");
    swap_buffers(p);
    p->carry_valid = false;
    p->halos_stale = false;
    return AMRB_OK;
}

// Lazy materialisation of the face halos (drop-in headers): the fused step never reads stored ghosts, so
// the post-condition of a reference step "face halos of the current buffer are filled"
// (amr_solver.hpp:351-352) only has to hold when something OBSERVES the padded patches: a download, the
// refinement criterion, a raw device pointer handed out, the refine / coarsen data motion.  A driver
// that advances one step per call (the reference's benchmarks) otherwise pays one halo launch per step
// (19 % of the GPU time of bench_fvm_solver_integration_active_amr, profiles/r01v_*).
amrb_status amrb_pool_set_lazy_halos(amrb_pool* p, int on)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    p->lazy_halos = on != 0;
    return AMRB_OK;
}
amrb_status amrb_pool_ensure_halos(amrb_pool* p)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (!p->halos_stale || !p->d_nbr || p->n_owned == 0) return AMRB_OK;
    return amrb_pool_halo_exchange(p);
}

// ------------------------------------------------------------------------------ stepping
amrb_status amrb_pool_compute_dt(amrb_pool* p, double* dt_out)
{
    if (!p || !dt_out) return fail(AMRB_ERR_ARGUMENT, "null argument");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    if (p->batch_open || p->batch_pending) return fail(AMRB_ERR_STATE, "a batch is in flight");
    AMRB_TRY(ensure_scalars(p, 1));
    init_scalars_kernel<<<1, 32, 0, p->stream>>>(p->d_dtmin, 0, 1, nullptr, 0.0, nullptr);
    AMRB_TRY(check_launch(p, "init_scalars_kernel"));
    StepArgs a;
    fill_step_args(p, a);
    p->ops->compute_dt(p->stream, a, p->d_dtmin);
    AMRB_TRY(check_launch(p, "compute_dt_kernel"));
    double raw = 0.0;
    AMRB_CUDA(cudaMemcpyAsync(&raw, p->d_dtmin, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    *dt_out        = p->cfl * raw; // amr_solver.hpp:412
    p->carry_valid = false;
    return AMRB_OK;
}

amrb_status amrb_pool_step(amrb_pool* p, double dt)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is open");
    if (p->mode == 1)
    {
        p->ops->halo_fill(p->stream, p->cur, p->d_nbr, p->d_meta, (int)p->n_owned);
        AMRB_TRY(check_launch(p, "halo_kernel"));
    }
    StepArgs a;
    fill_step_args(p, a);
    a.sc.fixed_dt = dt;
    (p->mode == 2 ? p->ops->step_v1 : p->ops->step)(p->stream, a, (int)p->n_owned);
    AMRB_TRY(check_launch(p, "step_kernel"));
    swap_buffers(p);
    p->carry_valid = false;
    p->halos_stale = true; // the ghosts of the new current buffer are copies / two steps old
    return AMRB_OK;
}

amrb_status amrb_pool_batch_begin(amrb_pool* p, size_t max_steps, double remaining)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is already open");
    if (p->batch_pending)
    {
        // same as wait_for_pending_dt_copy() at the top of advance_batch_async (amr_solver.hpp:163)
        AMRB_TRY(wait_batch(p));
    }
    const size_t last = p->batch_steps; // slot holding the dt-min of the state we start from
    const bool   carry = p->carry_valid && p->scal_cap >= max_steps + 2;
    if (carry && last != 0)
        AMRB_CUDA(cudaMemcpyAsync(p->d_dtmin, p->d_dtmin + last, sizeof(unsigned long long),
                                  cudaMemcpyDeviceToDevice, p->stream));
    AMRB_TRY(ensure_scalars(p, max_steps));
    const int first = carry ? 1 : 0;
    const int count = (int)max_steps + 1 - first;
    init_scalars_kernel<<<(count + 255) / 256 + 1, 256, 0, p->stream>>>(
        p->d_dtmin, first, count, p->d_remaining, remaining, p->d_dts);
    AMRB_TRY(check_launch(p, "init_scalars_kernel"));
    if (!carry)
    {
        StepArgs a;
        fill_step_args(p, a);
        p->ops->compute_dt(p->stream, a, p->d_dtmin);
        AMRB_TRY(check_launch(p, "compute_dt_kernel"));
    }
    p->batch_steps  = max_steps;
    p->batch_k      = 0;
    p->batch_open   = true;
    p->step_touched = false;
    p->carry_valid  = false;
    return AMRB_OK;
}

amrb_status amrb_pool_step_partial(amrb_pool* p, const int32_t* dev_list, size_t count)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (!p->batch_open) return fail(AMRB_ERR_STATE, "no open batch");
    if (p->batch_k >= p->batch_steps) return fail(AMRB_ERR_STATE, "batch step budget exhausted");
    AMRB_TRY(set_device(p));
    const size_t k = p->batch_k;
    if (p->mode == 1 && !p->step_touched)
    {
        p->ops->halo_fill(p->stream, p->cur, p->d_nbr, p->d_meta, (int)p->n_owned);
        AMRB_TRY(check_launch(p, "halo_kernel"));
    }
    StepArgs a;
    fill_step_args(p, a);
    a.list = dev_list;
    a.sc   = StepScalars{ p->d_dtmin + k,         p->d_dtmin + k + 1, p->d_remaining + k,
                          p->d_remaining + k + 1, p->d_dts + k,       0.0,
                          p->cfl };
    const int n_items = dev_list ? (int)count : (int)p->n_owned;
    if (n_items > 0)
    {
        // 3D Euler (plane-marching kernel): tasks beyond the first of each warp are drawn from a
        // device counter, zeroed in stream order before the launch (AMRB_QUEUE=0: static map)
        static const int use_queue = getenv("AMRB_QUEUE") ? atoi(getenv("AMRB_QUEUE")) : 1;
        if (use_queue && p->mode == 0 && p->lay.rank == 3 && p->lay.equation == AMRB_EQ_EULER)
        {
            if (!p->d_queue) AMRB_CUDA(cudaMalloc(&p->d_queue, sizeof(unsigned int)));
            AMRB_CUDA(cudaMemsetAsync(p->d_queue, 0, sizeof(unsigned int), p->stream));
            a.queue = p->d_queue;
        }
        (p->mode == 2 ? p->ops->step_v1 : p->ops->step)(p->stream, a, n_items);
        AMRB_TRY(check_launch(p, "step_kernel"));
    }
    p->step_touched = true;
    return AMRB_OK;
}

amrb_status amrb_pool_step_commit(amrb_pool* p)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (!p->batch_open || !p->step_touched) return fail(AMRB_ERR_STATE, "no step to commit");
    swap_buffers(p);
    ++p->batch_k;
    p->step_touched = false;
    return AMRB_OK;
}

amrb_status amrb_pool_batch_end(amrb_pool* p, int materialise_halos)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (!p->batch_open || p->step_touched) return fail(AMRB_ERR_STATE, "batch not open or step uncommitted");
    AMRB_TRY(set_device(p));
    if (materialise_halos && !p->dense)
    {
        // post-condition of every reference step: face halos of the current buffer are filled
        // (amr_solver.hpp:351-352)
        p->ops->halo_fill(p->stream, p->cur, p->d_nbr, p->d_meta, (int)p->n_owned);
        AMRB_TRY(check_launch(p, "halo_kernel"));
        p->halos_stale = false;
    }
    else if (p->dense)
        p->halos_stale = false;
    else
        p->halos_stale = true;
    p->batch_steps = p->batch_k; // slot index of the dt-min of the final state
    if (p->batch_k > 0)
        AMRB_CUDA(cudaMemcpyAsync(p->h_dts, p->d_dts, p->batch_k * sizeof(double),
                                  cudaMemcpyDeviceToHost, p->stream));
    // Under stream capture (the caller records the batch into a CUDA graph) an event recorded here
    // could not be waited for afterwards: the wait falls back to a stream synchronize.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    AMRB_CUDA(cudaStreamIsCapturing(p->stream, &cap));
    p->pending_in_graph = (cap != cudaStreamCaptureStatusNone);
    if (!p->pending_in_graph) AMRB_CUDA(cudaEventRecord(p->batch_done, p->stream));
    p->batch_open    = false;
    p->batch_pending = true;
    p->carry_valid   = true;
    return AMRB_OK;
}

double* amrb_pool_dtmin_slot(amrb_pool* p, size_t k)
{
    if (!p || !p->d_dtmin || k + 1 > p->scal_cap) return nullptr;
    return reinterpret_cast<double*>(p->d_dtmin + k);
}

amrb_status amrb_pool_advance_batch_async(amrb_pool* p, size_t steps, double remaining)
{
    AMRB_TRY(amrb_pool_batch_begin(p, steps, remaining));
    for (size_t k = 0; k < steps; ++k)
    {
        amrb_status s = amrb_pool_step_partial(p, nullptr, 0);
        if (s == AMRB_OK) s = amrb_pool_step_commit(p);
        if (s != AMRB_OK)
        {
            // close the batch: a failed launch must not leave the pool refusing every later call
            p->batch_open   = false;
            p->step_touched = false;
            p->carry_valid  = false;
            p->halos_stale  = true;
            return s;
        }
    }
    if (p->lazy_halos && p->mode != 1) return amrb_pool_batch_end(p, 0);
    return amrb_pool_batch_end(p, 1);
}

amrb_status amrb_pool_finish_advance_batch(amrb_pool* p, double* dt_sum, size_t* executed,
                                           double* dts, size_t dts_capacity)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (p->batch_open) return fail(AMRB_ERR_STATE, "batch still open");
    double sum = 0.0;
    size_t cnt = 0;
    if (p->batch_pending)
    {
        AMRB_TRY(set_device(p));
        AMRB_TRY(wait_batch(p));
    }
    // accumulate exactly like finalize_step_dt_kernel: acc += step_dt in step order, count the
    // steps with dt > 0 (src/cuda/fvm_time_step.cu:224-230)
    for (size_t k = 0; k < p->batch_k; ++k)
    {
        const double d = p->h_dts[k];
        sum += d;
        if (d > 0.0)
        {
            if (dts && cnt < dts_capacity) dts[cnt] = d;
            ++cnt;
        }
    }
    if (dt_sum) *dt_sum = sum;
    if (executed) *executed = cnt;
    return AMRB_OK;
}

// ------------------------------------------------------------------------------ ghost faces
size_t amrb_pool_face_slab_doubles(const amrb_pool* p, int direction)
{
    if (!p || direction < 0 || direction >= 2 * p->lay.rank) return 0;
    size_t n = (size_t)std::min(2 * p->lay.halo, p->lay.size[0]); // layers, see face_pack_kernel
    for (int k = 1; k < p->lay.rank; ++k) n *= (size_t)p->lay.size[0];
    return n;
}

static amrb_status faces(amrb_pool* p, const int32_t* entries, size_t count, double* buffer, int unpack)
{
    if (!p || (count && (!entries || !buffer))) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (count == 0) return AMRB_OK;
    AMRB_TRY(set_device(p));
    p->ops->faces(p->stream, p->cur, entries, (int)count, buffer, unpack);
    return check_launch(p, "face_pack_kernel");
}
amrb_status amrb_pool_pack_faces(amrb_pool* p, const int32_t* e, size_t n, double* b)
{
    return faces(p, e, n, b, 0);
}
amrb_status amrb_pool_unpack_faces(amrb_pool* p, const int32_t* e, size_t n, const double* b)
{
    return faces(p, e, n, const_cast<double*>(b), 1);
}

// ------------------------------------------------------------------------------ reconstruct
amrb_status amrb_pool_apply_plan(amrb_pool* p, size_t new_size, const int8_t* kind,
                                 const int32_t* src, const int8_t* child)
{
    if (!p || !kind || !src || !child) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (new_size == 0 || new_size > p->capacity) return fail(AMRB_ERR_CAPACITY, "new size exceeds capacity");
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is open");
    AMRB_TRY(set_device(p));
    AMRB_TRY(amrb_pool_ensure_halos(p)); // copied patches carry their halos along
    const size_t bytes = new_size * (sizeof(int8_t) * 2 + sizeof(int32_t));
    if (p->plan_cap < bytes)
    {
        cudaFree(p->d_plan);
        p->plan_cap = 0;
        AMRB_CUDA(cudaMalloc(&p->d_plan, bytes * 2));
        p->plan_cap = bytes * 2;
    }
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    int32_t* d_src   = static_cast<int32_t*>(p->d_plan);
    int8_t*  d_kind  = reinterpret_cast<int8_t*>(d_src + new_size);
    int8_t*  d_child = d_kind + new_size;
    AMRB_CUDA(cudaMemcpy(d_src, src, new_size * sizeof(int32_t), cudaMemcpyHostToDevice));
    AMRB_CUDA(cudaMemcpy(d_kind, kind, new_size, cudaMemcpyHostToDevice));
    AMRB_CUDA(cudaMemcpy(d_child, child, new_size, cudaMemcpyHostToDevice));
    p->ops->plan(p->stream, p->cur, p->nxt, d_kind, d_src, d_child, (int)new_size);
    AMRB_TRY(check_launch(p, "plan_kernel"));
    swap_buffers(p);
    p->carry_valid = false;
    // tables are stale until the caller uploads the new topology
    p->n_owned = 0;
    return AMRB_OK;
}

amrb_status amrb_pool_patch_max_flags(amrb_pool* p, int field, double refine_threshold,
                                      double coarsen_threshold, int min_level, int max_level,
                                      int8_t* flags)
{
    if (!p || !flags) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (field < 0 || field >= p->lay.nvar) return fail(AMRB_ERR_ARGUMENT, "field out of range");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    AMRB_TRY(amrb_pool_ensure_halos(p)); // the criterion looks at the padded patch
    if (p->flags_cap < p->n_owned)
    {
        cudaFree(p->d_flags);
        p->flags_cap = 0;
        AMRB_CUDA(cudaMalloc(&p->d_flags, p->n_owned * 2));
        p->flags_cap = p->n_owned * 2;
    }
    if (p->dense)
        p->ops->flags_dense(p->stream, p->cur.p[field], p->d_nbr, p->d_meta, p->d_level, (int)p->n_owned,
                            refine_threshold, coarsen_threshold, min_level, max_level, p->d_flags);
    else
        p->ops->flags(p->stream, p->cur.p[field], p->d_level, (int)p->n_owned, refine_threshold,
                      coarsen_threshold, min_level, max_level, p->d_flags);
    AMRB_TRY(check_launch(p, "patch_max_flags_kernel"));
    AMRB_CUDA(cudaMemcpyAsync(flags, p->d_flags, p->n_owned, cudaMemcpyDeviceToHost, p->stream));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    return AMRB_OK;
}

// the same criterion, flags left on the device (input of amrb_pool_reconstruct_device); asynchronous
amrb_status amrb_pool_flag_patches(amrb_pool* p, int field, double refine_threshold, double coarsen_threshold,
                                   int min_level, int max_level)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (field < 0 || field >= p->lay.nvar) return fail(AMRB_ERR_ARGUMENT, "field out of range");
    AMRB_TRY(need_topology(p));
    AMRB_TRY(set_device(p));
    AMRB_TRY(amrb_pool_ensure_halos(p));
    if (p->flags_cap < p->n_owned)
    {
        cudaFree(p->d_flags);
        p->flags_cap = 0;
        AMRB_CUDA(cudaMalloc(&p->d_flags, p->n_owned * 2));
        p->flags_cap = p->n_owned * 2;
    }
    if (p->dense)
        p->ops->flags_dense(p->stream, p->cur.p[field], p->d_nbr, p->d_meta, p->d_level, (int)p->n_owned,
                            refine_threshold, coarsen_threshold, min_level, max_level, p->d_flags);
    else
        p->ops->flags(p->stream, p->cur.p[field], p->d_level, (int)p->n_owned, refine_threshold,
                      coarsen_threshold, min_level, max_level, p->d_flags);
    return check_launch(p, "patch_max_flags_kernel");
}

amrb_status amrb_patch_max_flags_device(const double* dev_field, const int32_t* dev_levels,
                                        size_t num_patches, size_t cells_per_patch,
                                        double refine_threshold, double coarsen_threshold,
                                        int min_level, int max_level, int8_t* dev_decisions,
                                        void* stream)
{
    if (num_patches == 0) return AMRB_OK;
    if (!dev_field || !dev_levels || !dev_decisions || cells_per_patch == 0)
        return fail(AMRB_ERR_ARGUMENT, "FVM CUDA AMR inputs are smaller than the patch count");
    patch_max_flags_rt_kernel<<<(unsigned)num_patches, 128, 0, (cudaStream_t)stream>>>(
        dev_field, dev_levels, (int)num_patches, (int)cells_per_patch, refine_threshold,
        coarsen_threshold, min_level, max_level, dev_decisions);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("scalar_patch_amr_kernel launch: ") + cudaGetErrorString(e));
    return AMRB_OK;
}

const int32_t* amrb_pool_levels(const amrb_pool* p) { return p ? p->d_level : nullptr; }

// ------------------------------------------------------------------------------ kernel-level entry points
// (the reference's launch protocol on caller-owned arrays; integration/amrb_shim.cpp)
namespace
{
// the reference's halo_direction_metadata (include/cuda/halo_exchange.hpp:19-26), 36 bytes
struct RefHaloMeta
{
    int32_t neighbor;
    int32_t finer_ids[4];
    int32_t quadrant[3];
    int8_t  relation;
    int8_t  padding[3];
};
static_assert(sizeof(RefHaloMeta) == 36, "reference metadata record");

__global__ void ref_meta_to_tables_kernel(const RefHaloMeta* __restrict__ ref, int count, int rank,
                                          int32_t* __restrict__ nbr, uint8_t* __restrict__ meta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int         KF = 1 << (rank - 1);
    const RefHaloMeta r  = ref[i];
    int               m  = r.relation & 3;
    for (int k = 0; k < KF; ++k) nbr[(size_t)i * KF + k] = -1;
    if (m == 1 || m == 3) nbr[(size_t)i * KF] = r.neighbor;
    if (m == 2)
        for (int k = 0; k < KF; ++k) nbr[(size_t)i * KF + k] = r.finer_ids[k];
    if (m == 3)
        for (int k = 0; k < rank; ++k) m |= (r.quadrant[k] & 1) << (2 + k);
    meta[i] = (uint8_t)m;
}
__global__ void raw_finalize_dt_kernel(double* dt, double* acc, double* remaining, uint32_t* count, double cfl)
{
    // finalize_step_dt_kernel semantics (src/cuda/fvm_time_step.cu:204-233)
    double step = *dt * cfl;
    if (*remaining <= 0.0)
        step = 0.0;
    else if (step > *remaining)
        step = *remaining;
    *dt = step;
    *acc += step;
    if (step > 0.0)
    {
        *remaining -= step;
        ++*count;
    }
}
__global__ void raw_set_double_kernel(double* p, double v) { *p = v; }
__global__ void raw_set_uint32_kernel(uint32_t* p, uint32_t v) { *p = v; }

// grow-only device scratch of the raw entry points (tables converted from the reference's metadata, zeroed
// dummy tables and step scalars); the reference's protocol is single-threaded, default stream
struct RawScratch
{
    int32_t* nbr = nullptr;
    uint8_t* meta = nullptr;
    size_t   cap = 0;      // (patch, direction) records
    double*  scal = nullptr; // [0] remaining = DBL_MAX, [1] dt taken, [2] remaining out
    unsigned int* queue = nullptr;
};
RawScratch g_raw;

amrb_status raw_tables(size_t records, int rank, cudaStream_t st, bool zero)
{
    const int KF = 1 << (rank - 1);
    if (g_raw.cap < records)
    {
        cudaFree(g_raw.nbr);
        cudaFree(g_raw.meta);
        g_raw.cap = 0;
        const size_t cap = records * 2;
        AMRB_CUDA(cudaMalloc(&g_raw.nbr, cap * 4 * sizeof(int32_t)));
        AMRB_CUDA(cudaMalloc(&g_raw.meta, cap + 4));
        g_raw.cap = cap;
    }
    if (!g_raw.scal)
    {
        AMRB_CUDA(cudaMalloc(&g_raw.scal, 4 * sizeof(double)));
        AMRB_CUDA(cudaMalloc(&g_raw.queue, sizeof(unsigned int)));
        const double init[4] = { DBL_MAX, 0.0, 0.0, 0.0 };
        AMRB_CUDA(cudaMemcpy(g_raw.scal, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    if (zero)
    {
        AMRB_CUDA(cudaMemsetAsync(g_raw.meta, 0, records, st));
        AMRB_CUDA(cudaMemsetAsync(g_raw.nbr, 0, records * KF * sizeof(int32_t), st));
    }
    return AMRB_OK;
}

amrb_status raw_ops(const amrb_layout* l, const Ops** out)
{
    if (!l) return fail(AMRB_ERR_ARGUMENT, "null layout");
    if (l->storage != AMRB_STORAGE_PADDED) return fail(AMRB_ERR_UNSUPPORTED, "raw entry points work on padded arrays");
    const Ops* o = find_ops(*l);
    if (!o) return fail(AMRB_ERR_UNSUPPORTED, "no kernels instantiated for this patch shape / equation");
    AMRB_CUDA(o->prepare());
    *out = o;
    return AMRB_OK;
}

void raw_step_args(const amrb_layout* l, const double* root, double gamma, StepArgs& a)
{
    std::memset(&a, 0, sizeof(a));
    a.gamma    = gamma;
    a.task_map = 1;
    for (int lvl = 0; lvl <= kMaxLevel; ++lvl)
        for (int d = 0; d < 3; ++d)
            a.dx[lvl][d] = (d < l->rank) ? root[d] / (double)(1u << std::min(lvl, 30)) : 1.0; // fvm_time_step.cu:54-56
}
} // namespace

amrb_status amrb_raw_halo_exchange(const amrb_layout* layout, double* field_base, const void* ref_metadata,
                                   size_t metadata_count, size_t num_patches, void* stream)
{
    if (num_patches == 0) return AMRB_OK;
    if (!layout || !field_base || !ref_metadata) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (metadata_count != num_patches * 2 * (size_t)layout->rank)
        return fail(AMRB_ERR_ARGUMENT, "Halo exchange CUDA metadata size mismatch"); // halo_exchange.cu:363-366
    // one field at a time: the single-field (advection) instantiation of this patch shape
    amrb_layout one = *layout;
    one.nvar        = 1;
    one.equation    = AMRB_EQ_ADVECTION;
    const Ops* ops  = nullptr;
    AMRB_TRY(raw_ops(&one, &ops));
    cudaStream_t st = (cudaStream_t)stream;
    AMRB_TRY(raw_tables(metadata_count, layout->rank, st, false));
    ref_meta_to_tables_kernel<<<(unsigned)((metadata_count + 255) / 256), 256, 0, st>>>(
        static_cast<const RefHaloMeta*>(ref_metadata), (int)metadata_count, layout->rank, g_raw.nbr, g_raw.meta);
    FieldPtrs f{};
    f.p[0] = field_base;
    ops->halo_fill(st, f, g_raw.nbr, g_raw.meta, (int)num_patches);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("halo_kernel: ") + cudaGetErrorString(e));
    return AMRB_OK;
}

amrb_status amrb_raw_compute_dt(const amrb_layout* layout, const double* const* fields, const int32_t* levels,
                                size_t num_patches, const double* root_cell_size, double gamma, double* dev_dt,
                                void* stream)
{
    if (num_patches == 0) return AMRB_OK;
    if (!layout || !fields || !levels || !root_cell_size || !dev_dt) return fail(AMRB_ERR_ARGUMENT, "null argument");
    const Ops* ops = nullptr;
    AMRB_TRY(raw_ops(layout, &ops));
    cudaStream_t st = (cudaStream_t)stream;
    StepArgs     a;
    raw_step_args(layout, root_cell_size, gamma, a);
    for (int f = 0; f < layout->nvar; ++f) a.cur.p[f] = const_cast<double*>(fields[f]);
    a.level     = levels;
    a.n_patches = (int)num_patches;
    raw_set_double_kernel<<<1, 1, 0, st>>>(dev_dt, DBL_MAX);
    ops->compute_dt(st, a, reinterpret_cast<unsigned long long*>(dev_dt));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("compute_dt_kernel: ") + cudaGetErrorString(e));
    return AMRB_OK;
}

amrb_status amrb_raw_finalize_dt(double* dev_dt, double* dev_accumulator, double* dev_remaining,
                                 uint32_t* dev_step_count, double cfl, void* stream)
{
    if (!dev_dt || !dev_accumulator || !dev_remaining || !dev_step_count) return fail(AMRB_ERR_ARGUMENT, "null argument");
    raw_finalize_dt_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_dt, dev_accumulator, dev_remaining, dev_step_count, cfl);
    cudaGetLastError();
    return AMRB_OK;
}

amrb_status amrb_raw_time_step(const amrb_layout* layout, double* const* in, double* const* out,
                               const int32_t* levels, size_t num_patches, const double* root_cell_size,
                               double gamma, const double* dev_dt, void* stream)
{
    if (num_patches == 0) return AMRB_OK;
    if (!layout || !in || !out || !levels || !root_cell_size || !dev_dt) return fail(AMRB_ERR_ARGUMENT, "null argument");
    const Ops* ops = nullptr;
    AMRB_TRY(raw_ops(layout, &ops));
    cudaStream_t st = (cudaStream_t)stream;
    // zeroed tables: relation "none" everywhere = the stored ghosts of `in` are used as they are
    AMRB_TRY(raw_tables(num_patches * 2 * (size_t)layout->rank, layout->rank, st, true));
    StepArgs a;
    raw_step_args(layout, root_cell_size, gamma, a);
    for (int f = 0; f < layout->nvar; ++f)
    {
        a.cur.p[f] = in[f];
        a.nxt.p[f] = out[f];
    }
    a.nbr       = g_raw.nbr;
    a.meta      = g_raw.meta;
    a.level     = levels;
    a.n_patches = (int)num_patches;
    a.lazy_halo = 0;
    a.variant   = getenv("AMRB_VARIANT") ? atoi(getenv("AMRB_VARIANT")) : 0;
    // the step size is final at *dev_dt: dt = raw * 1.0, never clamped (remaining = DBL_MAX)
    a.sc = StepScalars{ reinterpret_cast<const unsigned long long*>(dev_dt), nullptr, g_raw.scal, g_raw.scal + 2,
                        g_raw.scal + 1, 0.0, 1.0 };
    if (layout->rank == 3 && layout->equation == AMRB_EQ_EULER)
    {
        AMRB_CUDA(cudaMemsetAsync(g_raw.queue, 0, sizeof(unsigned int), st));
        a.queue = g_raw.queue;
    }
    ops->step(st, a, (int)num_patches);
    if (g_prepare_error != cudaSuccess)
    {
        const cudaError_t pe = g_prepare_error;
        g_prepare_error      = cudaSuccess;
        return fail(AMRB_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(pe));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("step kernel: ") + cudaGetErrorString(e));
    return AMRB_OK;
}

amrb_status amrb_raw_set_double(double* dev, double value, void* stream)
{
    if (!dev) return fail(AMRB_ERR_ARGUMENT, "null argument");
    raw_set_double_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev, value);
    cudaGetLastError();
    return AMRB_OK;
}
amrb_status amrb_raw_set_uint32(uint32_t* dev, uint32_t value, void* stream)
{
    if (!dev) return fail(AMRB_ERR_ARGUMENT, "null argument");
    raw_set_uint32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev, value);
    cudaGetLastError();
    return AMRB_OK;
}

amrb_status amrb_profile_capture_start(void)
{
    AMRB_CUDA(cudaProfilerStart());
    return AMRB_OK;
}
amrb_status amrb_profile_capture_stop(void)
{
    AMRB_CUDA(cudaProfilerStop());
    return AMRB_OK;
}
// NVTX v3 is header-only (the injection library is looked up at run time, a no-op without a
// profiler attached): same range names as the reference's scoped_profile_range call sites
// (src/cuda/device_buffer.cu:229-237).
amrb_status amrb_profile_range_push(const char* label)
{
    nvtxRangePushA(label ? label : "");
    return AMRB_OK;
}
amrb_status amrb_profile_range_pop(void)
{
    nvtxRangePop();
    return AMRB_OK;
}

} // extern "C"
