// reconstruct_tree on the device: the refine / coarsen SELECTION — eligibility, 2:1
// ripple, coarsening veto — the new leaf ids in Morton order, the transfer plan, the data motion and the new
// halo tables, without the leaf set or the flags leaving the GPU.  The host only reads back two integers
// (did the mesh change, new leaf count).
//
// Replaces the host side of ndtree::reconstruct_tree (include/ndtree/ndtree.hpp:886-940 update_refine_flags /
// apply_refine_coarsen, 1127-1240 balancing, 1249-1271 reconstruct_tree) with the same set-based rules as
// csrc/amrb_topology.cpp: amrb_tree_reconstruct, evaluated over the halo tables the step kernels already use
// (relation, neighbor indices per (leaf, direction)):
//   refine[i]      flag Refine and level < depth;
//   ripple         a leaf with a COARSER neighbor across any face drags that neighbor into refine (2:1
//                  balance), repeated to the fixed point: every hop goes one level coarser -> <= depth sweeps;
//   recombination  all 2^R children of one parent are leaves (contiguous in Morton order, first child = aligned
//                  anchor) flagged Coarsen ...
//   veto           ... unless an OUTWARD face of a child has a finer neighbor, or a same-level one about to split;
//   new leaves     per old leaf 0 (merged away), 1 (kept / the parent of a merged family, emitted by its first
//                  child) or 2^R (children, ascending child number = ascending id); an exclusive scan of the counts
//                  gives every old leaf its place in the new Morton order (the order is preserved: plan monotone).
// Bit-identical to amrb_tree_reconstruct (tests/test_device_regrid.py: leaf ids and plans over multi-level 2D / 3D
// trees and random flags, several passes).
#include "amrb_pool.h"

#include <algorithm>
#include <cstring>

using namespace amrb;

namespace
{
__device__ __forceinline__ void rg_decode(uint64_t id, int rank, int depth, uint32_t (&c)[3], int& lvl)
{
    lvl              = (int)(id & 63u);
    const uint64_t m = id >> 6;
    c[0] = c[1] = c[2] = 0u;
    for (int b = 0; b <= depth; ++b)
        for (int a = 0; a < rank; ++a) c[a] |= (uint32_t)((m >> (rank * b + a)) & 1ull) << b;
}
__device__ __forceinline__ uint64_t rg_encode(int rank, int depth, const uint32_t (&c)[3], int lvl)
{
    uint64_t m = 0;
    for (int b = 0; b <= depth; ++b)
        for (int a = 0; a < rank; ++a) m |= (uint64_t)((c[a] >> b) & 1u) << (rank * b + a);
    return (m << 6) | (uint64_t)lvl;
}

__global__ void rg_init_kernel(const int8_t* __restrict__ flags, const int32_t* __restrict__ level, int n, int depth,
                               uint8_t* __restrict__ refine, uint8_t* __restrict__ merge, int* __restrict__ info)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0)
    {
        info[0] = 0; // any change
        info[1] = 0; // new leaf count
    }
    if (i >= n) return;
    refine[i] = (flags[i] == AMRB_REFINE && level[i] < depth) ? 1 : 0;
    merge[i]  = 0;
}

// one ripple sweep: thread per (leaf, direction)
__global__ void rg_ripple_kernel(const uint8_t* __restrict__ meta, const int32_t* __restrict__ nbr, int n, int nd,
                                 int kf, uint8_t* refine)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * nd) return;
    const int i = (int)(t / nd);
    if (!refine[i]) return;
    if ((meta[t] & 3) == AMRB_REL_COARSER) refine[nbr[t * kf]] = 1;
}

// recombination candidates + veto: thread per leaf (only first children do work)
__global__ void rg_merge_kernel(const uint64_t* __restrict__ ids, const int8_t* __restrict__ flags,
                                const uint8_t* __restrict__ meta, const int32_t* __restrict__ nbr,
                                const uint8_t* __restrict__ refine, int n, int rank, int depth,
                                uint8_t* __restrict__ merge)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int fan = 1 << rank, nd = 2 * rank, kf = 1 << (rank - 1);
    if (i >= n || i + fan > n || flags[i] != AMRB_COARSEN) return;
    uint32_t c[3];
    int      lvl;
    rg_decode(ids[i], rank, depth, c, lvl);
    if (lvl == 0) return;
    const uint32_t E = 1u << (depth - lvl + 1);
    for (int a = 0; a < rank; ++a)
        if ((c[a] & (E - 1)) != 0) return; // not the first child of its parent
    for (int j = 0; j < fan; ++j)
    {
        uint32_t cc[3] = { c[0], c[1], c[2] };
        for (int a = 0; a < rank; ++a)
            if ((j >> a) & 1) cc[a] += E >> 1;
        if (ids[i + j] != rg_encode(rank, depth, cc, lvl) || flags[i + j] != AMRB_COARSEN) return;
    }
    // veto: an outward face of a child with a finer neighbor, or a same-level one about to split
    for (int j = 0; j < fan; ++j)
        for (int d = 0; d < nd; ++d)
        {
            const int dim = d >> 1, pos = d & 1, ax = rank - 1 - dim;
            if (((j >> ax) & 1) != pos) continue; // inward: a sibling
            const size_t e   = (size_t)(i + j) * nd + d;
            const int    rel = meta[e] & 3;
            if (rel == AMRB_REL_FINER) return;
            if (rel == AMRB_REL_SAME && refine[nbr[e * kf]]) return;
        }
    merge[i] = 1;
}

// new leaves per old leaf
__global__ void rg_count_kernel(const uint64_t* __restrict__ ids, const uint8_t* __restrict__ refine,
                                const uint8_t* __restrict__ merge, int n, int rank, int depth,
                                int32_t* __restrict__ count, int* __restrict__ info)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cnt;
    if (merge[i])
        cnt = 1;
    else
    {
        // member of a merged family?  child number j of this leaf -> its first sibling is j slots before
        uint32_t c[3];
        int      lvl;
        rg_decode(ids[i], rank, depth, c, lvl);
        bool gone = false;
        if (lvl > 0)
        {
            const uint32_t e = 1u << (depth - lvl);
            int            j = 0;
            for (int a = 0; a < rank; ++a) j |= (int)((c[a] / e) & 1u) << a;
            const int first = i - j;
            if (j > 0 && first >= 0 && merge[first])
            {
                // the same parent (merge[first] guarantees the family is contiguous from `first`)
                const uint32_t E = e << 1;
                uint32_t       pc[3], fc[3];
                int            fl;
                rg_decode(ids[first], rank, depth, fc, fl);
                bool same = (fl == lvl);
                for (int a = 0; a < rank; ++a)
                {
                    pc[a] = c[a] & ~(E - 1);
                    same  = same && (pc[a] == fc[a]);
                }
                gone = same;
            }
        }
        cnt = gone ? 0 : (refine[i] ? (1 << rank) : 1);
    }
    count[i] = cnt;
    if (cnt != 1 || merge[i]) atomicOr(&info[0], 1);
}

// exclusive scan, three launches: per-block sums, scan of the block sums (one block), add back
constexpr int kScanBlock = 1024;
__global__ void rg_scan_blocks_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out,
                                      int32_t* __restrict__ block_sums)
{
    __shared__ int32_t s[kScanBlock];
    const int          i = blockIdx.x * kScanBlock + threadIdx.x;
    const int32_t      v = i < n ? in[i] : 0;
    s[threadIdx.x]       = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1)
    {
        const int32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = s[threadIdx.x] - v; // exclusive inside the block
    if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void rg_scan_sums_kernel(int32_t* block_sums, int nb, int* info)
{
    // one block: serial chunks per thread, then a shared-memory scan of the chunk totals
    __shared__ int32_t s[kScanBlock];
    const int          per = (nb + kScanBlock - 1) / kScanBlock;
    const int          b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    int32_t            acc = 0;
    for (int b = b0; b < b1; ++b) acc += block_sums[b];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1)
    {
        const int32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    int32_t run = s[threadIdx.x] - acc; // exclusive prefix of this thread's chunk
    for (int b = b0; b < b1; ++b)
    {
        const int32_t v = block_sums[b];
        block_sums[b]   = run;
        run += v;
    }
    if (threadIdx.x == kScanBlock - 1) info[1] = s[threadIdx.x];
}

// new ids + transfer plan at the scanned offsets
__global__ void rg_emit_kernel(const uint64_t* __restrict__ ids, const uint8_t* __restrict__ merge,
                               const int32_t* __restrict__ count, const int32_t* __restrict__ offs,
                               const int32_t* __restrict__ block_sums, int n, int rank, int depth,
                               uint64_t* __restrict__ new_ids, int32_t* __restrict__ src, int8_t* __restrict__ kind,
                               int8_t* __restrict__ child)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cnt = count[i];
    if (cnt == 0) return;
    const int o = offs[i] + block_sums[i / kScanBlock];
    uint32_t  c[3];
    int       lvl;
    rg_decode(ids[i], rank, depth, c, lvl);
    if (merge[i])
    {
        new_ids[o] = rg_encode(rank, depth, c, lvl - 1); // the first child's anchor is the parent's
        src[o]     = i;
        kind[o]    = 2;
        child[o]   = 0;
    }
    else if (cnt == 1)
    {
        new_ids[o] = ids[i];
        src[o]     = i;
        kind[o]    = 0;
        child[o]   = 0;
    }
    else
    {
        const uint32_t hh = 1u << (depth - lvl - 1);
        for (int j = 0; j < cnt; ++j)
        {
            uint32_t cc[3] = { c[0], c[1], c[2] };
            for (int a = 0; a < rank; ++a)
                if ((j >> a) & 1) cc[a] += hh;
            new_ids[o + j] = rg_encode(rank, depth, cc, lvl + 1);
            src[o + j]     = i;
            kind[o + j]    = 1;
            child[o + j]   = (int8_t)j;
        }
    }
}

struct Regrid // grow-only scratch per pool, kept in a side table keyed by the pool
{
    uint8_t *refine = nullptr, *merge = nullptr;
    int32_t *count = nullptr, *offs = nullptr, *bsum = nullptr;
    uint64_t* new_ids = nullptr;
    int*      d_info = nullptr;
    int*      h_info = nullptr;
    size_t    cap = 0, ids_cap = 0;
};
} // namespace

// scratch lives with the pool (amrb_pool has no member for it: keep a small registry)
#include <map>
#include <mutex>
namespace
{
std::map<amrb_pool*, Regrid> g_regrid;
std::mutex                   g_regrid_mutex;

amrb_status ensure(Regrid& r, amrb_pool* p, size_t n)
{
    if (r.cap < n)
    {
        cudaFree(r.refine);
        cudaFree(r.merge);
        cudaFree(r.count);
        cudaFree(r.offs);
        cudaFree(r.bsum);
        r.cap = 0;
        const size_t cap = std::max(n, std::min(p->capacity, n * 2));
        AMRB_CUDA(cudaMalloc(&r.refine, cap));
        AMRB_CUDA(cudaMalloc(&r.merge, cap));
        AMRB_CUDA(cudaMalloc(&r.count, cap * sizeof(int32_t)));
        AMRB_CUDA(cudaMalloc(&r.offs, cap * sizeof(int32_t)));
        AMRB_CUDA(cudaMalloc(&r.bsum, (cap / kScanBlock + 2) * sizeof(int32_t)));
        r.cap = cap;
    }
    if (!r.d_info)
    {
        AMRB_CUDA(cudaMalloc(&r.d_info, 2 * sizeof(int)));
        AMRB_CUDA(cudaHostAlloc(&r.h_info, 2 * sizeof(int), cudaHostAllocDefault));
    }
    if (r.ids_cap < p->capacity)
    {
        cudaFree(r.new_ids);
        r.ids_cap = 0;
        AMRB_CUDA(cudaMalloc(&r.new_ids, p->capacity * sizeof(uint64_t)));
        r.ids_cap = p->capacity;
    }
    return AMRB_OK;
}
} // namespace

namespace amrb
{
void regrid_release(amrb_pool* p)
{
    std::lock_guard<std::mutex> lock(g_regrid_mutex);
    auto                        it = g_regrid.find(p);
    if (it == g_regrid.end()) return;
    Regrid& r = it->second;
    cudaFree(r.refine);
    cudaFree(r.merge);
    cudaFree(r.count);
    cudaFree(r.offs);
    cudaFree(r.bsum);
    cudaFree(r.new_ids);
    cudaFree(r.d_info);
    if (r.h_info) cudaFreeHost(r.h_info);
    g_regrid.erase(it);
}
} // namespace amrb

extern "C" {

amrb_status amrb_pool_reconstruct_device(amrb_pool* p, const int8_t* dev_flags, int* changed, size_t* new_size)
{
    if (!p) return fail(AMRB_ERR_ARGUMENT, "null pool");
    if (changed) *changed = 0;
    if (new_size) *new_size = p->n_owned;
    if (p->batch_open) return fail(AMRB_ERR_STATE, "a batch is open");
    if (!p->d_nbr || p->n_owned == 0) return fail(AMRB_ERR_STATE, "pool has no topology yet");
    if (!p->d_ids || p->n_total != p->n_owned)
        return fail(AMRB_ERR_STATE, "device reconstruct needs the topology built from the leaf ids "
                                    "(amrb_pool_set_topology_from_ids), without ghost slots");
    const int8_t* flags = dev_flags ? dev_flags : p->d_flags;
    if (!flags) return fail(AMRB_ERR_STATE, "no refinement flags on the device (amrb_pool_flag_patches first)");
    AMRB_CUDA(cudaSetDevice(p->device));
    Regrid* rp;
    {
        std::lock_guard<std::mutex> lock(g_regrid_mutex);
        rp = &g_regrid[p];
    }
    Regrid&   r = *rp;
    const int n = (int)p->n_owned, R = p->lay.rank, depth = p->lay.depth, ND = 2 * R, KF = 1 << (R - 1);
    AMRB_TRY(ensure(r, p, (size_t)n));
    cudaStream_t st = p->stream;
    const int    T = 256, nb1 = (n + T - 1) / T;
    rg_init_kernel<<<nb1, T, 0, st>>>(flags, p->d_level, n, depth, r.refine, r.merge, r.d_info);
    const long nt = (long)n * ND;
    for (int sweep = 0; sweep < depth; ++sweep)
        rg_ripple_kernel<<<(unsigned)((nt + T - 1) / T), T, 0, st>>>(p->d_meta, p->d_nbr, n, ND, KF, r.refine);
    rg_merge_kernel<<<nb1, T, 0, st>>>(p->d_ids, flags, p->d_meta, p->d_nbr, r.refine, n, R, depth, r.merge);
    rg_count_kernel<<<nb1, T, 0, st>>>(p->d_ids, r.refine, r.merge, n, R, depth, r.count, r.d_info);
    const int nsb = (n + kScanBlock - 1) / kScanBlock;
    rg_scan_blocks_kernel<<<nsb, kScanBlock, 0, st>>>(r.count, n, r.offs, r.bsum);
    rg_scan_sums_kernel<<<1, kScanBlock, 0, st>>>(r.bsum, nsb, r.d_info);
    AMRB_CUDA(cudaMemcpyAsync(r.h_info, r.d_info, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    AMRB_CUDA(cudaStreamSynchronize(st));
    p->launches += 5 + depth;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("device reconstruct: ") + cudaGetErrorString(e));
    if (!r.h_info[0]) return AMRB_OK; // nothing refines, nothing recombines
    const size_t m = (size_t)r.h_info[1];
    if (m == 0 || m > p->capacity) return fail(AMRB_ERR_CAPACITY, "reconstruct would exceed the patch capacity");
    // plan buffers: int32 src | int8 kind | int8 child
    const size_t bytes = m * (sizeof(int32_t) + 2);
    if (p->plan_cap < bytes)
    {
        cudaFree(p->d_plan);
        p->plan_cap = 0;
        AMRB_CUDA(cudaMalloc(&p->d_plan, bytes * 2));
        p->plan_cap = bytes * 2;
    }
    int32_t* d_src   = static_cast<int32_t*>(p->d_plan);
    int8_t*  d_kind  = reinterpret_cast<int8_t*>(d_src + m);
    int8_t*  d_child = d_kind + m;
    rg_emit_kernel<<<nb1, T, 0, st>>>(p->d_ids, r.merge, r.count, r.offs, r.bsum, n, R, depth, r.new_ids, d_src,
                                      d_kind, d_child);
    // data motion (copy / prolongation / restriction fused with the re-sort), then the new tables
    AMRB_TRY(amrb_pool_ensure_halos(p)); // copied patches carry their halos along
    p->ops->plan(st, p->cur, p->nxt, d_kind, d_src, d_child, (int)m);
    std::swap(p->cur, p->nxt);
    if (p->table_cap < m || p->ids_cap < m)
    {
        // tables grow with the leaf count (never beyond the pool capacity)
        AMRB_CUDA(cudaStreamSynchronize(st));
        const size_t cap = std::max(m, std::min(p->capacity, m * 2));
        if (p->table_cap < m)
        {
            cudaFree(p->d_nbr);
            cudaFree(p->d_meta);
            cudaFree(p->d_level);
            p->d_nbr = nullptr;
            p->d_meta = nullptr;
            p->d_level = nullptr;
            p->table_cap = 0;
            AMRB_CUDA(cudaMalloc(&p->d_nbr, cap * ND * KF * sizeof(int32_t)));
            AMRB_CUDA(cudaMalloc(&p->d_meta, cap * ND));
            AMRB_CUDA(cudaMalloc(&p->d_level, cap * sizeof(int32_t)));
            p->table_cap = cap;
        }
        if (p->ids_cap < m)
        {
            cudaFree(p->d_ids);
            p->d_ids = nullptr;
            p->ids_cap = 0;
            AMRB_CUDA(cudaMalloc(&p->d_ids, cap * sizeof(uint64_t)));
            p->ids_cap = cap;
        }
    }
    AMRB_CUDA(cudaMemcpyAsync(p->d_ids, r.new_ids, m * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    const long threads = (long)m * ND;
    topology_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p->d_ids, (int)m, R, depth, p->d_level,
                                                                       p->d_meta, p->d_nbr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("device reconstruct (apply): ") + cudaGetErrorString(e));
    p->launches += 3;
    p->n_owned     = m;
    p->n_total     = m;
    p->carry_valid = false;
    p->halos_stale = !p->dense;
    if (changed) *changed = 1;
    if (new_size) *new_size = m;
    return AMRB_OK;
}

// leaf ids of the current device topology (ascending), e.g. to mirror a device reconstruct on the host
amrb_status amrb_pool_get_ids(amrb_pool* p, uint64_t* host_ids, size_t capacity)
{
    if (!p || !host_ids) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (!p->d_ids) return fail(AMRB_ERR_STATE, "the topology was not built from leaf ids");
    if (capacity < p->n_owned) return fail(AMRB_ERR_CAPACITY, "id buffer too small");
    AMRB_CUDA(cudaSetDevice(p->device));
    AMRB_CUDA(cudaMemcpyAsync(host_ids, p->d_ids, p->n_owned * sizeof(uint64_t), cudaMemcpyDeviceToHost, p->stream));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    return AMRB_OK;
}

// last transfer plan of a device reconstruct (tests): kind / src / child of the new leaves
amrb_status amrb_pool_get_plan(amrb_pool* p, int8_t* kind, int32_t* src, int8_t* child)
{
    if (!p || !kind || !src || !child) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (!p->d_plan) return fail(AMRB_ERR_STATE, "no plan on the device");
    const size_t m = p->n_owned;
    AMRB_CUDA(cudaSetDevice(p->device));
    AMRB_CUDA(cudaStreamSynchronize(p->stream));
    const int32_t* d_src = static_cast<const int32_t*>(p->d_plan);
    const int8_t*  d_kind = reinterpret_cast<const int8_t*>(d_src + m);
    AMRB_CUDA(cudaMemcpy(src, d_src, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
    AMRB_CUDA(cudaMemcpy(kind, d_kind, m, cudaMemcpyDeviceToHost));
    AMRB_CUDA(cudaMemcpy(child, d_kind + m, m, cudaMemcpyDeviceToHost));
    return AMRB_OK;
}

} // extern "C"
