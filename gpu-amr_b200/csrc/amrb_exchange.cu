// Inter-GPU ghost exchange over peer-mapped memory (NVLink / NVSwitch), no collective library on the data
// path.  One process per GPU; rank r owns a Morton range and keeps copies of the remote patches its halos
// read in ghost slots (include/gpuamr_b200.h, section 6).  Per time step ONE kernel per rank
//
//     face_push_kernel   gathers the interior slabs the peers' halos need and STORES them straight into the
//                        peers' receive buffers over NVLink (coalesced remote stores), then -- last CTA
//                        done, after a system-scope fence -- drops this rank's CFL minimum into every peer's
//                        mailbox and raises its arrival flag there (pack + send + signal fused);
//
// followed on the receiving side by exchange_unpack_kernel: acquire-spin on the peers' flags, fold of their CFL
// minima into this rank's dt-min slot (= the all-reduce(min)) and the unpack into the ghost slots in one launch
// (exchange_wait_kernel + the pool's unpack kernel do the same in two launches: overlapped schedule, per-phase
// timing, drivers that hold several ranks in one process).  Receive buffers and mailbox slots are double-buffered by generation parity: a peer can be at most
// one exchange ahead (its next push needs this rank's push of the current generation).  The K-step batch
// loop lives here, in C++ (amrb_exchange_advance_batch_async): nothing on the host runs between two steps.
//
// No reference counterpart: the reference is single-GPU.
#include "amrb_pool.h"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace amrb;

namespace
{
constexpr int kMaxWorld = 8;

struct Mailbox // device memory of the RECEIVING rank, written by its peers
{
    unsigned long long flag[kMaxWorld];     // flag[r]  = last DATA generation rank r has pushed here
    unsigned long long tflag[kMaxWorld];    // tflag[r] = last CFL-minimum generation rank r has delivered
    unsigned long long dtmin[2][kMaxWorld]; // bits of rank r's local dt-min, by dt-generation parity
};

struct PushArgs
{
    FieldPtrs                 cur;        // this rank's current buffers
    const int32_t*            entries;    // [count][2] = {patch, direction}, sorted by destination rank
    int                       count;
    int                       world, rank;
    int                       seg_start[kMaxWorld + 1]; // entries for destination r: [seg_start[r], seg_start[r+1])
    double*                   peer_recv[kMaxWorld];     // peer r's receive buffer of this generation + my segment
    Mailbox*                  peer_box[kMaxWorld];      // peer r's mailbox (nullptr: not connected / self)
    unsigned int*             done;       // CTA completion counter (this rank)
    const unsigned long long* my_dtmin;   // this rank's dt-min slot of the step about to run (may be null)
    unsigned long long        gen;        // data generation of this push (0: no slabs, CFL minimum only)
    unsigned long long        tgen;       // dt generation (0: no CFL minimum in this push)
    int                       R, S, HS, NV, T, stored; // geometry: rank, size, stored halo, fields, layers, doubles per field-patch
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// slab of entry e = the T interior layers next to face `direction` of every field, laid out
// [field][layer][face cell] exactly like face_pack_kernel (amrb_kernels.cuh), so the receiver unpacks with
// the same kernel.  Runtime geometry: one instantiation serves every patch shape (bandwidth-trivial kernel).
__global__ void __launch_bounds__(128) face_push_kernel(const __grid_constant__ PushArgs a)
{
    __shared__ bool last;
    const int e = blockIdx.x;
    if (e < a.count)
    {
        const int p = a.entries[2 * e], dw = a.entries[2 * e + 1], d = dw & 15;
        const int dim = d >> 1, pos = d & 1;
        int       dst = 0;
        while (e >= a.seg_start[dst + 1]) ++dst;
        const int P    = a.S + 2 * a.HS;
        const int face = (a.R == 2) ? a.S : a.S * a.S;
        const int slab = a.T * face;
        const int part = ((dw >> 4) > 0 && (dw >> 4) < a.T) ? (dw >> 4) * face : slab; // layers on the wire
        double*   out  = a.peer_recv[dst] + (size_t)(e - a.seg_start[dst]) * a.NV * slab;
        for (int i2 = threadIdx.x; i2 < a.NV * part; i2 += blockDim.x)
        {
            const int f = i2 / part, r = i2 % part, layer = r / face, it = f * slab + r;
            int       t = r % face, gl = 0, pitch = 1;
            for (int k = a.R - 1; k >= 0; --k)
            {
                int i;
                if (k == dim)
                    i = pos ? (a.HS + a.S - 1 - layer) : (a.HS + layer);
                else
                {
                    i = a.HS + (t % a.S);
                    t /= a.S;
                }
                gl += i * pitch;
                pitch *= P;
            }
            out[it] = a.cur.p[f][(size_t)p * a.stored + gl];
        }
    }
    // ---- completion: the last CTA to finish signals every peer.  One system-scope fence per CTA, by the thread
    // that counts the CTA in after the block barrier (the barrier orders the other threads' remote stores before
    // it: the cooperative-groups grid-barrier pattern); 128 fences per CTA made the kernel wait 128 times for the
    // NVLink acknowledgements
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        last = (atomicAdd(a.done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < a.world && threadIdx.x != a.rank && a.peer_box[threadIdx.x] != nullptr)
    {
        __threadfence_system();
        Mailbox* box = a.peer_box[threadIdx.x];
        if (a.tgen != 0)
        {
            box->dtmin[a.tgen & 1][a.rank] = *a.my_dtmin;
            st_release_sys(&box->tflag[a.rank], a.tgen);
        }
        if (a.gen != 0) st_release_sys(&box->flag[a.rank], a.gen);
    }
    if (threadIdx.x == 0) *a.done = 0u; // ready for the next launch (stream-ordered)
}

// one warp: lane r waits for rank r's push of generation `gen`, then the warp folds the peers' dt-minima
// into this rank's slot (bit patterns of positive doubles order like unsigned integers)
__global__ void exchange_wait_kernel(Mailbox* box, int world, int rank, unsigned long long gen,
                                     unsigned long long tgen, unsigned long long* my_dtmin, int* timed_out,
                                     long long timeout_cycles)
{
    const int          lane = threadIdx.x;
    unsigned long long v    = ~0ull;
    if (lane < world && lane != rank)
    {
        const long long t0 = clock64();
        bool            ok = true;
        while (ld_acquire_sys(&box->flag[lane]) < gen || ld_acquire_sys(&box->tflag[lane]) < tgen)
        {
            if (clock64() - t0 > timeout_cycles)
            {
                ok = false;
                break;
            }
            __nanosleep(64);
        }
        if (!ok)
            atomicExch(timed_out, 1);
        else if (my_dtmin != nullptr)
            v = box->dtmin[tgen & 1][lane];
    }
    else if (lane == rank && my_dtmin != nullptr)
        v = *my_dtmin;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v                          = t < v ? t : v;
    }
    if (lane == 0 && my_dtmin != nullptr) *my_dtmin = v;
}

// wait + all-reduce(min) fold + unpack in ONE launch (the plain schedule's receiving side): every CTA's first warp
// acquire-spins on the peers' data flags of this generation, CTA 0 also on the CFL-minimum flags and folds the
// minima into this rank's dt-min slot; then the CTA copies its entry's slab from the receive buffer (read past
// L1: the lines are written by the peers) into the ghost slot.  Same slab layout as face_push_kernel.
struct UnpackArgs
{
    FieldPtrs           cur;
    const int32_t*      entries; // [count][2] = {ghost slot, direction}
    int                 count;
    const double*       buffer;  // my receive buffer of this generation
    Mailbox*            box;
    int                 world, rank;
    unsigned long long  gen, tgen;
    unsigned long long* my_dtmin; // may be null: no CFL minimum in this exchange
    int*                timed_out;
    long long           timeout_cycles;
    int                 R, S, HS, NV, T, stored;
};
__global__ void __launch_bounds__(128) exchange_unpack_kernel(const __grid_constant__ UnpackArgs a)
{
    if (threadIdx.x < 32)
    {
        const int          lane = threadIdx.x;
        const bool         fold = (blockIdx.x == 0) && a.my_dtmin != nullptr;
        unsigned long long v    = ~0ull;
        if (lane < a.world && lane != a.rank)
        {
            const long long t0 = clock64();
            bool            ok = true;
            while (ld_acquire_sys(&a.box->flag[lane]) < a.gen || (fold && ld_acquire_sys(&a.box->tflag[lane]) < a.tgen))
            {
                if (clock64() - t0 > a.timeout_cycles)
                {
                    ok = false;
                    break;
                }
                __nanosleep(64);
            }
            if (!ok)
                atomicExch(a.timed_out, 1);
            else if (fold)
                v = a.box->dtmin[a.tgen & 1][lane];
        }
        else if (lane == a.rank && fold)
            v = *a.my_dtmin;
        if (fold)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
                v                          = t < v ? t : v;
            }
            if (lane == 0) *a.my_dtmin = v;
        }
    }
    __syncthreads();
    const int e = blockIdx.x;
    if (e >= a.count) return;
    const int     p = a.entries[2 * e], dw = a.entries[2 * e + 1], d = dw & 15;
    const int     dim = d >> 1, pos = d & 1;
    const int     P    = a.S + 2 * a.HS;
    const int     face = (a.R == 2) ? a.S : a.S * a.S;
    const int     slab = a.T * face;
    const int     part = ((dw >> 4) > 0 && (dw >> 4) < a.T) ? (dw >> 4) * face : slab;
    const double* in   = a.buffer + (size_t)e * a.NV * slab;
    for (int i2 = threadIdx.x; i2 < a.NV * part; i2 += blockDim.x)
    {
        const int f = i2 / part, r = i2 % part, layer = r / face, it = f * slab + r;
        int       t = r % face, gl = 0, pitch = 1;
        for (int k = a.R - 1; k >= 0; --k)
        {
            int i;
            if (k == dim)
                i = pos ? (a.HS + a.S - 1 - layer) : (a.HS + layer);
            else
            {
                i = a.HS + (t % a.S);
                t /= a.S;
            }
            gl += i * pitch;
            pitch *= P;
        }
        a.cur.p[f][(size_t)p * a.stored + gl] = __ldcg(in + it);
    }
}
} // namespace

struct amrb_exchange
{
    amrb_pool*   pool = nullptr;
    int          rank = 0, world = 1;
    Mailbox*     box  = nullptr;            // mine (device)
    double*      recv[2] = { nullptr, nullptr }; // my receive buffers by generation parity
    size_t       recv_doubles = 0;
    unsigned int* d_done = nullptr;
    int*         h_timeout = nullptr;       // mapped pinned
    int32_t*     d_send = nullptr;          // entries I push
    int32_t*     d_recv = nullptr;          // entries I unpack (ghost slot, direction)
    size_t       n_send = 0, n_recv = 0;
    int          seg_start[kMaxWorld + 1] = {};
    size_t       seg_offset[kMaxWorld] = {}; // where my segment starts inside peer r's receive buffer [doubles]
    Mailbox*     peer_box[kMaxWorld] = {};
    double*      peer_recv[kMaxWorld][2] = {};
    unsigned long long gen = 0;             // data generation (slab pushes)
    unsigned long long tgen = 0;            // dt generation (CFL minima delivered)
    int32_t*     d_boundary = nullptr;      // owned patches with a remote neighbor / without (overlap schedule)
    int32_t*     d_interior = nullptr;
    size_t       n_boundary = 0, n_interior = 0;
    cudaStream_t side = nullptr;            // the slab push of step k+1 runs here, under the interior launch of step k
    cudaEvent_t  ev_boundary = nullptr, ev_push = nullptr;
    int          slab_doubles = 0;          // per entry, all fields
    uint64_t     launches = 0;
    // optional per-phase timing of the plain schedule (amrb_exchange_set_timing): events around push / wait /
    // unpack / step of every step of a batch
    bool                     timing = false;
    std::vector<cudaEvent_t> tev;
    size_t                   tsteps = 0;
};

namespace
{
// one push: the slabs of `src` (data = true) and / or this rank's CFL minimum (dtmin_slot != null) + flags
amrb_status push(amrb_exchange* ex, const unsigned long long* dtmin_slot, bool data = true,
                 const FieldPtrs* src = nullptr, cudaStream_t stream = nullptr)
{
    amrb_pool* p = ex->pool;
    if (data) ++ex->gen;
    if (dtmin_slot) ++ex->tgen;
    PushArgs a{};
    a.cur     = src ? *src : p->cur;
    a.entries = ex->d_send;
    a.count   = data ? (int)ex->n_send : 0;
    a.world   = ex->world;
    a.rank    = ex->rank;
    std::memcpy(a.seg_start, ex->seg_start, sizeof(a.seg_start));
    for (int r = 0; r < ex->world; ++r)
    {
        a.peer_box[r]  = (r == ex->rank) ? nullptr : ex->peer_box[r];
        a.peer_recv[r] = ex->peer_recv[r][ex->gen & 1] ? ex->peer_recv[r][ex->gen & 1] + ex->seg_offset[r] : nullptr;
    }
    a.done     = data ? ex->d_done : ex->d_done + 1; // the two kinds of push may overlap: separate counters
    a.my_dtmin = dtmin_slot;
    a.gen      = data ? ex->gen : 0;
    a.tgen     = dtmin_slot ? ex->tgen : 0;
    a.R        = p->lay.rank;
    a.S        = p->lay.size[0];
    a.HS       = p->dense ? 0 : p->lay.halo;
    a.NV       = p->lay.nvar;
    a.T        = std::min(2 * p->lay.halo, p->lay.size[0]);
    a.stored   = (int)p->flat;
    face_push_kernel<<<(unsigned)std::max<size_t>(a.count, 1), 128, 0, stream ? stream : p->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("face_push_kernel: ") + cudaGetErrorString(e));
    ++ex->launches;
    return AMRB_OK;
}

amrb_status wait_and_unpack(amrb_exchange* ex, unsigned long long* dtmin_slot)
{
    amrb_pool* p = ex->pool;
    // ~4 s at 2 GHz: a peer that died must not hang this GPU
    exchange_wait_kernel<<<1, 32, 0, p->stream>>>(ex->box, ex->world, ex->rank, ex->gen, ex->tgen, dtmin_slot,
                                                  ex->h_timeout, 8000000000ll);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("exchange_wait_kernel: ") + cudaGetErrorString(e));
    ++ex->launches;
    if (ex->n_recv)
    {
        p->ops->faces(p->stream, p->cur, ex->d_recv, (int)ex->n_recv, ex->recv[ex->gen & 1], 1);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("face unpack: ") + cudaGetErrorString(e));
        ++ex->launches;
    }
    return AMRB_OK;
}

// the same in one launch (exchange_unpack_kernel)
amrb_status wait_unpack_fused(amrb_exchange* ex, unsigned long long* dtmin_slot)
{
    amrb_pool* p = ex->pool;
    UnpackArgs a{};
    a.cur            = p->cur;
    a.entries        = ex->d_recv;
    a.count          = (int)ex->n_recv;
    a.buffer         = ex->recv[ex->gen & 1];
    a.box            = ex->box;
    a.world          = ex->world;
    a.rank           = ex->rank;
    a.gen            = ex->gen;
    a.tgen           = ex->tgen;
    a.my_dtmin       = dtmin_slot;
    a.timed_out      = ex->h_timeout;
    a.timeout_cycles = 8000000000ll; // ~4 s at 2 GHz: a peer that died must not hang this GPU
    a.R              = p->lay.rank;
    a.S              = p->lay.size[0];
    a.HS             = p->dense ? 0 : p->lay.halo;
    a.NV             = p->lay.nvar;
    a.T              = std::min(2 * p->lay.halo, p->lay.size[0]);
    a.stored         = (int)p->flat;
    exchange_unpack_kernel<<<(unsigned)std::max<size_t>(ex->n_recv, 1), 128, 0, p->stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMRB_ERR_CUDA, std::string("exchange_unpack_kernel: ") + cudaGetErrorString(e));
    ++ex->launches;
    return AMRB_OK;
}
} // namespace

extern "C" {

amrb_status amrb_exchange_create(amrb_pool* pool, int rank, int world, const int32_t* send_entries,
                                 const int64_t* send_counts, const int64_t* send_offsets,
                                 const int32_t* recv_entries, size_t n_recv, amrb_exchange** out)
{
    if (!pool || !out || !send_counts || !send_offsets) return fail(AMRB_ERR_ARGUMENT, "null argument");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return fail(AMRB_ERR_ARGUMENT, "world size must be 1..8 and 0 <= rank < world");
    AMRB_CUDA(cudaSetDevice(pool->device));
    amrb_exchange* ex = new amrb_exchange();
    ex->pool  = pool;
    ex->rank  = rank;
    ex->world = world;
    size_t n_send = 0;
    for (int r = 0; r < world; ++r)
    {
        ex->seg_start[r] = (int)n_send;
        n_send += (size_t)send_counts[r];
    }
    for (int r = world; r <= kMaxWorld; ++r) ex->seg_start[r] = (int)n_send;
    ex->n_send = n_send;
    ex->n_recv = n_recv;
    if ((n_send && !send_entries) || (n_recv && !recv_entries))
    {
        delete ex;
        return fail(AMRB_ERR_ARGUMENT, "null entry list");
    }
    int T = std::min(2 * pool->lay.halo, pool->lay.size[0]);
    size_t face = 1;
    for (int k = 1; k < pool->lay.rank; ++k) face *= (size_t)pool->lay.size[0];
    ex->slab_doubles = (int)(T * face * pool->lay.nvar);
    for (int r = 0; r < world; ++r) ex->seg_offset[r] = (size_t)send_offsets[r] * ex->slab_doubles;
    ex->recv_doubles = std::max<size_t>(n_recv * ex->slab_doubles, 1);
    AMRB_CUDA(cudaMalloc(&ex->box, sizeof(Mailbox)));
    AMRB_CUDA(cudaMemset(ex->box, 0, sizeof(Mailbox)));
    for (int b = 0; b < 2; ++b)
    {
        AMRB_CUDA(cudaMalloc(&ex->recv[b], ex->recv_doubles * sizeof(double)));
        AMRB_CUDA(cudaMemset(ex->recv[b], 0, ex->recv_doubles * sizeof(double)));
    }
    AMRB_CUDA(cudaMalloc(&ex->d_done, 2 * sizeof(unsigned int)));
    AMRB_CUDA(cudaMemset(ex->d_done, 0, 2 * sizeof(unsigned int)));
    AMRB_CUDA(cudaHostAlloc(&ex->h_timeout, sizeof(int), cudaHostAllocMapped));
    *ex->h_timeout = 0;
    if (n_send)
    {
        AMRB_CUDA(cudaMalloc(&ex->d_send, n_send * 2 * sizeof(int32_t)));
        AMRB_CUDA(cudaMemcpy(ex->d_send, send_entries, n_send * 2 * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    if (n_recv)
    {
        AMRB_CUDA(cudaMalloc(&ex->d_recv, n_recv * 2 * sizeof(int32_t)));
        AMRB_CUDA(cudaMemcpy(ex->d_recv, recv_entries, n_recv * 2 * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    AMRB_CUDA(cudaDeviceSynchronize());
    *out = ex;
    return AMRB_OK;
}

amrb_status amrb_exchange_destroy(amrb_exchange* ex)
{
    if (!ex) return AMRB_OK;
    cudaSetDevice(ex->pool->device);
    cudaStreamSynchronize(ex->pool->stream);
    cudaFree(ex->box);
    cudaFree(ex->recv[0]);
    cudaFree(ex->recv[1]);
    cudaFree(ex->d_done);
    cudaFree(ex->d_send);
    cudaFree(ex->d_recv);
    cudaFree(ex->d_boundary);
    cudaFree(ex->d_interior);
    if (ex->side) cudaStreamDestroy(ex->side);
    if (ex->ev_boundary) cudaEventDestroy(ex->ev_boundary);
    if (ex->ev_push) cudaEventDestroy(ex->ev_push);
    if (ex->h_timeout) cudaFreeHost(ex->h_timeout);
    for (cudaEvent_t e : ex->tev) cudaEventDestroy(e);
    cudaGetLastError();
    delete ex;
    return AMRB_OK;
}

// which: 0 = mailbox, 1 / 2 = receive buffer of generation parity 0 / 1
void* amrb_exchange_buffer(amrb_exchange* ex, int which)
{
    if (!ex) return nullptr;
    return which == 0 ? (void*)ex->box : (which == 1 ? (void*)ex->recv[0] : (which == 2 ? (void*)ex->recv[1] : nullptr));
}

amrb_status amrb_ipc_export(void* dev_ptr, void* handle64)
{
    if (!dev_ptr || !handle64) return fail(AMRB_ERR_ARGUMENT, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    AMRB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle64, &h, 64);
    return AMRB_OK;
}
amrb_status amrb_ipc_open(const void* handle64, void** out)
{
    if (!handle64 || !out) return fail(AMRB_ERR_ARGUMENT, "null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    AMRB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return AMRB_OK;
}
amrb_status amrb_ipc_close(void* dev_ptr)
{
    AMRB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return AMRB_OK;
}

amrb_status amrb_exchange_connect(amrb_exchange* ex, int peer, void* mailbox, void* recv0, void* recv1)
{
    if (!ex || peer < 0 || peer >= ex->world || peer == ex->rank) return fail(AMRB_ERR_ARGUMENT, "bad peer");
    if (!mailbox || !recv0 || !recv1) return fail(AMRB_ERR_ARGUMENT, "null peer buffer");
    ex->peer_box[peer]     = static_cast<Mailbox*>(mailbox);
    ex->peer_recv[peer][0] = static_cast<double*>(recv0);
    ex->peer_recv[peer][1] = static_cast<double*>(recv1);
    return AMRB_OK;
}

// one standalone exchange of the CURRENT buffers' slabs into the peers' ghost slots (before a halo
// materialisation or the first step after an upload); asynchronous
amrb_status amrb_exchange_halo(amrb_exchange* ex)
{
    if (!ex) return fail(AMRB_ERR_ARGUMENT, "null exchange");
    AMRB_CUDA(cudaSetDevice(ex->pool->device));
    AMRB_TRY(push(ex, nullptr));
    return wait_and_unpack(ex, nullptr);
}

// amr_solver::advance_batch_async (solver/amr_solver.hpp:155-243) over the sharded mesh: per step one
// push (slabs + CFL minimum + flag), one wait (+ all-reduce(min) fold), the local unpack, the fused step.
// owned patches with / without a remote neighbor (ShardPlan.boundary / .interior): enables the overlapped
// schedule of amrb_exchange_advance_batch_async
amrb_status amrb_exchange_set_lists(amrb_exchange* ex, const int32_t* boundary, size_t n_boundary,
                                    const int32_t* interior, size_t n_interior)
{
    if (!ex || (n_boundary && !boundary) || (n_interior && !interior)) return fail(AMRB_ERR_ARGUMENT, "null argument");
    AMRB_CUDA(cudaSetDevice(ex->pool->device));
    AMRB_CUDA(cudaStreamSynchronize(ex->pool->stream));
    cudaFree(ex->d_boundary);
    cudaFree(ex->d_interior);
    ex->d_boundary = ex->d_interior = nullptr;
    ex->n_boundary = n_boundary;
    ex->n_interior = n_interior;
    if (n_boundary)
    {
        AMRB_CUDA(cudaMalloc(&ex->d_boundary, n_boundary * sizeof(int32_t)));
        AMRB_CUDA(cudaMemcpy(ex->d_boundary, boundary, n_boundary * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    if (n_interior)
    {
        AMRB_CUDA(cudaMalloc(&ex->d_interior, n_interior * sizeof(int32_t)));
        AMRB_CUDA(cudaMemcpy(ex->d_interior, interior, n_interior * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    if (!ex->side)
    {
        // highest priority: the push CTAs must be dispatched ahead of the persistent CTAs of the interior launch
        // (which take the whole register file of an SM once resident)
        int lo = 0, hi = 0;
        AMRB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        AMRB_CUDA(cudaStreamCreateWithPriority(&ex->side, cudaStreamNonBlocking, hi));
        AMRB_CUDA(cudaEventCreateWithFlags(&ex->ev_boundary, cudaEventDisableTiming));
        AMRB_CUDA(cudaEventCreateWithFlags(&ex->ev_push, cudaEventDisableTiming));
    }
    return AMRB_OK;
}

// amr_solver::advance_batch_async (solver/amr_solver.hpp:155-243) over the sharded mesh.
//   plain schedule (overlap = 0):  per step push (slabs + CFL minimum + flags), wait + all-reduce(min) fold +
//       unpack (one launch; three when per-phase timing is on), the fused step over all owned patches;
//   overlapped schedule (overlap = 1, needs amrb_exchange_set_lists): per step wait + unpack, the fused step over
//       the BOUNDARY patches, then -- on a side stream, under the launch over the INTERIOR patches -- the slab push
//       of the next step straight from the next buffer; the CFL minimum follows in a push of its own once both
//       launches are done.  The NVLink transfer hides behind the interior launch; pays when a rank's share is
//       large (C3: ~250 000 patches per rank), costs two launches per step when it is small.
amrb_status amrb_exchange_advance_batch_async(amrb_exchange* ex, size_t steps, double remaining, int overlap)
{
    if (!ex) return fail(AMRB_ERR_ARGUMENT, "null exchange");
    amrb_pool* p = ex->pool;
    const bool ov = overlap && ex->side && ex->n_boundary > 0 && ex->n_interior > 0 && steps > 0;
    AMRB_TRY(amrb_pool_batch_begin(p, steps, remaining));
    amrb_status s = AMRB_OK;
    auto slot = [&](size_t k) { return reinterpret_cast<unsigned long long*>(amrb_pool_dtmin_slot(p, k)); };
    if (!ov)
    {
        if (ex->timing)
        {
            while (ex->tev.size() < 4 * steps + 1)
            {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) return fail(AMRB_ERR_CUDA, "cudaEventCreate");
                ex->tev.push_back(e);
            }
            ex->tsteps = steps;
            cudaEventRecord(ex->tev[0], p->stream);
        }
        for (size_t k = 0; k < steps && s == AMRB_OK; ++k)
        {
            s = push(ex, slot(k));
            if (ex->timing) cudaEventRecord(ex->tev[4 * k + 1], p->stream);
            if (s == AMRB_OK && !ex->timing)
                s = wait_unpack_fused(ex, slot(k));
            else if (s == AMRB_OK)
            {
                // wait and unpack timed separately
                exchange_wait_kernel<<<1, 32, 0, p->stream>>>(ex->box, ex->world, ex->rank, ex->gen, ex->tgen, slot(k),
                                                              ex->h_timeout, 8000000000ll);
                ++ex->launches;
                if (ex->timing) cudaEventRecord(ex->tev[4 * k + 2], p->stream);
                if (ex->n_recv)
                {
                    p->ops->faces(p->stream, p->cur, ex->d_recv, (int)ex->n_recv, ex->recv[ex->gen & 1], 1);
                    ++ex->launches;
                }
                if (ex->timing) cudaEventRecord(ex->tev[4 * k + 3], p->stream);
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) s = fail(AMRB_ERR_CUDA, std::string("exchange wait / unpack: ") + cudaGetErrorString(e));
            }
            if (s == AMRB_OK) s = amrb_pool_step_partial(p, nullptr, 0);
            if (ex->timing) cudaEventRecord(ex->tev[4 * k + 4], p->stream);
            if (s == AMRB_OK) s = amrb_pool_step_commit(p);
        }
    }
    else
    {
        s = push(ex, slot(0)); // the first exchange has nothing to hide behind
        for (size_t k = 0; k < steps && s == AMRB_OK; ++k)
        {
            s = wait_and_unpack(ex, slot(k));
            // the slab push of the previous step read the buffer this step writes two steps later: done by now
            if (s == AMRB_OK && k > 0 && cudaStreamWaitEvent(p->stream, ex->ev_push, 0) != cudaSuccess)
                s = fail(AMRB_ERR_CUDA, "cudaStreamWaitEvent");
            if (s == AMRB_OK) s = amrb_pool_step_partial(p, ex->d_boundary, ex->n_boundary);
            if (s == AMRB_OK && k + 1 < steps)
            {
                // boundary patches of the new state are in the NEXT buffer: push them while the interior runs
                if (cudaEventRecord(ex->ev_boundary, p->stream) != cudaSuccess ||
                    cudaStreamWaitEvent(ex->side, ex->ev_boundary, 0) != cudaSuccess)
                    s = fail(AMRB_ERR_CUDA, "cudaEventRecord / cudaStreamWaitEvent");
                if (s == AMRB_OK) s = push(ex, nullptr, true, &p->nxt, ex->side);
                if (s == AMRB_OK && cudaEventRecord(ex->ev_push, ex->side) != cudaSuccess)
                    s = fail(AMRB_ERR_CUDA, "cudaEventRecord");
            }
            if (s == AMRB_OK) s = amrb_pool_step_partial(p, ex->d_interior, ex->n_interior);
            if (s == AMRB_OK && k + 1 < steps) s = push(ex, slot(k + 1), false); // CFL minimum of the new state
            if (s == AMRB_OK) s = amrb_pool_step_commit(p);
        }
        if (s == AMRB_OK && steps > 1 && cudaStreamWaitEvent(p->stream, ex->ev_push, 0) != cudaSuccess)
            s = fail(AMRB_ERR_CUDA, "cudaStreamWaitEvent");
    }
    if (s == AMRB_OK) s = push(ex, nullptr);
    if (s == AMRB_OK) s = ov ? wait_and_unpack(ex, nullptr) : wait_unpack_fused(ex, nullptr);
    if (s != AMRB_OK)
    {
        p->batch_open   = false;
        p->step_touched = false;
        p->carry_valid  = false;
        p->halos_stale  = true;
        return s;
    }
    return amrb_pool_batch_end(p, 1);
}

// the two halves of one exchange, for drivers that interleave several ranks themselves (all shards of a mesh
// in ONE process on one GPU: every rank's push has to be enqueued before any rank's wait).  with_dt: fold the
// CFL minimum of batch slot k into the exchange.
amrb_status amrb_exchange_push(amrb_exchange* ex, int with_dt, size_t k)
{
    if (!ex) return fail(AMRB_ERR_ARGUMENT, "null exchange");
    AMRB_CUDA(cudaSetDevice(ex->pool->device));
    unsigned long long* slot = with_dt ? reinterpret_cast<unsigned long long*>(amrb_pool_dtmin_slot(ex->pool, k)) : nullptr;
    if (with_dt && !slot) return fail(AMRB_ERR_STATE, "no dt-min slot: open a batch first");
    return push(ex, slot);
}
amrb_status amrb_exchange_wait(amrb_exchange* ex, int with_dt, size_t k)
{
    if (!ex) return fail(AMRB_ERR_ARGUMENT, "null exchange");
    AMRB_CUDA(cudaSetDevice(ex->pool->device));
    unsigned long long* slot = with_dt ? reinterpret_cast<unsigned long long*>(amrb_pool_dtmin_slot(ex->pool, k)) : nullptr;
    if (with_dt && !slot) return fail(AMRB_ERR_STATE, "no dt-min slot: open a batch first");
    return wait_unpack_fused(ex, slot); // the kernel of the batch loop's plain schedule
}

// per-phase device times [ms] of the last batch run with the plain schedule after amrb_exchange_set_timing(ex, 1):
// out[0..3] = push, wait (skew + flags), unpack, fused step, summed over the batch's steps
amrb_status amrb_exchange_set_timing(amrb_exchange* ex, int on)
{
    if (!ex) return fail(AMRB_ERR_ARGUMENT, "null exchange");
    ex->timing = on != 0;
    return AMRB_OK;
}
amrb_status amrb_exchange_get_timing(amrb_exchange* ex, double* out4)
{
    if (!ex || !out4) return fail(AMRB_ERR_ARGUMENT, "null argument");
    for (int i = 0; i < 4; ++i) out4[i] = 0.0;
    if (!ex->timing || ex->tsteps == 0) return fail(AMRB_ERR_STATE, "no timed batch");
    AMRB_CUDA(cudaStreamSynchronize(ex->pool->stream));
    for (size_t k = 0; k < ex->tsteps; ++k)
        for (int i = 0; i < 4; ++i)
        {
            float ms = 0.f;
            AMRB_CUDA(cudaEventElapsedTime(&ms, ex->tev[4 * k + i], ex->tev[4 * k + i + 1]));
            out4[i] += ms;
        }
    return AMRB_OK;
}

// 1 when a wait kernel gave up on a peer since the last call (the state is then undefined)
int amrb_exchange_timed_out(amrb_exchange* ex)
{
    if (!ex || !ex->h_timeout) return 0;
    const int v    = *ex->h_timeout;
    *ex->h_timeout = 0;
    return v;
}

uint64_t amrb_exchange_launch_count(const amrb_exchange* ex) { return ex ? ex->launches : 0; }

} // extern "C"
