// Micro-benchmark: throughput of cp.async.bulk (TMA 1-D bulk copy) global -> shared on sm_100a as a
// function of copy size and source/destination alignment.  Each warp keeps NS copies in flight into
// a private shared-memory ring (the access pattern of the fused step kernels).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_bench tools/tma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// each warp: `copies` copies of `bytes` bytes, source stride `stride` bytes starting at byte offset `off`
template <int NS, int WPC>
__global__ void k(const char* src, size_t span, int bytes, int stride, int off, int copies, int slot, double* sink, int multi)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[WPC * NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* ring = smem + (size_t)warp * NS * slot;
    uint64_t*      bar  = bars + warp * NS;
    if (lane == 0)
        for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
    __syncwarp();
    const size_t gw = (size_t)blockIdx.x * WPC + warp, nw = (size_t)gridDim.x * WPC;
    // stride > 0: the whole grid sweeps one contiguous window (copy i of warp w at (w + i*nw)*stride)
    // stride < 0: private streams -- warp w walks its own region of -stride bytes sequentially, then
    //             jumps to region w + nw (the per-warp row streams of the fused step kernels)
    auto addr = [&](int i) {
        if (stride > 0) return src + (((gw + (size_t)i * nw) * (size_t)stride) % span) + off;
        const size_t region = (size_t)(-stride), per = region / bytes;
        const size_t r = gw + ((size_t)i / per) * nw;
        return src + ((r * region + ((size_t)i % per) * bytes) % span) + off;
    };
    for (int s = 0; s < NS && s < copies; ++s)
        if (lane == (multi ? s : 0))
        {
            mbar_expect_tx(&bar[s], bytes);
            bulk_g2s(ring + s * slot, addr(s), bytes, &bar[s]);
        }
    double acc = 0;
    int    st = 0, ph = 0;
    for (int i = 0; i < copies; ++i)
    {
        mbar_wait(&bar[st], ph);
        acc += reinterpret_cast<const double*>(ring + st * slot)[lane];
        __syncwarp();
        if (lane == (multi ? st : 0) && i + NS < copies)
        {
            mbar_expect_tx(&bar[st], bytes);
            bulk_g2s(ring + st * slot, addr(i + NS), bytes, &bar[st]);
        }
        if (++st == NS)
        {
            st = 0;
            ph ^= 1;
        }
    }
    if (acc == 1.2345) sink[0] = acc;
}

template <int NS, int WPC>
void run(const char* d, size_t span, int bytes, int stride, int off, int ctas_per_sm, double* sink, int multi = 0)
{
    const int slot   = (bytes + 127) / 128 * 128;
    const int smem   = NS * WPC * slot;
    const int copies = (int)((size_t)(1u << 30) / ((size_t)148 * ctas_per_sm * WPC * bytes)) + 1;
    cudaFuncSetAttribute(k<NS, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<NS, WPC><<<148 * ctas_per_sm, WPC * 32, smem>>>(d, span, bytes, stride, off, 8, slot, sink, multi);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NS, WPC><<<148 * ctas_per_sm, WPC * 32, smem>>>(d, span, bytes, stride, off, copies, slot, sink, multi);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)148 * ctas_per_sm * WPC * copies * bytes;
    printf("%sNS %d warps/SM %2d bytes %6d stride %6d off %3d : %8.1f GB/s  (%s)\n", multi ? "multi-lane " : "", NS,
           ctas_per_sm * WPC, bytes, stride, off, total / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t span = (size_t)3 << 30;
    char*        d;
    cudaMalloc(&d, span + (1 << 20));
    cudaMemset(d, 1, span + (1 << 20));
    double* sink;
    cudaMalloc(&sink, 8);
    // aligned vs unaligned sources, sizes of the fused step kernels' chunks
    for (int off : { 0, 16, 32, 64 })
    {
        run<3, 4>(d, span, 1024, 1024, off, 2, sink);
        run<3, 4>(d, span, 1056, 1056, off, 2, sink);
        run<3, 4>(d, span, 4224, 4224, off, 2, sink);
        run<2, 4>(d, span, 9504, 9504, off, 2, sink);
    }
    printf("private streams (region 38016 B = 18 rows x 4 fields x 528 B)\n");
    run<3, 4>(d, span, 1056, -38016, 0, 2, sink);
    run<3, 4>(d, span, 1056, -38016, 0, 4, sink);
    run<3, 4>(d, span, 3168, -38016, 0, 2, sink);
    run<3, 4>(d, span, 4224, -38016, 0, 2, sink);
    run<2, 4>(d, span, 9504, -38016, 0, 2, sink);
    run<2, 4>(d, span, 12672, -38016, 0, 2, sink);
    run<3, 4>(d, span, 1056, -139392, 0, 2, sink);
    run<3, 4>(d, span, 4224, -139392, 0, 2, sink);
    printf("outstanding copies per SM (private streams, 1056 B and 2112 B copies)\n");
    run<2, 4>(d, span, 1056, -38016, 0, 2, sink);
    run<6, 4>(d, span, 1056, -38016, 0, 2, sink);
    run<12, 4>(d, span, 1056, -38016, 0, 2, sink);
    run<3, 4>(d, span, 1056, -38016, 0, 8, sink);
    run<6, 4>(d, span, 1056, -38016, 0, 4, sink);
    run<12, 4>(d, span, 1056, -38016, 0, 4, sink);
    run<3, 4>(d, span, 2112, -38016, 0, 2, sink);
    run<6, 4>(d, span, 2112, -38016, 0, 2, sink);
    run<3, 4>(d, span, 2112, -38016, 0, 4, sink);
    run<4, 4>(d, span, 1056, -38016, 0, 2, sink, 1);
    run<8, 4>(d, span, 1056, -38016, 0, 2, sink, 1);
    run<12, 4>(d, span, 1056, -38016, 0, 2, sink, 1);
    run<4, 4>(d, span, 1056, -38016, 0, 4, sink, 1);
    printf("3D plane streams (region 8000 B = one 10^3 field-patch)\n");
    run<3, 4>(d, span, 1600, -8000, 0, 2, sink);
    run<4, 4>(d, span, 800, -8000, 0, 2, sink);
    run<3, 4>(d, span, 800, -8000, 0, 3, sink);
    run<2, 4>(d, span, 4000, -8000, 0, 2, sink);
    run<2, 4>(d, span, 8000, -8000, 0, 2, sink);
    run<3, 4>(d, span, 1600, -8000, 0, 2, sink, 1);
    run<4, 4>(d, span, 1152, -46656, 0, 2, sink);
    printf("grid sweep\n");
    run<3, 4>(d, span, 1056, 1056, 0, 4, sink);
    run<6, 4>(d, span, 1056, 1056, 0, 2, sink);
    run<6, 4>(d, span, 1024, 1024, 0, 2, sink);
    run<3, 4>(d, span, 4096, 4096, 0, 2, sink);
    run<3, 4>(d, span, 8192, 8192, 0, 2, sink);
    run<3, 4>(d, span, 8192, 8192, 16, 2, sink);
    run<3, 4>(d, span, 16384, 16384, 0, 2, sink);
    run<3, 1>(d, span, 16384, 16384, 0, 2, sink);
    run<3, 1>(d, span, 32768, 32768, 0, 2, sink);
    run<2, 1>(d, span, 65536, 65536, 0, 1, sink);
    return 0;
}
