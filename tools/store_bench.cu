// Micro-benchmark: global store throughput on sm_100a for the row-store shapes of the fused step
// kernels.  Every warp writes rows of 64 doubles (512 B) into padded rows of 66 doubles (528 B pitch,
// interior starts at +8 B), 4 fields (separate arrays), private row streams.
//   mode 0: lane owns cells (2l, 2l+1): two STG.64 with a 16-byte lane stride (march kernel)
//   mode 1: lane owns cells (l, l+32): two dense STG.64 (256 contiguous bytes per instruction)
//   mode 2: one STG.128 per lane at a 16-byte aligned address (shifted one cell left; same volume)
//   mode 4: mode 2 plus the two ghost columns: the whole 528-byte padded row, STG.128
//   mode 3: dense, unpadded rows (512 B pitch), STG.128: the upper bound
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_bench tools/store_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* f0, double* f1, double* f2, double* f3, int rows_per_warp, size_t total_rows)
{
    const int    lane = threadIdx.x & 31;
    const size_t gw   = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nw   = (size_t)gridDim.x * (blockDim.x >> 5);
    double*      f[4] = { f0, f1, f2, f3 };
    const double v    = 1.0 + lane;
    // warp w owns bands of 16 consecutive rows: band index w, w + nw, ...
    for (int r = 0; r < rows_per_warp; ++r)
    {
        const size_t band = gw + (size_t)(r / 16) * nw;
        const size_t row  = (band * 16 + (r % 16)) % total_rows;
        const size_t pitch = (MODE == 3) ? 64 : 66;
        const size_t o     = row * pitch;
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
            if (MODE == 0)
            {
                f[q][o + 1 + 2 * lane] = v;
                f[q][o + 2 + 2 * lane] = v + r;
            }
            else if (MODE == 1)
            {
                f[q][o + 1 + lane]      = v;
                f[q][o + 1 + lane + 32] = v + r;
            }
            else if (MODE == 2)
                *reinterpret_cast<double2*>(f[q] + o + 2 * lane) = make_double2(v, v + r);
            else if (MODE == 4)
            {
                // whole padded row (ghost columns included): no sector is left partially written
                *reinterpret_cast<double2*>(f[q] + o + 2 * lane) = make_double2(v, v + r);
                if (lane == 31) *reinterpret_cast<double2*>(f[q] + o + 64) = make_double2(v, v + r);
            }
            else
                *reinterpret_cast<double2*>(f[q] + o + 2 * lane) = make_double2(v, v + r);
        }
    }
}

template <int MODE>
void run(double** f, size_t total_rows, int ctas_per_sm, int wpc)
{
    const int    grid = 148 * ctas_per_sm;
    const size_t nw   = (size_t)grid * wpc;
    const int    rows_per_warp = (int)(((size_t)1 << 21) / nw) / 16 * 16; // ~2M rows x 4 fields x 512 B = 4.3 GB
    k<MODE><<<grid, wpc * 32>>>(f[0], f[1], f[2], f[3], 32, total_rows);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, wpc * 32>>>(f[0], f[1], f[2], f[3], rows_per_warp, total_rows);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)nw * rows_per_warp * 4 * 512;
    printf("mode %d warps/SM %2d : %8.1f GB/s (%s)\n", MODE, ctas_per_sm * wpc, bytes / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t total_rows = (size_t)1 << 21;
    double*      f[4];
    for (int q = 0; q < 4; ++q)
    {
        cudaMalloc(&f[q], total_rows * 66 * 8 + 4096);
        cudaMemset(f[q], 0, total_rows * 66 * 8 + 4096);
    }
    for (int w : { 2, 4 })
    {
        run<0>(f, total_rows, w, 4);
        run<1>(f, total_rows, w, 4);
        run<2>(f, total_rows, w, 4);
        run<3>(f, total_rows, w, 4);
        run<4>(f, total_rows, w, 4);
    }
    return 0;
}
