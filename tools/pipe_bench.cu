// Micro-benchmark of the sm_100a issue rates that bound the Euler kernels: DFMA, DADD, fmax(double),
// MUFU.RCP64H / RSQ64H (rcp/rsqrt.approx.ftz.f64), F2F f64<->f32, FP32 MUFU.  Prints warp-instr / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double seed)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            if (OP == 0) a[i] = fma(a[i], 1.0000001, 1e-9);
            if (OP == 1) a[i] = a[i] + 1e-9;
            if (OP == 2) a[i] = fmax(a[i], a[(i + 1) & 7] * 0.5);
            if (OP == 3) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(a[i]) : "d"(a[i]));
            if (OP == 4) asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(a[i]) : "d"(a[i]));
            if (OP == 5) a[i] = (double)(float)a[i] + 1e-9;
            if (OP == 6) { float f = (float)a[i]; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(f)); a[i] = f; }
            if (OP == 7) a[i] = 1.0 / a[i];
            if (OP == 8) a[i] = sqrt(a[i]);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, double ops_per_iter)
{
    double* d;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
    const int iters = 4096;
    k<OP><<<148 * 8, 256>>>(d, 16, 1.5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<148 * 8, 256>>>(d, iters, 1.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_ops = 148.0 * 8 * 8 * iters * 8 * ops_per_iter; // CTAs * warps * iters * chains
    const double cycles   = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %7.3f warp-ops/clk/SM (at %d MHz nominal)\n", name, ms, warp_ops / cycles / 148.0, clk / 1000);
    cudaFree(d);
}
int main()
{
    run<0>("DFMA", 1);
    run<1>("DADD", 1);
    run<2>("fmax(double)+DMUL", 1);
    run<3>("MUFU.RCP64H", 1);
    run<4>("MUFU.RSQ64H", 1);
    run<5>("F2F f64->f32->f64 + DADD", 1);
    run<6>("F2F + MUFU.RCP f32 + F2F", 1);
    run<7>("IEEE 1.0/x (fp64)", 1);
    run<8>("IEEE sqrt (fp64)", 1);
    return 0;
}
