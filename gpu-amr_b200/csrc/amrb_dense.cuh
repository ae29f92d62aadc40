// Kernels of interior-only pools that have no padded counterpart: the padded image of a patch (for
// host downloads and VTK output) and the refinement criterion over that image.
#pragma once
#include "amrb_kernels.cuh"

namespace amrb
{

// padded image of patches [first, first + n) of one field: interior cells from the dense pool, face
// ghosts (all H layers) gathered from the neighbor interiors exactly like halo_kernel would have
// written them (halo_source: same / coarser injection / finer restriction), zeros in the edge / corner
// ghosts nothing ever writes and in the ghosts of slots without tables (ghost slots).
template <int R, int S, int H>
__global__ void __launch_bounds__(256)
dense_export_kernel(const double* __restrict__ field, const int32_t* __restrict__ nbr,
                    const uint8_t* __restrict__ meta, int first, int n, int n_tabled,
                    double* __restrict__ staging)
{
    using GP    = Geo<R, S, H, H>;
    using GD    = Geo<R, S, H, 0>;
    const int j = blockIdx.x;
    if (j >= n) return;
    const int p = first + j;
    for (int i = threadIdx.x; i < GP::FLAT; i += blockDim.x)
    {
        int r = i, idx[R], outside = 0, od = 0;
#pragma unroll
        for (int k = R - 1; k >= 0; --k)
        {
            idx[k] = r % GP::P;
            r /= GP::P;
            if (idx[k] < H || idx[k] >= H + S)
            {
                ++outside;
                od = 2 * k + (idx[k] >= H + S ? 1 : 0);
            }
        }
        double v = 0.0;
        if (outside == 0)
        {
            int off = 0;
#pragma unroll
            for (int k = 0; k < R; ++k) off += (idx[k] - H) * GD::pitch(k);
            v = field[(size_t)p * GD::FLAT + off];
        }
        else if (outside == 1 && p < n_tabled)
        {
            const int m = meta[(size_t)p * GD::NDIR + od];
            if ((m & 3) != 0)
                v = halo_source<R, S, H, 0>(field, nbr + ((size_t)p * GD::NDIR + od) * GD::KF, m, od, idx);
        }
        staging[(size_t)j * GP::FLAT + i] = v;
    }
}

// refinement decisions of the benchmark criterion (max over ALL flat cells of the padded patch,
// src/cuda/fvm_refinement_criterion.cu:27-67) for an interior-only pool: the padded image is never
// built, its maximum is max(interior cells, gathered face ghosts, 0 for the edge / corner ghosts).
template <int R, int S, int H>
__global__ void __launch_bounds__(128)
dense_flags_kernel(const double* __restrict__ field, const int32_t* __restrict__ nbr,
                   const uint8_t* __restrict__ meta, const int32_t* __restrict__ level, int n_patches,
                   double refine_thr, double coarsen_thr, int min_level, int max_level,
                   int8_t* __restrict__ flags)
{
    using GD = Geo<R, S, H, 0>;
    __shared__ double red[4];
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    double m = 0.0;
    for (int i = threadIdx.x; i < GD::FLAT; i += 128)
    {
        const double v = field[(size_t)p * GD::FLAT + i];
        m              = v > m ? v : m;
    }
    constexpr int PER_DIR = H * GD::FACE;
    for (int it = threadIdx.x; it < GD::NDIR * PER_DIR; it += 128)
    {
        const int d = it / PER_DIR, r = it % PER_DIR;
        const int mm = meta[(size_t)p * GD::NDIR + d];
        if ((mm & 3) == 0) continue;
        int idx[R];
        slab_index<R, S, H>(d, r / GD::FACE, r % GD::FACE, idx);
        const double v = halo_source<R, S, H, 0>(field, nbr + ((size_t)p * GD::NDIR + d) * GD::KF, mm, d, idx);
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

} // namespace amrb
