// Kernel dispatch table shared by the translation units of the library: one Ops per instantiated
// (rank, patch size, halo, equation, storage) combination.
#pragma once
#include "amrb_kernels.cuh"

namespace amrb
{

struct Ops
{
    int    rank, size, halo, eq;
    int    storage;    // 0 = padded (reference device layout), 1 = interior-only
    int    bands;      // CTAs per patch in the thread-per-cell step
    size_t step_smem;  // dynamic shared memory of the thread-per-cell step
    cudaError_t (*prepare)();
    void (*halo_fill)(cudaStream_t, const FieldPtrs&, const int32_t*, const uint8_t*, int);
    void (*step)(cudaStream_t, const StepArgs&, int n_items);
    void (*step_v1)(cudaStream_t, const StepArgs&, int n_items); // thread-per-cell variant (A/B)
    void (*compute_dt)(cudaStream_t, const StepArgs&, unsigned long long*);
    void (*plan)(cudaStream_t, const FieldPtrs&, const FieldPtrs&, const int8_t*, const int32_t*,
                 const int8_t*, int);
    void (*flags)(cudaStream_t, const double*, const int32_t*, int, double, double, int, int,
                  int8_t*);
    // padded staging <-> pool: padded pools copy the interior, interior-only pools are the dense side
    void (*interior)(cudaStream_t, double* padded, double* dense, int n, int to_padded);
    void (*faces)(cudaStream_t, const FieldPtrs&, const int32_t*, int, double*, int);
    // interior-only pools: padded image (interior + gathered face ghosts, zeros elsewhere) of patches
    // [first, first + n) of one field into a staging buffer; criterion over that same image
    void (*export_padded)(cudaStream_t, const double* field, const int32_t* nbr, const uint8_t* meta,
                          int first, int n, int n_tabled, double* staging);
    void (*flags_dense)(cudaStream_t, const double* field, const int32_t* nbr, const uint8_t* meta,
                        const int32_t* level, int n, double, double, int, int, int8_t*);
};

// per-device caches (function attributes and the SM count belong to a device; amrb_api.cu)
int  device_sm_count();
struct DevicePrepared
{
    bool done[64] = {};
    bool ensure(const void* kernel, int smem_bytes);
};

// interior-only 3D instantiations (amrb_dense.cu)
const Ops* dense_ops(int* count);

} // namespace amrb
