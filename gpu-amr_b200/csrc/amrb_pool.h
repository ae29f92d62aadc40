// Internal: the device patch pool as the translation units of the library see it.
#pragma once
#include "../../include/gpuamr_b200.h"
#include "amrb_ops.cuh"

#include <string>

struct amrb_pool
{
    amrb_layout  lay{};
    const amrb::Ops* ops = nullptr;
    int          device   = 0;
    cudaStream_t stream   = nullptr;
    bool         own_stream = false, own_mem = false;
    size_t       capacity = 0, n_owned = 0, n_total = 0, data = 0;
    size_t       flat  = 0; // doubles per field-patch as STORED (padded: prod(size+2h); interior-only: prod(size))
    size_t       pflat = 0; // doubles per padded field-patch (the host exchange format)
    bool         dense = false;
    amrb::FieldPtrs cur{}, nxt{};
    int32_t*     d_nbr   = nullptr;
    uint8_t*     d_meta  = nullptr;
    int32_t*     d_level = nullptr;
    size_t       table_cap = 0;
    double       lengths[3] = { 1.0, 1.0, 1.0 }, gamma = 1.4, cfl = 0.3;
    double       dx[amrb::kMaxLevel + 1][3]{};
    // batch scalars
    unsigned long long* d_dtmin = nullptr;
    unsigned int*       d_queue = nullptr; // dynamic task counter of the 3D marching kernel
    uint64_t*           d_ids = nullptr;   // leaf ids (device-side table build)
    size_t              ids_cap = 0;
    double*      d_remaining = nullptr;
    double*      d_dts       = nullptr;
    double*      h_dts       = nullptr; // pinned
    size_t       scal_cap = 0, batch_steps = 0, batch_k = 0;
    bool         batch_open = false, batch_pending = false, carry_valid = false;
    bool         pending_in_graph = false; // the batch was enqueued under stream capture: no event
    bool         lazy_halos = false;       // materialise face halos only when something observes them
    bool         halos_stale = false;
    bool         step_touched = false;
    cudaEvent_t  batch_done = nullptr;
    // staging
    double*      d_stage = nullptr;
    size_t       stage_cap = 0;
    int8_t*      d_flags = nullptr;
    size_t       flags_cap = 0;
    void*        d_plan = nullptr;
    size_t       plan_cap = 0;
    uint64_t     launches = 0;
    int          mode     = 0;
    int          variant  = 0; // kernel variant for A/B runs (amrb_pool_set_variant; default AMRB_VARIANT)
};

namespace amrb
{
// status + message of the calling thread (amrb_last_error)
amrb_status fail(amrb_status code, const std::string& what);
// scratch of the device reconstruct (amrb_regrid.cu), released with the pool
void regrid_release(amrb_pool* p);
} // namespace amrb

#define AMRB_CUDA(expr)                                                                          \
    do                                                                                           \
    {                                                                                            \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return amrb::fail(AMRB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define AMRB_TRY(expr)                                                                           \
    do                                                                                           \
    {                                                                                            \
        amrb_status s_ = (expr);                                                                 \
        if (s_ != AMRB_OK) return s_;                                                            \
    } while (0)
