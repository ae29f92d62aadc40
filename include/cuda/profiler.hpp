// amr::cuda profiling hooks with the reference's names (include/cuda/profiler.hpp).
#ifndef AMR_INCLUDED_CUDA_PROFILER
#define AMR_INCLUDED_CUDA_PROFILER
#include "gpuamr_b200.h"

namespace amr::cuda
{
inline auto profile_capture_start() -> void { amrb_profile_capture_start(); }
inline auto profile_capture_stop() -> void { amrb_profile_capture_stop(); }
inline auto profile_range_push(const char* label) -> void { amrb_profile_range_push(label); }
inline auto profile_range_pop() noexcept -> void { amrb_profile_range_pop(); }

class scoped_profile_range
{
public:
    explicit scoped_profile_range(const char* label) { profile_range_push(label); }
    ~scoped_profile_range() { profile_range_pop(); }
    scoped_profile_range(scoped_profile_range const&)                    = delete;
    auto operator=(scoped_profile_range const&) -> scoped_profile_range& = delete;
};
} // namespace amr::cuda
#endif
