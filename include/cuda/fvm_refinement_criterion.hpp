// amr::cuda::compute_scalar_patch_amr_decisions_from_device with the reference's signature
// (include/cuda/fvm_refinement_criterion.hpp:8-27): per-patch max over all flat cells -> decision.
#ifndef AMR_INCLUDED_CUDA_FVM_REFINEMENT_CRITERION
#define AMR_INCLUDED_CUDA_FVM_REFINEMENT_CRITERION
#include "amrb_check.hpp"
#include <cstddef>
#include <cstdint>
#include <stdexcept>

namespace amr::cuda
{
struct scalar_patch_amr_launch_config
{
    std::size_t num_patches;
    std::size_t cells_per_patch;
    double      refine_threshold;
    double      coarsen_threshold;
    int         min_level;
    int         max_level;
};

inline auto compute_scalar_patch_amr_decisions_from_device(
    const double* device_patch_data, const int* device_patch_levels, std::size_t level_count,
    const scalar_patch_amr_launch_config& config, std::int8_t* device_decisions,
    std::size_t decision_count) -> void
{
    if (config.num_patches == 0) return;
    if (decision_count < config.num_patches || level_count < config.num_patches)
        throw std::runtime_error("FVM CUDA AMR inputs are smaller than the patch count");
    detail::check(amrb_patch_max_flags_device(device_patch_data, device_patch_levels, config.num_patches,
                                              config.cells_per_patch, config.refine_threshold,
                                              config.coarsen_threshold, config.min_level,
                                              config.max_level, device_decisions, nullptr),
                  "scalar_patch_amr_kernel launch");
}
} // namespace amr::cuda
#endif
