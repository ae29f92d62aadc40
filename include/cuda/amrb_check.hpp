// Re-throws C-ABI failures as std::runtime_error, the error behaviour of the reference's CUDA
// wrappers (throw_if_cuda_error, src/cuda/halo_exchange.cu:16-26).
#ifndef AMRB_CUDA_CHECK_HPP
#define AMRB_CUDA_CHECK_HPP
#include "gpuamr_b200.h"
#include <stdexcept>
#include <string>

namespace amr::cuda::detail
{
inline auto check(amrb_status st, const char* context) -> void
{
    if (st != AMRB_OK) throw std::runtime_error(std::string(context) + ": " + amrb_last_error());
}
} // namespace amr::cuda::detail
#endif
