// amr::ndt::utils::patches — host-side helpers over padded patch layouts
// (include/ndtree/patch_utils.hpp:35-58 of the reference).
#ifndef AMRB_NDTREE_PATCH_UTILS_HPP
#define AMRB_NDTREE_PATCH_UTILS_HPP
#include "patch_layout.hpp"

namespace amr::ndt::utils::patches
{
// true when the padded linear index lies in the ghost frame of any dim
template <typename Layout>
[[nodiscard]] constexpr auto is_halo_cell(typename Layout::index_t linear_index) noexcept -> bool
{
    using padded_t       = typename Layout::padded_layout_t;
    constexpr auto sizes = padded_t::sizes();
    constexpr auto h     = Layout::halo_width();
    const auto     m     = padded_t::multi_index(linear_index);
    for (std::size_t k = 0; k != padded_t::rank(); ++k)
        if (m[k] < h || m[k] >= sizes[k] - h) return true;
    return false;
}
} // namespace amr::ndt::utils::patches
#endif
