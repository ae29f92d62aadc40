// amr::ndt::utils::patches — host-side helpers over padded patch layouts
// (include/ndtree/patch_utils.hpp:35-58, 203-234 of the reference).  The halo operators of that file (same_t /
// finer_t / coarser_t, halo_apply) are the device kernels here (csrc/amrb_kernels.cuh: halo_source).
#ifndef AMRB_NDTREE_PATCH_UTILS_HPP
#define AMRB_NDTREE_PATCH_UTILS_HPP
#include "patch_layout.hpp"

#include <array>
#include <concepts>
#include <cstddef>

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

namespace detail
{
// linear padded indices of the N^rank cells of the hypercube whose lowest corner is `idx`, last layout dim
// fastest (include/ndtree/patch_utils.hpp:203-234 of the reference): the children of a coarse cell under
// refinement by N, in the order linear_interpolator::restriction sums them -- and the order the kernels'
// finer_t gather (`fine_mean5`, `halo_source`) reproduces on the device
template <typename Patch_Layout, std::integral auto N>
[[nodiscard]] constexpr auto hypercube_offset(typename Patch_Layout::index_t idx) noexcept
{
    using index_t        = typename Patch_Layout::index_t;
    using padded_t       = typename Patch_Layout::padded_layout_t;
    constexpr auto rank  = padded_t::rank();
    constexpr auto count = []
    {
        std::size_t k = 1;
        for (std::size_t i = 0; i != rank; ++i) k *= static_cast<std::size_t>(N);
        return k;
    }();
    constexpr auto              strides = padded_t::strides();
    std::array<index_t, count>  out{};
    for (std::size_t c = 0; c != count; ++c)
    {
        index_t     o = idx;
        std::size_t r = c;
        for (std::size_t k = rank; k-- > 0;) // digit of the last dim varies fastest
        {
            o += static_cast<index_t>(r % static_cast<std::size_t>(N)) * static_cast<index_t>(strides[k]);
            r /= static_cast<std::size_t>(N);
        }
        out[c] = o;
    }
    return out;
}
} // namespace detail
} // namespace amr::ndt::utils::patches
#endif
