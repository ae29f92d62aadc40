// amr::ndt::intergrid_operator::linear_interpolator — the transfer operator between refinement levels:
// piecewise-constant injection coarse -> fine, arithmetic mean of the 2^rank children fine -> coarse
// (include/ndtree/intergrid_operator.hpp:18-106 of the reference).
//
// On the device path the operator is a SELECTOR: the arithmetic lives in the CUDA kernels (halo_source: coarser_t
// injection / finer_t mean in this summation order; plan_kernel: prolongation / restriction of whole patches), and
// ndtree static_asserts that its IntergridOperator is this type.  The three static members below are the host
// form of the same arithmetic, with the reference's signatures, so that host code written against the operator
// (custom steppers over get_patch / get_out_patch, tests) keeps compiling and computes what the kernels compute.
#ifndef AMRB_NDTREE_INTERGRID_OPERATOR_HPP
#define AMRB_NDTREE_INTERGRID_OPERATOR_HPP
#include "ndconcepts.hpp"
#include "patch_layout.hpp"

#include <array>
#include <concepts>
#include <type_traits>

namespace amr::ndt::intergrid_operator
{
template <typename Patch_Layout>
struct linear_interpolator
{
    using patch_layout_t = Patch_Layout;
    using index_t        = typename Patch_Layout::index_t;
    static constexpr bool device_native = true; // implemented by the kernels of libgpuamr_b200

    // fine cell <- the coarse cell covering it (child_offset: which of the 2^rank children; unused by injection)
    static constexpr auto interpolation(auto& to, index_t const& to_idx, [[maybe_unused]] index_t const& child_offset,
                                        auto const& from, index_t const from_idx) noexcept -> void
    {
        to[to_idx] = from[from_idx];
    }
    // all children of one coarse cell at once (patch prolongation during reconstruct_tree)
    template <std::integral auto N>
    static constexpr auto interpolation(auto& to, std::array<index_t, N> const& to_idxs, auto const& from,
                                        index_t const from_idx) noexcept -> void
    {
        for (index_t i{}; i != index_t{ N }; ++i) interpolation(to, to_idxs[i], i, from, from_idx);
    }
    // coarse cell <- mean of its N fine cells, summed in the order given (the kernels: last layout dim fastest)
    template <std::integral auto N>
    static constexpr auto restriction(auto& to, index_t const to_idx, auto const& from,
                                      std::array<index_t, N> const& from_idxs) noexcept -> void
    {
        using value_type = typename std::remove_cvref_t<decltype(from)>::value_type;
        value_type sum{};
        for (auto const i : from_idxs) sum += from[i];
        to[to_idx] = sum / static_cast<value_type>(N);
    }
};
} // namespace amr::ndt::intergrid_operator
#endif
