// amr::ndt::solver::physics_system — geometry policy: Morton id -> patch origin / cell size.
// Same static interface and formulas as include/solver/physics_system.hpp:58-138 of the reference:
// physical axis i pairs with layout dim rank-1-i (x is the fastest layout dim).
#ifndef AMRB_SOLVER_PHYSICS_SYSTEM_HPP
#define AMRB_SOLVER_PHYSICS_SYSTEM_HPP
#include <array>
#include <cstddef>
#include <cstdint>

namespace amr::ndt::solver
{
template <typename Patch_Index, typename Patch_Layout, auto Domain_Sizes>
class physics_system
{
public:
    using patch_index_t   = Patch_Index;
    using patch_layout_t  = Patch_Layout;
    using physics_coord_t = double;
    using size_type       = typename patch_index_t::size_type;
    static constexpr size_type n_dimension = patch_index_t::rank();
    using physics_coord_arr_t              = std::array<physics_coord_t, n_dimension>;

    [[nodiscard]] static constexpr auto lengths() noexcept -> physics_coord_arr_t
    {
        physics_coord_arr_t r{};
        for (size_type i = 0; i != n_dimension; ++i) r[i] = static_cast<double>(Domain_Sizes[i]);
        return r;
    }
    // physical extent of the patch: L_i * 2^(Depth-level) / 2^Depth
    [[nodiscard]] static constexpr auto patch_sizes(patch_index_t const& id) noexcept -> physics_coord_arr_t
    {
        const auto          span = static_cast<double>(1u << (patch_index_t::max_depth() - id.level()));
        const auto          full = static_cast<double>(1u << patch_index_t::max_depth());
        physics_coord_arr_t r{};
        for (size_type i = 0; i != n_dimension; ++i) r[i] = static_cast<double>(Domain_Sizes[i]) * span / full;
        return r;
    }
    [[nodiscard]] static constexpr auto cell_sizes(patch_index_t const& id) noexcept -> physics_coord_arr_t
    {
        constexpr auto cells = patch_layout_t::data_layout_t::sizes();
        auto           r     = patch_sizes(id);
        for (size_type i = 0; i != n_dimension; ++i) r[i] /= static_cast<double>(cells[n_dimension - 1 - i]);
        return r;
    }
    // origin of the patch: L_i * anchor_i / 2^Depth
    [[nodiscard]] static constexpr auto patch_coord(patch_index_t const& id) noexcept -> physics_coord_arr_t
    {
        const auto          c    = id.coords();
        const auto          full = static_cast<double>(1u << patch_index_t::max_depth());
        physics_coord_arr_t r{};
        for (size_type i = 0; i != n_dimension; ++i)
            r[i] = static_cast<double>(Domain_Sizes[i]) * static_cast<double>(c[i]) / full;
        return r;
    }
    // lower corner of padded cell `linear_idx`
    [[nodiscard]] static constexpr auto cell_coord(patch_index_t const& id, std::size_t linear_idx) noexcept
        -> physics_coord_arr_t
    {
        using padded_t   = typename patch_layout_t::padded_layout_t;
        const auto m     = padded_t::multi_index(linear_idx);
        const auto dx    = cell_sizes(id);
        auto       r     = patch_coord(id);
        const auto h     = static_cast<double>(patch_layout_t::halo_width());
        for (size_type i = 0; i != n_dimension; ++i)
            r[i] += (static_cast<double>(m[n_dimension - 1 - i]) - h) * dx[i];
        return r;
    }
};
} // namespace amr::ndt::solver
#endif
