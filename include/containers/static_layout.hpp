// amr::containers::static_layout — row-major (last dim fastest) strides over a static_shape.
// API mirror of include/containers/static_layout.hpp:28-37, 86-125 of the reference.
#ifndef AMRB_CONTAINERS_STATIC_LAYOUT_HPP
#define AMRB_CONTAINERS_STATIC_LAYOUT_HPP
#include "static_shape.hpp"

namespace amr::containers
{
template <typename Shape>
struct static_layout
{
    using shape_t     = Shape;
    using size_type   = typename Shape::size_type;
    using rank_t      = typename Shape::rank_t;
    using index_t     = size_type;
    using multi_idx_t = std::array<size_type, Shape::rank()>;

    [[nodiscard]] static constexpr auto rank() noexcept -> rank_t { return Shape::rank(); }
    [[nodiscard]] static constexpr auto sizes() noexcept { return Shape::sizes(); }
    [[nodiscard]] static constexpr auto size(rank_t i) noexcept -> size_type { return Shape::size(i); }
    [[nodiscard]] static constexpr auto flat_size() noexcept -> size_type { return Shape::elements(); }
    [[nodiscard]] static constexpr auto elements() noexcept -> size_type { return Shape::elements(); }
    [[nodiscard]] static constexpr auto strides() noexcept -> multi_idx_t
    {
        multi_idx_t s{};
        size_type   acc = 1;
        for (rank_t k = rank(); k-- > 0;)
        {
            s[k] = acc;
            acc *= Shape::size(k);
        }
        return s;
    }
    [[nodiscard]] static constexpr auto stride(rank_t i) noexcept -> size_type { return strides()[i]; }
    [[nodiscard]] static constexpr auto linear_index(multi_idx_t const& idx) noexcept -> index_t
    {
        constexpr auto s = strides();
        index_t        l = 0;
        for (rank_t k = 0; k != rank(); ++k) l += idx[k] * s[k];
        return l;
    }
    [[nodiscard]] static constexpr auto multi_index(index_t linear) noexcept -> multi_idx_t
    {
        constexpr auto s = strides();
        multi_idx_t    m{};
        for (rank_t k = 0; k != rank(); ++k)
        {
            m[k] = linear / s[k];
            linear %= s[k];
        }
        return m;
    }
};
} // namespace amr::containers
#endif
