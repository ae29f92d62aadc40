// amr::containers::static_shape — compile-time tensor extents.
// API mirror of the reference's include/containers/static_shape.hpp (rank(), sizes(), elements());
// written for the B200 hot path, where shapes only parameterise patch layouts.
#ifndef AMRB_CONTAINERS_STATIC_SHAPE_HPP
#define AMRB_CONTAINERS_STATIC_SHAPE_HPP
#include <array>
#include <cstddef>

namespace amr::containers
{
template <auto... Ns>
struct static_shape
{
    static_assert(sizeof...(Ns) > 0, "a shape needs at least one extent");
    static_assert(((Ns > 0) && ...), "extents must be positive");
    using size_type = std::size_t;
    using rank_t    = std::size_t;

    [[nodiscard]] static constexpr auto rank() noexcept -> rank_t { return sizeof...(Ns); }
    [[nodiscard]] static constexpr auto sizes() noexcept -> std::array<size_type, sizeof...(Ns)>
    {
        return { static_cast<size_type>(Ns)... };
    }
    [[nodiscard]] static constexpr auto size(rank_t i) noexcept -> size_type { return sizes()[i]; }
    [[nodiscard]] static constexpr auto elements() noexcept -> size_type
    {
        return (static_cast<size_type>(Ns) * ...);
    }
};
} // namespace amr::containers
#endif
