// amr::ndt::neighbors — direction system and neighbor variants of the patch tree.
//
// Public vocabulary of the reference's include/ndtree/neighbor.hpp:22-255 and
// docs/DIRECTION_SYSTEM_EXPLANATION.md:19-32 (index d: dimension = d / 2 == LAYOUT dim, even = negative,
// odd = positive, opposite(d) = d ^ 1; a face neighbor is none | same{id} | finer{ids[2^(rank-1)]} |
// coarser{id, contact_quadrant[rank]}).  The reference maintains these variants incrementally inside the
// tree (neighbor_utils, :291-572); here they are VIEWS decoded on demand from the tables the device
// kernels use (amrb_tree_tables: rel / nbr / quad per (patch, direction)), so this header holds no
// topology logic.
#ifndef AMRB_NDTREE_NEIGHBOR_HPP
#define AMRB_NDTREE_NEIGHBOR_HPP
#include <array>
#include <compare>
#include <cstddef>
#include <string_view>
#include <type_traits>
#include <variant>

namespace amr::ndt::neighbors
{
// Face neighbor of a leaf across one direction.  Fanout_1D children per dim (2), Rank dims,
// Identifier = patch index (get_neighbor_at) or linear index (neighbor_linear_index).
template <auto Fanout_1D, auto Rank, typename Identifier>
struct neighbor_variant
{
    using identifier_t = Identifier;
    using index_t      = decltype(Fanout_1D);
    static constexpr auto s_rank      = Rank;
    static constexpr auto s_1d_fanout = Fanout_1D;

    struct none
    {
        static constexpr std::string_view s_repr = "None";
    };
    struct same
    {
        static constexpr std::string_view s_repr = "Same";
        identifier_t id;
    };
    struct coarser
    {
        static constexpr std::string_view s_repr = "Coarser";
        using container_t = std::array<index_t, static_cast<std::size_t>(Rank)>;
        identifier_t id;
        container_t  contact_quadrant; // which half of the coarse face this patch touches, per layout dim
    };
    struct finer
    {
        static constexpr std::string_view s_repr = "Finer";
        [[nodiscard]] static constexpr auto num_neighbors() noexcept -> std::size_t
        {
            std::size_t n = 1;
            for (std::size_t k = 1; k < static_cast<std::size_t>(Rank); ++k) n *= static_cast<std::size_t>(Fanout_1D);
            return n;
        }
        using container_t = std::array<identifier_t, num_neighbors()>;
        container_t ids; // non-normal layout dims ascending, lowest dim = fastest bit
    };

    using type = std::variant<none, same, finer, coarser>;
    type data  = none{};

    [[nodiscard]] constexpr auto repr() const noexcept -> std::string_view
    {
        return std::visit([](auto const& v) { return std::remove_cvref_t<decltype(v)>::s_repr; }, data);
    }
};

// One of the 2 * Dim face directions.  Iterate with  for (auto d = first(); d != sentinel(); d.advance()).
template <std::integral auto Dim>
struct direction
{
    using index_t        = decltype(Dim);
    using size_type      = index_t;
    using signed_index_t = std::make_signed_t<index_t>;
    using vector_t       = std::array<signed_index_t, static_cast<std::size_t>(Dim)>;

    [[nodiscard]] static constexpr auto rank() noexcept -> size_type { return Dim; }
    [[nodiscard]] static constexpr auto elements() noexcept -> size_type { return size_type{ 2 } * Dim; }
    [[nodiscard]] static constexpr auto first() noexcept -> direction { return direction{ index_t{ 0 } }; }
    [[nodiscard]] static constexpr auto sentinel() noexcept -> direction { return direction{ elements() }; }
    [[nodiscard]] static constexpr auto from_index(index_t i) noexcept -> direction { return direction{ i }; }

    [[nodiscard]] static constexpr auto dimension_offset(direction const& d) noexcept -> index_t { return d.idx_ % index_t{ 2 }; }
    [[nodiscard]] static constexpr auto is_negative(direction const& d) noexcept -> bool { return dimension_offset(d) == 0; }
    [[nodiscard]] static constexpr auto is_positive(direction const& d) noexcept -> bool { return !is_negative(d); }
    [[nodiscard]] static constexpr auto opposite(direction const& d) noexcept -> direction
    {
        return direction{ static_cast<index_t>(d.idx_ ^ index_t{ 1 }) };
    }
    [[nodiscard]] static constexpr auto advance(direction d) noexcept -> direction
    {
        d.advance();
        return d;
    }
    [[nodiscard]] static constexpr auto unit_vector(direction const& d) noexcept -> vector_t
    {
        vector_t v{};
        v[static_cast<std::size_t>(d.dimension())] = d.is_negative() ? signed_index_t{ -1 } : signed_index_t{ 1 };
        return v;
    }

    constexpr auto advance() noexcept -> void { ++idx_; }
    [[nodiscard]] constexpr auto index() const noexcept -> index_t { return idx_; }
    [[nodiscard]] constexpr auto dimension() const noexcept -> index_t { return idx_ / index_t{ 2 }; }
    [[nodiscard]] constexpr auto is_negative() const noexcept -> bool { return is_negative(*this); }
    [[nodiscard]] constexpr auto is_positive() const noexcept -> bool { return is_positive(*this); }
    [[nodiscard]] constexpr auto repr() const noexcept -> std::string_view
    {
        constexpr std::string_view names[] = { "0-", "0+", "1-", "1+", "2-", "2+" };
        return static_cast<bool>(*this) ? names[static_cast<std::size_t>(idx_)] : std::string_view{ "??" };
    }
    [[nodiscard]] explicit constexpr operator bool() const noexcept { return idx_ >= index_t{} && idx_ < elements(); }
    [[nodiscard]] constexpr auto operator<=>(direction const&) const noexcept = default;

    index_t idx_{};

private:
    explicit constexpr direction(index_t i) noexcept : idx_{ i } {}
};
} // namespace amr::ndt::neighbors
#endif
