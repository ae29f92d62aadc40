// amr::ndt::morton::morton_id<Depth, Rank> — patch index of the linear 2^d-tree.
// id = interleave(x, y[, z]) << 6 | level; x occupies bit 0 of every interleaved group; coordinates
// are anchor coordinates in units of the finest level.  Same encoding, ordering and member names as
// include/morton/morton_id.hpp:21-229 / 232-450 of the reference, implemented with plain bit loops
// (the host side only touches ids during topology changes).
#ifndef AMRB_MORTON_MORTON_ID_HPP
#define AMRB_MORTON_MORTON_ID_HPP
#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>

namespace amr::ndt::morton
{
template <unsigned Depth, unsigned Rank>
class morton_id
{
    static_assert(Rank == 2 || Rank == 3, "morton_id supports rank 2 and 3");
    static_assert(Depth >= 1 && Depth * Rank + 6 <= 64, "id does not fit 64 bits");

public:
    using id_t         = std::uint64_t;
    using size_type    = std::uint32_t;
    using level_t      = std::uint8_t;
    using coord_t      = std::uint32_t;
    using coord_array_t = std::array<coord_t, Rank>;
    using offset_t     = std::uint32_t;
    static constexpr unsigned s_level_bits = 6;

    constexpr morton_id() noexcept = default;
    constexpr explicit morton_id(id_t raw) noexcept : m_id(raw) {}
    constexpr morton_id(coord_array_t const& coords, level_t level) noexcept : m_id(encode(coords, level)) {}

    [[nodiscard]] static constexpr auto rank() noexcept -> size_type { return Rank; }
    [[nodiscard]] static constexpr auto max_depth() noexcept -> size_type { return Depth; }
    [[nodiscard]] static constexpr auto fanout() noexcept -> size_type { return 1u << Rank; }
    [[nodiscard]] static constexpr auto nd_fanout() noexcept -> size_type { return 1u << Rank; }
    [[nodiscard]] static constexpr auto root() noexcept -> morton_id { return morton_id{ id_t{ 0 } }; }

    [[nodiscard]] constexpr auto id() const noexcept -> id_t { return m_id; }
    [[nodiscard]] constexpr auto level() const noexcept -> level_t { return static_cast<level_t>(m_id & 63u); }

    [[nodiscard]] static constexpr auto encode(coord_array_t const& c, level_t level) noexcept -> id_t
    {
        id_t m = 0;
        for (unsigned b = 0; b != Depth + 1; ++b)
            for (unsigned a = 0; a != Rank; ++a) m |= static_cast<id_t>((c[a] >> b) & 1u) << (Rank * b + a);
        return (m << s_level_bits) | level;
    }
    [[nodiscard]] static constexpr auto decode(id_t raw) noexcept -> std::pair<coord_array_t, level_t>
    {
        coord_array_t c{};
        const id_t    m = raw >> s_level_bits;
        for (unsigned b = 0; b != Depth + 1; ++b)
            for (unsigned a = 0; a != Rank; ++a) c[a] |= static_cast<coord_t>((m >> (Rank * b + a)) & 1u) << b;
        return { c, static_cast<level_t>(raw & 63u) };
    }
    [[nodiscard]] constexpr auto coords() const noexcept -> coord_array_t { return decode(m_id).first; }

    // child `offset` (bit a of offset advances coordinate a by half the parent's extent)
    [[nodiscard]] static constexpr auto child_of(morton_id const& parent, offset_t offset) noexcept -> morton_id
    {
        auto [c, l]     = decode(parent.m_id);
        const coord_t h = coord_t{ 1 } << (Depth - l - 1);
        for (unsigned a = 0; a != Rank; ++a)
            if ((offset >> a) & 1u) c[a] += h;
        return morton_id{ c, static_cast<level_t>(l + 1) };
    }
    [[nodiscard]] static constexpr auto parent_of(morton_id const& child) noexcept -> morton_id
    {
        auto [c, l]     = decode(child.m_id);
        const coord_t e = coord_t{ 1 } << (Depth - l + 1);
        for (unsigned a = 0; a != Rank; ++a) c[a] &= ~(e - 1);
        return morton_id{ c, static_cast<level_t>(l - 1) };
    }

    [[nodiscard]] friend constexpr auto operator==(morton_id const& a, morton_id const& b) noexcept -> bool { return a.m_id == b.m_id; }
    [[nodiscard]] friend constexpr auto operator<(morton_id const& a, morton_id const& b) noexcept -> bool { return a.m_id < b.m_id; }
    [[nodiscard]] friend constexpr auto operator<=>(morton_id const& a, morton_id const& b) noexcept { return a.m_id <=> b.m_id; }

private:
    id_t m_id = 0;
};
} // namespace amr::ndt::morton

template <unsigned D, unsigned R>
struct std::hash<amr::ndt::morton::morton_id<D, R>>
{
    auto operator()(amr::ndt::morton::morton_id<D, R> const& v) const noexcept -> std::size_t
    {
        return static_cast<std::size_t>(v.id());
    }
};
#endif
