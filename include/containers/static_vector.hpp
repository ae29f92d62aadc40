// amr::containers::static_vector — fixed-size value vector (aggregate, brace-initialisable).
// API mirror of include/containers/static_vector.hpp of the reference (operator[], size, data).
#ifndef AMRB_CONTAINERS_STATIC_VECTOR_HPP
#define AMRB_CONTAINERS_STATIC_VECTOR_HPP
#include <cstddef>

namespace amr::containers
{
template <typename T, std::size_t N>
struct static_vector
{
    using value_type = T;
    using size_type  = std::size_t;
    T m_data[N];

    [[nodiscard]] static constexpr auto size() noexcept -> size_type { return N; }
    [[nodiscard]] constexpr auto operator[](size_type i) noexcept -> T& { return m_data[i]; }
    [[nodiscard]] constexpr auto operator[](size_type i) const noexcept -> T const& { return m_data[i]; }
    [[nodiscard]] constexpr auto data() noexcept -> T* { return m_data; }
    [[nodiscard]] constexpr auto data() const noexcept -> T const* { return m_data; }
    [[nodiscard]] constexpr auto begin() noexcept -> T* { return m_data; }
    [[nodiscard]] constexpr auto end() noexcept -> T* { return m_data + N; }
    [[nodiscard]] constexpr auto begin() const noexcept -> T const* { return m_data; }
    [[nodiscard]] constexpr auto end() const noexcept -> T const* { return m_data + N; }
};
} // namespace amr::containers
#endif
