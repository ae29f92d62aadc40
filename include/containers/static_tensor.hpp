// amr::containers::static_tensor — flat storage addressed through a static_layout.
#ifndef AMRB_CONTAINERS_STATIC_TENSOR_HPP
#define AMRB_CONTAINERS_STATIC_TENSOR_HPP
#include "static_layout.hpp"

namespace amr::containers
{
template <typename T, typename Layout>
struct static_tensor
{
    using value_type  = T;
    using layout_t    = Layout;
    using size_type   = typename Layout::size_type;
    using multi_idx_t = typename Layout::multi_idx_t;
    std::array<T, Layout::flat_size()> m_data;

    [[nodiscard]] static constexpr auto flat_size() noexcept -> size_type { return Layout::flat_size(); }
    [[nodiscard]] constexpr auto operator[](size_type i) noexcept -> T& { return m_data[i]; }
    [[nodiscard]] constexpr auto operator[](size_type i) const noexcept -> T const& { return m_data[i]; }
    [[nodiscard]] constexpr auto operator[](multi_idx_t const& i) noexcept -> T& { return m_data[Layout::linear_index(i)]; }
    [[nodiscard]] constexpr auto operator[](multi_idx_t const& i) const noexcept -> T const& { return m_data[Layout::linear_index(i)]; }
    [[nodiscard]] constexpr auto begin() noexcept { return m_data.begin(); }
    [[nodiscard]] constexpr auto end() noexcept { return m_data.end(); }
    [[nodiscard]] constexpr auto begin() const noexcept { return m_data.begin(); }
    [[nodiscard]] constexpr auto end() const noexcept { return m_data.end(); }
};
} // namespace amr::containers
#endif
