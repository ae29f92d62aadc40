// amr::ndt::patches::patch<T, PatchLayout> — host mirror of one field-patch: a flat padded buffer
// (sizeof(patch) == flat_size * sizeof(T), as the reference asserts in
// examples/fvm_solver_advection.e.cpp:45-48).
#ifndef AMRB_NDTREE_PATCH_HPP
#define AMRB_NDTREE_PATCH_HPP
#include "patch_layout.hpp"
#include <array>

namespace amr::ndt::patches
{
template <typename T, typename Patch_Layout>
class patch
{
public:
    using value_type     = T;
    using patch_layout_t = Patch_Layout;
    using size_type      = typename Patch_Layout::size_type;
    using container_t    = std::array<T, Patch_Layout::flat_size()>;

    [[nodiscard]] constexpr auto operator[](size_type i) noexcept -> T& { return m_data[i]; }
    [[nodiscard]] constexpr auto operator[](size_type i) const noexcept -> T const& { return m_data[i]; }
    [[nodiscard]] constexpr auto data() noexcept -> container_t& { return m_data; }
    [[nodiscard]] constexpr auto data() const noexcept -> container_t const& { return m_data; }

private:
    container_t m_data;
};
} // namespace amr::ndt::patches
#endif
