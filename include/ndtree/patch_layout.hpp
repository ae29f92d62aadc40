// amr::ndt::patches::patch_layout<DataLayout, Halo> — a patch is one padded row-major tensor,
// every dim padded by 2*halo (include/ndtree/patch_layout.hpp:14-116, containers/container_utils.hpp:
// 56-64 of the reference).
#ifndef AMRB_NDTREE_PATCH_LAYOUT_HPP
#define AMRB_NDTREE_PATCH_LAYOUT_HPP
#include "containers/static_layout.hpp"
#include <utility>

namespace amr::ndt::patches
{
namespace detail
{
template <typename Shape, std::size_t Pad, typename Seq>
struct padded_shape;
template <typename Shape, std::size_t Pad, std::size_t... I>
struct padded_shape<Shape, Pad, std::index_sequence<I...>>
{
    using type = containers::static_shape<(Shape::sizes()[I] + 2 * Pad)...>;
};
} // namespace detail

template <typename Data_Layout, auto Halo_Width>
class patch_layout
{
public:
    using data_layout_t = Data_Layout;
    using shape_t       = typename Data_Layout::shape_t;
    using size_type     = typename Data_Layout::size_type;
    using index_t       = size_type;
    using padded_shape_t =
        typename detail::padded_shape<shape_t, static_cast<std::size_t>(Halo_Width),
                                      std::make_index_sequence<shape_t::rank()>>::type;
    using padded_layout_t = containers::static_layout<padded_shape_t>;

    [[nodiscard]] static constexpr auto rank() noexcept { return Data_Layout::rank(); }
    [[nodiscard]] static constexpr auto halo_width() noexcept -> size_type { return static_cast<size_type>(Halo_Width); }
    [[nodiscard]] static constexpr auto flat_size() noexcept -> size_type { return padded_layout_t::flat_size(); }
    [[nodiscard]] static constexpr auto data_size() noexcept -> size_type { return Data_Layout::flat_size(); }
};
} // namespace amr::ndt::patches
#endif
