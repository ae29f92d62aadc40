// amr::ndt::concepts — what the template policies of the tree / solver must provide.
//
// The reference constrains ndtree<Cell, PatchIndex, PatchLayout, IntergridOperator> and amr_solver with
// these concepts (include/ndtree/ndconcepts.hpp:13-127); drivers written against the reference name them
// in their own templates, so the same names exist here.  Each concept lists only what THIS implementation
// reads from the policy type (a policy that satisfies the reference's concept satisfies these).
#ifndef AMRB_NDTREE_NDCONCEPTS_HPP
#define AMRB_NDTREE_NDCONCEPTS_HPP
#include <concepts>
#include <cstddef>
#include <string_view>
#include <tuple>

namespace amr::ndt::concepts
{
// one scalar field of the cell payload (solver/cell_types.hpp): value type + position in the SoA store
template <typename T>
concept MapType = requires {
    typename T::type;
    { T::index() } -> std::convertible_to<std::size_t>;
    { T::name() } -> std::convertible_to<std::string_view>;
};

namespace detail
{
template <typename>
inline constexpr bool all_map_types = false;
template <template <class...> class Tuple, class... Ts>
inline constexpr bool all_map_types<Tuple<Ts...>> = (MapType<Ts> && ...);
} // namespace detail

template <typename T>
concept MapTypeTuple = detail::all_map_types<T>;

// cell payload = a tuple of fields the tree stores as structure of arrays
template <typename T>
concept DeconstructibleType = requires { typename T::deconstructed_types_map_t; } &&
                              MapTypeTuple<typename T::deconstructed_types_map_t>;

// patch index (morton/morton_id.hpp): ordering key of the linear tree
template <typename I>
concept PatchIndex = requires(I const i) {
    typename I::size_type;
    { I::rank() } -> std::integral;
    { I::fanout() } -> std::integral;
    { I::max_depth() } -> std::integral;
    { I::root() } -> std::same_as<I>;
    { I::parent_of(i) } -> std::same_as<I>;
    { i.id() } -> std::unsigned_integral;
    { i.level() } -> std::integral;
    { i < i } -> std::convertible_to<bool>;
} && std::equality_comparable<I>;

template <typename L>
concept PatchLayout = requires {
    typename L::index_t;
    typename L::size_type;
    typename L::data_layout_t;
    typename L::padded_layout_t;
    { L::rank() } -> std::integral;
    { L::flat_size() } -> std::integral;
    { L::halo_width() } -> std::integral;
};

template <typename P>
concept Patch = requires(P p, P const cp, typename P::size_type i) {
    typename P::value_type;
    typename P::container_t;
    { p.data() } -> std::same_as<typename P::container_t&>;
    { cp.data() } -> std::same_as<typename P::container_t const&>;
    { cp[i] } -> std::same_as<typename P::value_type const&>;
};

template <typename D>
concept Direction = requires(D const d) {
    typename D::index_t;
    typename D::size_type;
    { D::rank() } -> std::same_as<typename D::size_type>;
    { D::elements() } -> std::same_as<typename D::size_type>;
    { D::unit_vector(d) } -> std::same_as<typename D::vector_t>;
    { d.dimension() } -> std::same_as<typename D::index_t>;
    { d.is_negative() } -> std::same_as<bool>;
    { d.is_positive() } -> std::same_as<bool>;
};

// transfer operator between refinement levels; the device path implements linear_interpolator only
template <typename IO>
concept IntergridOperator = requires {
    typename IO::patch_layout_t;
    typename IO::index_t;
};

// policy object naming the four face operators of the halo exchange (patch_utils.hpp:303-441)
template <typename HEO>
concept HaloExchangeOperator = requires {
    HEO::boundary;
    HEO::same;
    HEO::finer;
    HEO::coarser;
};

template <typename T>
concept TreeType = requires {
    typename T::linear_index_t;
    typename T::patch_layout_t;
    typename T::patch_index_t;
} && PatchLayout<typename T::patch_layout_t>;
} // namespace amr::ndt::concepts
#endif
