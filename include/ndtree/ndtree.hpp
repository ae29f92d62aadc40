// amr::ndt::tree::ndtree — the Morton-ordered leaf patch store, device resident.
//
// Same class template, member names and call protocol as the reference's
// include/ndtree/ndtree.hpp (ndtree<Cell, PatchIndex, PatchLayout, IntergridOperator>: size,
// get_node_index_at, get_patch<Map>, reconstruct_tree, halo_exchange_update, swap-free stepping
// through amr_solver, and the CUDA-mode extras get_device_buffer, sync_*_{to,from}_device,
// build_patch_levels_on_device, get_device_patch_level_buffer, get_device_refine_status_buffer,
// sync_refine_status_from_device; ndtree.hpp:94-99, 561-679, 1249-1271, 1862-1877), but a different
// machine underneath: the patch data lives in an amrb_pool on the GPU (SoA, current + next buffer),
// the topology in an amrb_tree (set of leaf ids; neighbor tables derived by key lookup and uploaded
// once per change), and refine/coarsen data motion is one device gather (amrb_pool_apply_plan).
// The host keeps a staging mirror of the current buffer for get_patch(); it is made coherent
// lazily (download before host reads, upload before device work after host writes), so drivers
// written against the CPU reference keep working without explicit sync calls.
#ifndef AMRB_NDTREE_NDTREE_HPP
#define AMRB_NDTREE_NDTREE_HPP
#include "cuda/amrb_check.hpp"
#include "gpuamr_b200.h"
#include "intergrid_operator.hpp"
#include "ndconcepts.hpp"
#include "neighbor.hpp"
#include "patch.hpp"
#include "patch_layout.hpp"
#include "patch_utils.hpp"
#include "utility/logging.hpp"

#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

namespace amr::ndt::tree
{
namespace detail
{
template <typename T, typename Tuple>
struct tuple_index;
template <typename T, typename... Ts>
struct tuple_index<T, std::tuple<T, Ts...>> : std::integral_constant<std::size_t, 0>
{
};
template <typename T, typename U, typename... Ts>
struct tuple_index<T, std::tuple<U, Ts...>>
    : std::integral_constant<std::size_t, 1 + tuple_index<T, std::tuple<Ts...>>::value>
{
};
} // namespace detail

template <typename Cell, typename Patch_Index, typename Patch_Layout, typename Intergrid_Operator>
class ndtree
{
public:
    using cell_t               = Cell;
    using patch_index_t        = Patch_Index;
    using patch_layout_t       = Patch_Layout;
    using intergrid_operator_t = Intergrid_Operator;
    using size_type            = std::size_t;
    using linear_index_t       = std::size_t;
    using fields_t             = typename Cell::deconstructed_types_map_t;
    template <typename Map>
    using patch_t = patches::patch<typename Map::type, patch_layout_t>;
    // direction system and neighbor variants (ndtree/neighbor.hpp; reference ndtree.hpp:100-125)
    using patch_direction_t               = neighbors::direction<size_type{ patch_index_t::rank() }>;
    using neighbor_patch_index_variant_t  = neighbors::neighbor_variant<size_type{ 2 }, size_type{ patch_index_t::rank() }, patch_index_t>;
    using neighbor_linear_index_variant_t = neighbors::neighbor_variant<size_type{ 2 }, size_type{ patch_index_t::rank() }, linear_index_t>;

    enum struct RefinementStatus : std::int8_t
    {
        Stable  = 0,
        Refine  = 1,
        Coarsen = 2,
    };
    using refine_status_t = RefinementStatus;

    static constexpr auto       current_buffer = 0;
    static constexpr auto       next_buffer    = 1;
    static constexpr size_type  s_nvar         = std::tuple_size_v<fields_t>;
    static constexpr size_type  s_rank         = patch_index_t::rank();
    static constexpr size_type  s_flat         = patch_layout_t::flat_size();

    static_assert(std::is_same_v<Intergrid_Operator, intergrid_operator::linear_interpolator<Patch_Layout>>,
                  "the device path implements linear_interpolator only (no CPU fallback)");
    static_assert(s_rank == patch_layout_t::rank(), "patch index and layout ranks differ");

    explicit ndtree(size_type capacity, int device = 0) : m_capacity(capacity)
    {
        constexpr auto sizes = patch_layout_t::data_layout_t::sizes();
        m_layout.rank        = static_cast<std::int32_t>(s_rank);
        for (size_type k = 0; k != 3; ++k) m_layout.size[k] = k < s_rank ? static_cast<std::int32_t>(sizes[k]) : 1;
        m_layout.halo     = static_cast<std::int32_t>(patch_layout_t::halo_width());
        m_layout.nvar     = static_cast<std::int32_t>(s_nvar);
        m_layout.equation = (s_nvar == 1) ? AMRB_EQ_ADVECTION : AMRB_EQ_EULER;
        m_layout.depth    = static_cast<std::int32_t>(patch_index_t::max_depth());
        check(amrb_tree_create(static_cast<int>(s_rank), m_layout.depth, &m_topo), "amrb_tree_create");
        check(amrb_pool_create(&m_layout, capacity, device, &m_pool), "amrb_pool_create");
        // face halos are materialised when the padded patches are observed, not after every step
        check(amrb_pool_set_lazy_halos(m_pool, 1), "amrb_pool_set_lazy_halos");
        void* host = nullptr;
        check(amrb_host_pinned_malloc(&host, capacity * s_flat * s_nvar * sizeof(double)), "host mirror");
        m_host = static_cast<double*>(host);
        std::memset(m_host, 0, capacity * s_flat * s_nvar * sizeof(double));
        void* flags = nullptr;
        check(amrb_device_malloc(&flags, capacity), "refine status buffer");
        m_device_refine_status = static_cast<std::int8_t*>(flags);
        refresh_topology();
    }
    ndtree(ndtree const&)                    = delete;
    auto operator=(ndtree const&) -> ndtree& = delete;
    ~ndtree() noexcept
    {
        amrb_pool_destroy(m_pool);
        amrb_tree_destroy(m_topo);
        amrb_host_pinned_free(m_host);
        if (m_host_next != nullptr) amrb_host_pinned_free(m_host_next);
        amrb_device_free(m_device_refine_status);
    }

    // ---- topology queries
    [[nodiscard]] auto size() const noexcept -> size_type { return m_ids.size(); }
    [[nodiscard]] auto capacity() const noexcept -> size_type { return m_capacity; }
    [[nodiscard]] auto get_node_index_at(linear_index_t i) const noexcept -> patch_index_t { return m_ids[i]; }
    [[nodiscard]] auto get_linear_index_at(patch_index_t const& id) const -> linear_index_t
    {
        // leaves are sorted by raw id
        size_type lo = 0, hi = m_ids.size();
        while (lo < hi)
        {
            const size_type mid = (lo + hi) / 2;
            if (m_ids[mid] < id)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo == m_ids.size() || !(m_ids[lo] == id)) throw std::out_of_range("patch index is not a leaf");
        return lo;
    }
    [[nodiscard]] auto get_refine_status(linear_index_t i) const noexcept -> refine_status_t
    {
        return static_cast<refine_status_t>(m_status[i]);
    }

    // ---- neighbor queries (reference ndtree.hpp:681-725).  Views decoded from the same tables the device
    // kernels read (amrb_tree_tables), fetched once per topology
    [[nodiscard]] auto get_neighbor_at(linear_index_t i, patch_direction_t const& d) const -> neighbor_patch_index_variant_t
    {
        using V = neighbor_patch_index_variant_t;
        const auto l = neighbor_linear_at(i, d);
        return std::visit(
            [this](auto const& n) -> V
            {
                using N = std::remove_cvref_t<decltype(n)>;
                using L = neighbor_linear_index_variant_t;
                if constexpr (std::is_same_v<N, typename L::same>)
                    return V{ typename V::same{ m_ids[n.id] } };
                else if constexpr (std::is_same_v<N, typename L::coarser>)
                    return V{ typename V::coarser{ m_ids[n.id], n.contact_quadrant } };
                else if constexpr (std::is_same_v<N, typename L::finer>)
                {
                    typename V::finer f{};
                    for (size_type k = 0; k != f.ids.size(); ++k) f.ids[k] = m_ids[n.ids[k]];
                    return V{ f };
                }
                else
                    return V{};
            },
            l.data);
    }
    [[nodiscard]] auto get_neighbor_at(patch_index_t const& id, patch_direction_t const& d) const -> neighbor_patch_index_variant_t
    {
        return get_neighbor_at(get_linear_index_at(id), d);
    }
    [[nodiscard]] auto neighbor_linear_index(neighbor_patch_index_variant_t const& n) const -> neighbor_linear_index_variant_t
    {
        using V = neighbor_patch_index_variant_t;
        using L = neighbor_linear_index_variant_t;
        return std::visit(
            [this](auto const& v) -> L
            {
                using N = std::remove_cvref_t<decltype(v)>;
                if constexpr (std::is_same_v<N, typename V::same>)
                    return L{ typename L::same{ get_linear_index_at(v.id) } };
                else if constexpr (std::is_same_v<N, typename V::coarser>)
                    return L{ typename L::coarser{ get_linear_index_at(v.id), v.contact_quadrant } };
                else if constexpr (std::is_same_v<N, typename V::finer>)
                {
                    typename L::finer f{};
                    for (size_type k = 0; k != f.ids.size(); ++k) f.ids[k] = get_linear_index_at(v.ids[k]);
                    return L{ f };
                }
                else
                    return L{};
            },
            n.data);
    }
    // the same query in linear indices, straight from the tables
    [[nodiscard]] auto neighbor_linear_at(linear_index_t i, patch_direction_t const& d) const -> neighbor_linear_index_variant_t
    {
        using L = neighbor_linear_index_variant_t;
        ensure_host_tables();
        constexpr size_type ND = 2 * s_rank, KF = size_type{ 1 } << (s_rank - 1);
        const size_type     e  = i * ND + static_cast<size_type>(d.index());
        switch (m_rel[e])
        {
        case AMRB_REL_SAME: return L{ typename L::same{ static_cast<linear_index_t>(m_nbr[e * KF]) } };
        case AMRB_REL_FINER:
        {
            typename L::finer f{};
            for (size_type k = 0; k != KF; ++k) f.ids[k] = static_cast<linear_index_t>(m_nbr[e * KF + k]);
            return L{ f };
        }
        case AMRB_REL_COARSER:
        {
            typename L::coarser c{ static_cast<linear_index_t>(m_nbr[e * KF]), {} };
            for (size_type k = 0; k != s_rank; ++k) c.contact_quadrant[k] = static_cast<size_type>(m_quad[e * s_rank + k]);
            return L{ c };
        }
        default: return L{};
        }
    }

    // ---- host view of the CURRENT buffer (staging mirror, lazily coherent)
    template <typename Map>
    [[nodiscard]] auto get_patch(linear_index_t i) -> patch_t<Map>&
    {
        make_host_current();
        m_host_modified = true; // a mutable reference escapes
        return host_patches<Map>()[i];
    }
    template <typename Map>
    [[nodiscard]] auto get_patch(linear_index_t i) const -> patch_t<Map> const&
    {
        const_cast<ndtree*>(this)->make_host_current();
        return const_cast<ndtree*>(this)->template host_patches<Map>()[i];
    }
    template <typename Map>
    [[nodiscard]] auto get_patch(patch_index_t const& id) -> patch_t<Map>&
    {
        return get_patch<Map>(get_linear_index_at(id));
    }
    template <typename Map>
    [[nodiscard]] auto get_patch(patch_index_t const& id) const -> patch_t<Map> const&
    {
        return get_patch<Map>(get_linear_index_at(id));
    }

    // ---- explicit buffer selection (reference ndtree.hpp:516-559).  next_buffer is a second host mirror,
    // allocated on first use: a host-side stepper fills it and calls swap_buffers(), which uploads it into
    // the device's next buffer before the pointer swap.
    template <typename Map, auto Buffer>
    [[nodiscard]] auto get_out_patch(linear_index_t i) -> patch_t<Map>&
    {
        static_assert(Buffer == current_buffer || Buffer == next_buffer, "unknown buffer tag");
        if constexpr (Buffer == current_buffer)
            return get_patch<Map>(i);
        else
        {
            ensure_host_next();
            m_host_next_modified = true;
            return reinterpret_cast<patch_t<Map>*>(m_host_next + field_index<Map>() * m_capacity * s_flat)[i];
        }
    }
    template <typename Map, auto Buffer>
    [[nodiscard]] auto get_out_patch(patch_index_t const& id) -> patch_t<Map>&
    {
        return get_out_patch<Map, Buffer>(get_linear_index_at(id));
    }
    // current <-> next (reference ndtree.hpp:1558-1579): pointer swap on the device
    auto swap_buffers() -> void
    {
        make_device_current();
        if (m_host_next_modified)
        {
            static_assert(sizeof(double) == 8);
            for (size_type f = 0; f != s_nvar; ++f)
                check(amrb_pool_upload_next(m_pool, static_cast<int>(f), 0, size(), m_host_next + f * m_capacity * s_flat),
                      "swap_buffers");
        }
        check(amrb_pool_swap_buffers(m_pool), "swap_buffers");
        if (m_host_next != nullptr) std::swap(m_host, m_host_next);
        // the mirror of the new current buffer is exact only if the host just wrote all of it
        m_device_newer       = !m_host_next_modified;
        m_host_modified      = false;
        m_host_next_modified = false;
    }

    // ---- device view (CUDA-mode extras of the reference)
    template <typename Map>
    [[nodiscard]] auto get_device_buffer() -> patch_t<Map>*
    {
        make_device_current();
        check(amrb_pool_ensure_halos(m_pool), "get_device_buffer"); // the pointee is a padded patch array
        // the raw pointer is handed to code running on other streams (the reference works on the
        // legacy default stream): order it after everything queued on the pool's stream
        check(amrb_stream_synchronize(amrb_pool_stream(m_pool)), "get_device_buffer");
        return reinterpret_cast<patch_t<Map>*>(amrb_pool_field(m_pool, static_cast<int>(field_index<Map>())));
    }
    [[nodiscard]] auto get_device_patch_level_buffer() const noexcept -> const int*
    {
        return reinterpret_cast<const int*>(amrb_pool_levels(m_pool));
    }
    [[nodiscard]] auto get_device_refine_status_buffer() noexcept -> refine_status_t*
    {
        return reinterpret_cast<refine_status_t*>(m_device_refine_status);
    }
    auto build_patch_levels_on_device() -> void {} // levels travel with the topology upload
    auto sync_refine_status_from_device() -> void
    {
        check(amrb_stream_synchronize(amrb_pool_stream(m_pool)), "sync_refine_status_from_device");
        check(amrb_device_synchronize(), "sync_refine_status_from_device");
        check(amrb_copy_device_to_host(m_status.data(), m_device_refine_status, size()), "sync_refine_status_from_device");
    }
    auto sync_current_to_device() -> void
    {
        // the device already holds the authoritative state and the mirror was not written:
        // uploading would clobber it with stale staging data
        if (m_device_newer && !m_host_modified) return;
        for (size_type f = 0; f != s_nvar; ++f)
            check(amrb_pool_upload(m_pool, static_cast<int>(f), 0, size(), m_host + f * m_capacity * s_flat),
                  "sync_current_to_device");
        m_host_modified = false;
        m_device_newer  = false;
    }
    auto sync_current_from_device() -> void
    {
        for (size_type f = 0; f != s_nvar; ++f)
            check(amrb_pool_download(m_pool, static_cast<int>(f), 0, size(), m_host + f * m_capacity * s_flat),
                  "sync_current_from_device");
        m_device_newer = false;
    }
    auto sync_next_to_device() -> void {}   // the next buffer never leaves the device
    auto sync_next_from_device() -> void {}

    // ---- halo fill of the current buffer, all fields, one launch (ndtree.hpp:1862-1877)
    auto halo_exchange_update() -> void
    {
        make_device_current();
        check(amrb_pool_halo_exchange(m_pool), "halo_exchange_update");
        m_device_newer = true;
    }

    // ---- refine / coarsen (ndtree.hpp:1249-1271).  fn is either a callable
    // (patch_index_t const&) -> refine_status_t or an object with fill_refine_flags(tree&)
    template <typename Fn>
    auto reconstruct_tree(Fn&& fn) -> void
    {
        m_status.assign(size(), 0);
        if constexpr (requires(Fn& f, ndtree& t) { f.fill_refine_flags(t); })
        {
            // the criterion writes its decisions into the device refine-status buffer
            // (compute_scalar_patch_amr_decisions_from_device): the whole reconstruct -- selection, 2:1 ripple,
            // coarsening veto, data motion, new tables -- stays on the device, the host mirrors the leaf ids
            fn.fill_refine_flags(*this);
            make_device_current();
            int         changed = 0;
            std::size_t n_new   = 0;
            check(amrb_pool_reconstruct_device(m_pool, m_device_refine_status, &changed, &n_new), "reconstruct_tree");
            if (!changed) return;
            std::vector<std::uint64_t> raw(n_new);
            check(amrb_pool_get_ids(m_pool, raw.data(), n_new), "amrb_pool_get_ids");
            check(amrb_tree_assign(m_topo, raw.data(), n_new), "amrb_tree_assign");
            m_ids.resize(n_new);
            for (size_type i = 0; i != n_new; ++i) m_ids[i] = patch_index_t{ raw[i] };
            m_status.assign(n_new, 0);
            m_host_tables_valid = false;
            m_device_newer      = true;
            return;
        }
        else
        {
            for (size_type i = 0; i != size(); ++i) m_status[i] = static_cast<std::int8_t>(fn(m_ids[i]));
        }
        make_device_current();
        int changed = 0;
        check(amrb_tree_reconstruct(m_topo, m_status.data(), m_capacity, &changed), "reconstruct_tree");
        if (!changed) return;
        const size_type        n = amrb_tree_plan_size(m_topo);
        std::vector<std::int8_t>  kind(n), child(n);
        std::vector<std::int32_t> src(n);
        check(amrb_tree_plan(m_topo, kind.data(), src.data(), child.data()), "amrb_tree_plan");
        check(amrb_pool_apply_plan(m_pool, n, kind.data(), src.data(), child.data()), "amrb_pool_apply_plan");
        refresh_topology();
        m_device_newer = true;
    }

    // ---- used by amr_solver
    [[nodiscard]] auto pool() noexcept -> amrb_pool* { return m_pool; }
    [[nodiscard]] auto layout() const noexcept -> amrb_layout const& { return m_layout; }
    auto make_device_current() -> void
    {
        if (m_host_modified) sync_current_to_device();
    }
    auto mark_device_newer() noexcept -> void { m_device_newer = true; }
    auto make_host_current() -> void
    {
        if (m_device_newer) sync_current_from_device();
    }
    template <typename Map>
    [[nodiscard]] static constexpr auto field_index() noexcept -> size_type
    {
        return detail::tuple_index<Map, fields_t>::value;
    }

private:
    static auto check(amrb_status st, const char* what) -> void { amr::cuda::detail::check(st, what); }

    template <typename Map>
    auto host_patches() noexcept -> patch_t<Map>*
    {
        return reinterpret_cast<patch_t<Map>*>(m_host + field_index<Map>() * m_capacity * s_flat);
    }

    auto refresh_topology() -> void
    {
        const size_type  n   = amrb_tree_size(m_topo);
        const auto*      raw = amrb_tree_ids(m_topo);
        m_ids.resize(n);
        for (size_type i = 0; i != n; ++i) m_ids[i] = patch_index_t{ raw[i] };
        m_status.assign(n, 0);
        // neighbor / halo tables: built on the device from the ascending leaf ids (one 8-byte-per-leaf
        // copy + one kernel) instead of the reference's host loop + metadata upload
        // (ndtree.hpp:1606-1700)
        check(amrb_pool_set_topology_from_ids(m_pool, raw, n), "amrb_pool_set_topology_from_ids");
        m_host_tables_valid = false;
    }

    auto ensure_host_tables() const -> void
    {
        if (m_host_tables_valid) return;
        constexpr size_type ND = 2 * s_rank, KF = size_type{ 1 } << (s_rank - 1);
        const size_type     n  = m_ids.size();
        m_lvl.resize(n);
        m_rel.resize(n * ND);
        m_nbr.resize(n * ND * KF);
        m_quad.resize(n * ND * s_rank);
        check(amrb_tree_tables(m_topo, m_lvl.data(), m_rel.data(), m_nbr.data(), m_quad.data()), "amrb_tree_tables");
        m_host_tables_valid = true;
    }
    auto ensure_host_next() -> void
    {
        if (m_host_next != nullptr) return;
        void* host = nullptr;
        check(amrb_host_pinned_malloc(&host, m_capacity * s_flat * s_nvar * sizeof(double)), "host mirror (next)");
        m_host_next = static_cast<double*>(host);
        std::memset(m_host_next, 0, m_capacity * s_flat * s_nvar * sizeof(double));
    }

    size_type                  m_capacity;
    amrb_layout                m_layout{};
    amrb_tree*                 m_topo = nullptr;
    amrb_pool*                 m_pool = nullptr;
    double*                    m_host = nullptr; // [field][capacity][flat], pinned
    double*                    m_host_next = nullptr; // mirror of the next buffer (get_out_patch), lazily allocated
    bool                       m_host_next_modified = false;
    mutable std::vector<std::int32_t> m_lvl, m_nbr; // host copy of the neighbor tables (get_neighbor_at)
    mutable std::vector<std::int8_t>  m_rel, m_quad;
    mutable bool                      m_host_tables_valid = false;
    std::int8_t*               m_device_refine_status = nullptr;
    std::vector<patch_index_t> m_ids;
    std::vector<std::int8_t>   m_status;
    bool                       m_host_modified = false; // host mirror may hold newer data
    bool                       m_device_newer  = false; // device holds newer data than the mirror
};
} // namespace amr::ndt::tree
#endif
