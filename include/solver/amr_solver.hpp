// amr_solver<TreeT, GeometryT, EquationT, DIM> — first-order finite-volume driver over the device
// patch store.  Same constructor, members and stepping contract as the reference's
// include/solver/amr_solver.hpp:62-262 (get_tree, initialize, advance, advance_batch_async,
// finish_advance_batch): a batch of k steps runs entirely on the GPU — each step ONE fused launch
// (ghost gather + Rusanov fluxes + conservative update + CFL reduction for the next step) with
// device-resident dt / remaining-time / step-count scalars — and the post-condition of every batch
// is the reference's: current buffers hold the new state with face halos filled.
#ifndef AMR_SOLVER_HPP
#define AMR_SOLVER_HPP
#include "AdvectionPhysics.hpp"
#include "EulerPhysics.hpp"
#include "containers/static_vector.hpp"
#include "cuda/amrb_check.hpp"
#include "gpuamr_b200.h"
#include "physics_system.hpp"

#include <array>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <utility>

template <typename TreeT, typename GeometryT, typename EquationT, int DIM>
class amr_solver
{
public:
    using tree_t         = TreeT;
    using geometry_t     = GeometryT;
    using equation_t     = EquationT;
    using arithmetic_t   = double;
    using patch_layout_t = typename TreeT::patch_layout_t;
    using patch_index_t  = typename TreeT::patch_index_t;
    static constexpr int NVAR = EquationT::NVAR;
    static_assert(static_cast<std::size_t>(NVAR) == TreeT::s_nvar, "equation and cell payload disagree");
    static_assert(static_cast<std::size_t>(DIM) == TreeT::s_rank, "DIM must equal the tree rank");
    // the device kernels implement exactly these two equation policies (the reference instantiates its
    // GPU kernels for the same set, src/cuda/fvm_time_step.cu:286-332); any other policy with 1 or
    // DIM+2 fields must not silently run the built-in physics
    static constexpr bool s_is_advection = std::is_same_v<EquationT, AdvectionPhysics<DIM>>;
    static constexpr bool s_is_euler     = std::is_same_v<EquationT, EulerPhysics<DIM>>;
    static_assert(s_is_advection || s_is_euler,
                  "the device path implements AdvectionPhysics<DIM> and EulerPhysics<DIM> only (no CPU fallback)");

    // same defaults as the reference constructor (amr_solver.hpp:62-69)
    amr_solver(std::size_t capacity, double gamma = 1.4, double cfl = 0.1)
        : m_tree(capacity), m_gamma(gamma), m_cfl(cfl)
    {
        if (m_tree.layout().equation != (s_is_advection ? AMRB_EQ_ADVECTION : AMRB_EQ_EULER))
            throw std::logic_error("amr_solver: the tree's cell payload does not match the equation policy");
        const auto len = GeometryT::lengths();
        double     L[3] = { 1.0, 1.0, 1.0 };
        for (int i = 0; i < DIM; ++i) L[i] = len[static_cast<std::size_t>(i)];
        check(amrb_pool_set_physics(m_tree.pool(), L, gamma, cfl), "amrb_pool_set_physics");
    }

    [[nodiscard]] auto get_tree() noexcept -> TreeT& { return m_tree; }
    [[nodiscard]] auto get_tree() const noexcept -> TreeT const& { return m_tree; }
    [[nodiscard]] auto get_gamma() const noexcept -> double { return m_gamma; }
    [[nodiscard]] auto get_cfl() const noexcept -> double { return m_cfl; }

    // ic(cell-centre coordinates) -> primitive state; stored as conservative variables in the
    // interior of every patch (amr_solver.hpp:105-145 of the reference).  Evaluated on the host
    // into the staging mirror; the next device operation uploads it.
    template <typename InitFn>
    auto initialize(InitFn&& ic) -> void
    {
        using padded_t   = typename patch_layout_t::padded_layout_t;
        constexpr auto h = patch_layout_t::halo_width();
        constexpr auto ps = padded_t::sizes();
        for (std::size_t p = 0; p != m_tree.size(); ++p)
        {
            const auto id = m_tree.get_node_index_at(p);
            const auto dx = GeometryT::cell_sizes(id);
            for (std::size_t l = 0; l != padded_t::flat_size(); ++l)
            {
                const auto m       = padded_t::multi_index(l);
                bool       interior = true;
                for (std::size_t k = 0; k != padded_t::rank(); ++k)
                    interior = interior && m[k] >= h && m[k] < ps[k] - h;
                if (!interior) continue;
                auto c = GeometryT::cell_coord(id, l);
                for (std::size_t i = 0; i != c.size(); ++i) c[i] += 0.5 * dx[i];
                const auto prim = ic(c);
                typename EquationT::state_t cons{};
                EquationT::primitiveToConservative(prim, cons, m_gamma);
                [&]<std::size_t... I>(std::index_sequence<I...>) {
                    ((m_tree.template get_patch<std::tuple_element_t<I, typename EquationT::FieldTags>>(p)[l] = cons[I]), ...);
                }(std::make_index_sequence<static_cast<std::size_t>(NVAR)>{});
            }
        }
    }

    auto advance() -> arithmetic_t
    {
        advance_batch_async(1);
        return finish_advance_batch();
    }

    auto advance_batch_async(std::size_t step_count,
                             arithmetic_t remaining_time = std::numeric_limits<arithmetic_t>::max()) -> void
    {
        m_tree.make_device_current();
        check(amrb_pool_advance_batch_async(m_tree.pool(), step_count, remaining_time), "advance_batch_async");
        m_tree.mark_device_newer();
    }

    auto finish_advance_batch(std::size_t* executed_step_count = nullptr) -> arithmetic_t
    {
        double      sum = 0.0;
        std::size_t n   = 0;
        check(amrb_pool_finish_advance_batch(m_tree.pool(), &sum, &n, nullptr, 0), "finish_advance_batch");
        if (executed_step_count != nullptr) *executed_step_count = n;
        return sum;
    }

    // cfl * min dx / speed over the current state (compute_time_step of the reference)
    [[nodiscard]] auto compute_time_step() -> arithmetic_t
    {
        m_tree.make_device_current();
        double dt = 0.0;
        check(amrb_pool_compute_dt(m_tree.pool(), &dt), "compute_time_step");
        return dt;
    }

private:
    static auto check(amrb_status st, const char* what) -> void { amr::cuda::detail::check(st, what); }

    TreeT  m_tree;
    double m_gamma;
    double m_cfl;
};

template <typename TreeT, typename GeometryT, typename EquationT>
using amr_solver_2d = amr_solver<TreeT, GeometryT, EquationT, 2>;
template <typename TreeT, typename GeometryT, typename EquationT>
using amr_solver_3d = amr_solver<TreeT, GeometryT, EquationT, 3>;
#endif // AMR_SOLVER_HPP
