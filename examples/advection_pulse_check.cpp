// Property checks of the reference's tests/advection_equation_amr.t.cpp (PulseCenterMovement and
// PulseCenterMovementAMR, :50-232) as a plain program against the B200 headers: a Gaussian pulse
// advected with velocity {1.0, 0.5} must have its arg-max within one cell width of x0 + v t, without
// and with refine/coarsen every step.  Written against the CPU-style API (no explicit sync calls):
// the tree's staging mirror keeps host reads coherent.  Exit code 0 = all checks passed.
#include "containers/static_layout.hpp"
#include "containers/static_shape.hpp"
#include "containers/static_vector.hpp"
#include "morton/morton_id.hpp"
#include "ndtree/intergrid_operator.hpp"
#include "ndtree/ndtree.hpp"
#include "ndtree/patch_layout.hpp"
#include "ndtree/patch_utils.hpp"
#include "solver/AdvectionPhysics.hpp"
#include "solver/amr_solver.hpp"
#include "solver/cell_types.hpp"
#include "solver/physics_system.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>

template <std::size_t N>
struct Config
{
    static constexpr int                   DIM     = 2;
    static constexpr std::array<double, 2> lengths = { 1.0, 1.0 };
    using shape_t        = amr::containers::static_shape<N, N>;
    using layout_t       = amr::containers::static_layout<shape_t>;
    using patch_index_t  = amr::ndt::morton::morton_id<7u, 2u>;
    using patch_layout_t = amr::ndt::patches::patch_layout<layout_t, 2>;
    using op_t           = amr::ndt::intergrid_operator::linear_interpolator<patch_layout_t>;
    using tree_t     = amr::ndt::tree::ndtree<amr::cell::AdvectionCell, patch_index_t, patch_layout_t, op_t>;
    using geometry_t = amr::ndt::solver::physics_system<patch_index_t, patch_layout_t, lengths>;
    using solver_t   = amr_solver<tree_t, geometry_t, AdvectionPhysics<2>, DIM>;
};

template <typename C>
static bool run(bool with_amr, const char* label)
{
    typename C::solver_t solver(2000, 1.0, 0.4);
    auto&                tree = solver.get_tree();
    const double         x0 = 0.2, y0 = 0.2, t_end = 0.5;
    auto ic = [&](auto const& c) -> amr::containers::static_vector<double, 1>
    { return { std::exp(-((c[0] - x0) * (c[0] - x0) + (c[1] - y0) * (c[1] - y0)) / 0.005) }; };
    auto criterion = [&](typename C::patch_index_t const& id)
    {
        auto const& patch = tree.template get_patch<amr::cell::Scalar>(id);
        double      mx    = 0.0;
        for (auto v : patch.data()) mx = std::max(mx, v);
        using st = typename C::tree_t::refine_status_t;
        if (mx > 0.1 && id.level() < 5) return st::Refine;
        if (mx < 0.05 && id.level() > 1) return st::Coarsen;
        return st::Stable;
    };
    solver.initialize(ic);
    tree.halo_exchange_update();
    double t = 0.0;
    int    steps = 0;
    while (t < t_end && steps < 100000)
    {
        const double dt = solver.advance();
        tree.halo_exchange_update();
        if (with_amr)
        {
            tree.reconstruct_tree(criterion);
            tree.halo_exchange_update();
        }
        t += dt;
        ++steps;
    }
    double                max_val = -1.0;
    std::array<double, 2> pos{};
    for (std::size_t p = 0; p < tree.size(); ++p)
    {
        const auto  id    = tree.get_node_index_at(p);
        auto const& patch = tree.template get_patch<amr::cell::Scalar>(p);
        for (std::size_t l = 0; l < C::patch_layout_t::flat_size(); ++l)
        {
            if (amr::ndt::utils::patches::is_halo_cell<typename C::patch_layout_t>(l)) continue;
            if (patch[l] > max_val)
            {
                max_val       = patch[l];
                const auto c  = C::geometry_t::cell_coord(id, l);
                const auto dx = C::geometry_t::cell_sizes(id);
                pos           = { c[0] + 0.5 * dx[0], c[1] + 0.5 * dx[1] };
            }
        }
    }
    const auto   dx0 = C::geometry_t::cell_sizes(tree.get_node_index_at(0));
    const double ex = x0 + AdvectionPhysics<2>::Velocity[0] * t, ey = y0 + AdvectionPhysics<2>::Velocity[1] * t;
    const bool   ok = std::abs(pos[0] - ex) <= dx0[0] && std::abs(pos[1] - ey) <= dx0[1] && max_val > 0.0;
    std::printf("%-28s steps %5d patches %4zu t %.4f max %.4f at (%.4f, %.4f) expected (%.4f, %.4f) tol %.4f : %s\n",
                label, steps, tree.size(), t, max_val, pos[0], pos[1], ex, ey, dx0[0], ok ? "OK" : "FAIL");
    return ok;
}

int main()
{
    bool ok = true;
    ok &= run<Config<10>>(false, "PulseCenterMovement<10>");
    ok &= run<Config<10>>(true, "PulseCenterMovementAMR<10>");
    std::printf(ok ? "ALL OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
