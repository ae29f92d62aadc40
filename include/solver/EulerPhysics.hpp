// EulerPhysics<DIM> — equation policy of the compressible Euler system (ideal gas).
// Public contract of include/solver/EulerPhysics.hpp of the reference: NVAR, FieldTags,
// primitiveToConservative, rusanovFlux, getMaxSpeed.  The device kernels implement the same
// formulas (gpu-amr_b200/csrc/amrb_step_euler.cuh); the host versions below serve initialisation
// and host-side checks.
#ifndef AMRB_SOLVER_EULER_PHYSICS_HPP
#define AMRB_SOLVER_EULER_PHYSICS_HPP
#include "cell_types.hpp"
#include "containers/static_vector.hpp"
#include "gpuamr_b200.h"
#include <algorithm>
#include <cmath>
#include <tuple>
#include <type_traits>

template <int DIM>
class EulerPhysics
{
    static_assert(DIM == 2 || DIM == 3);

public:
    static constexpr int NVAR          = DIM + 2;
    static constexpr int amrb_equation = AMRB_EQ_EULER;
    using state_t                      = amr::containers::static_vector<double, NVAR>;
    using FieldTags = std::conditional_t<
        DIM == 2, std::tuple<amr::cell::Rho, amr::cell::Rhou, amr::cell::Rhov, amr::cell::E2D>,
        std::tuple<amr::cell::Rho, amr::cell::Rhou, amr::cell::Rhov, amr::cell::Rhow, amr::cell::E3D>>;

    static constexpr int getNumVars() { return NVAR; }

    // [rho, u, v, (w), p] -> [rho, rho u, rho v, (rho w), E]
    static void primitiveToConservative(state_t const& prim, state_t& cons, double gamma)
    {
        const double rho = prim[0];
        double       v2  = 0.0;
        cons[0]          = rho;
        for (int d = 0; d < DIM; ++d)
        {
            cons[1 + d] = rho * prim[1 + d];
            v2 += prim[1 + d] * prim[1 + d];
        }
        cons[DIM + 1] = prim[DIM + 1] / (gamma - 1.0) + 0.5 * rho * v2;
    }

    struct side_t
    {
        double inv_rho, p, a, un;
    };
    static side_t side(state_t const& U, int direction, double gamma)
    {
        side_t s{};
        s.inv_rho = 1.0 / U[0];
        double K  = 0.0;
        for (int d = 0; d < DIM; ++d) K += U[1 + d] * U[1 + d];
        K *= 0.5 * s.inv_rho;
        s.p  = (gamma - 1.0) * (U[DIM + 1] - K);
        s.a  = std::sqrt(gamma * s.p * s.inv_rho);
        s.un = U[1 + direction] * s.inv_rho;
        return s;
    }

    // local Lax-Friedrichs flux across a face normal to `direction`
    static void rusanovFlux(state_t const& UL, state_t const& UR, state_t& flux, int direction, double gamma)
    {
        const side_t L = side(UL, direction, gamma), R = side(UR, direction, gamma);
        const double smax = std::max(std::abs(L.un) + L.a, std::abs(R.un) + R.a);
        flux[0] = 0.5 * (UL[1 + direction] + UR[1 + direction] - smax * (UR[0] - UL[0]));
        for (int d = 0; d < DIM; ++d)
        {
            double fl = UL[1 + d] * L.un, fr = UR[1 + d] * R.un;
            if (d == direction)
            {
                fl += L.p;
                fr += R.p;
            }
            flux[1 + d] = 0.5 * (fl + fr - smax * (UR[1 + d] - UL[1 + d]));
        }
        const double el = L.un * (UL[DIM + 1] + L.p), er = R.un * (UR[DIM + 1] + R.p);
        flux[DIM + 1]   = 0.5 * (el + er - smax * (UR[DIM + 1] - UL[DIM + 1]));
    }

    // |u_dir| + a of the cell `idx` of a tuple of field patches
    template <typename PatchTuple>
    static double getMaxSpeed(PatchTuple const& patches, std::size_t idx, int direction, double gamma)
    {
        state_t U{};
        [&]<std::size_t... I>(std::index_sequence<I...>) { ((U[I] = std::get<I>(patches)[idx]), ...); }
        (std::make_index_sequence<NVAR>{});
        const side_t s = side(U, direction, gamma);
        return std::abs(s.un) + s.a;
    }
};

using EulerPhysics2D = EulerPhysics<2>;
using EulerPhysics3D = EulerPhysics<3>;

namespace Direction
{
constexpr int X = 0;
constexpr int Y = 1;
constexpr int Z = 2;
} // namespace Direction
#endif
