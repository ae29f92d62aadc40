// AdvectionPhysics<DIM> — equation policy of linear scalar advection with the reference's fixed
// velocity {1.0, 0.5, 0.0} (include/solver/AdvectionPhysics.hpp:24,54,82).
#ifndef AMRB_SOLVER_ADVECTION_PHYSICS_HPP
#define AMRB_SOLVER_ADVECTION_PHYSICS_HPP
#include "cell_types.hpp"
#include "containers/static_vector.hpp"
#include "gpuamr_b200.h"
#include <cmath>
#include <tuple>

template <int DIM>
class AdvectionPhysics
{
public:
    static constexpr int NVAR          = 1;
    static constexpr int amrb_equation = AMRB_EQ_ADVECTION;
    using state_t                      = amr::containers::static_vector<double, NVAR>;
    using FieldTags                    = std::tuple<amr::cell::Scalar>;
    static constexpr double Velocity[3] = { 1.0, 0.5, 0.0 };

    static constexpr int getNumVars() { return NVAR; }
    static void primitiveToConservative(state_t const& prim, state_t& cons, double /*gamma*/) { cons[0] = prim[0]; }
    static void rusanovFlux(state_t const& UL, state_t const& UR, state_t& flux, int direction, double /*gamma*/)
    {
        const double v = Velocity[direction];
        flux[0]        = 0.5 * (UL[0] * v + UR[0] * v) - 0.5 * std::abs(v) * (UR[0] - UL[0]);
    }
    template <typename PatchTuple>
    static double getMaxSpeed(PatchTuple const&, std::size_t, int direction, double)
    {
        return std::abs(Velocity[direction]);
    }
};
#endif
