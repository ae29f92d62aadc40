// amr::cell — field map types and cell payload descriptors of the FVM drivers
// (include/solver/cell_types.hpp:12-238 of the reference).  Every field is one double per cell; a
// cell type only lists its fields (structure-of-arrays storage, one patch array per field).
#ifndef AMRB_SOLVER_CELL_TYPES_HPP
#define AMRB_SOLVER_CELL_TYPES_HPP
#include <cstddef>
#include <string_view>
#include <tuple>

namespace amr::cell
{
#define AMRB_FIELD(NAME, INDEX, LABEL)                                                           \
    struct NAME                                                                                  \
    {                                                                                            \
        using type = double;                                                                     \
        static constexpr auto index() noexcept -> std::size_t { return INDEX; }                  \
        static constexpr auto name() noexcept -> std::string_view { return LABEL; }              \
    };
AMRB_FIELD(Rho, 0, "Rho")
AMRB_FIELD(Rhou, 1, "Rhou")
AMRB_FIELD(Rhov, 2, "Rhov")
AMRB_FIELD(Rhow, 3, "Rhow")
AMRB_FIELD(E2D, 3, "E2D")
AMRB_FIELD(E3D, 4, "E3D")
AMRB_FIELD(Scalar, 0, "Scalar")
#undef AMRB_FIELD

struct EulerCell2D
{
    using deconstructed_types_map_t = std::tuple<Rho, Rhou, Rhov, E2D>;
    static constexpr std::size_t n_fields = 4;
};
struct EulerCell3D
{
    using deconstructed_types_map_t = std::tuple<Rho, Rhou, Rhov, Rhow, E3D>;
    static constexpr std::size_t n_fields = 5;
};
struct AdvectionCell
{
    using deconstructed_types_map_t = std::tuple<Scalar>;
    static constexpr std::size_t n_fields = 1;
};
} // namespace amr::cell
#endif
