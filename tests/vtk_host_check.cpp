// TEST INFRASTRUCTURE (host only, no GPU): feeds include/ndtree/vtk_print.hpp of this repo with a
// minimal tree view (leaf ids + padded patch data read from a file) and writes
// vtk_output/check.vtk, which tests/test_vtk_print.py compares byte for byte with the file the
// reference's own vtk_print wrote for the same tree (tests/golden/vtk/*.npz).
//   -DCFG_RANK=2|3 -DCFG_S=<cells per dim> -DCFG_H=<halo> -DCFG_DEPTH=<depth> -DCFG_EQ=0|1 -DCFG_L=<length>
// input file: u64 n_patches | u64 ids[n] | f64 data[nvar][n][flat]
#include "containers/static_layout.hpp"
#include "containers/static_shape.hpp"
#include "morton/morton_id.hpp"
#include "ndtree/patch.hpp"
#include "ndtree/patch_layout.hpp"
#include "ndtree/vtk_print.hpp"
#include "solver/cell_types.hpp"
#include "solver/physics_system.hpp"

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <tuple>
#include <vector>

namespace
{
constexpr int RANK = CFG_RANK;
template <int R>
struct shape_of;
template <>
struct shape_of<2>
{
    using type = amr::containers::static_shape<std::size_t{ CFG_S }, std::size_t{ CFG_S }>;
    static constexpr std::array<double, 2> lengths = { CFG_L, CFG_L };
    using euler = amr::cell::EulerCell2D;
};
template <>
struct shape_of<3>
{
    using type = amr::containers::static_shape<std::size_t{ CFG_S }, std::size_t{ CFG_S }, std::size_t{ CFG_S }>;
    static constexpr std::array<double, 3> lengths = { CFG_L, CFG_L, CFG_L };
    using euler = amr::cell::EulerCell3D;
};
using S_              = shape_of<RANK>;
using layout_t        = amr::containers::static_layout<typename S_::type>;
using patch_index_t   = amr::ndt::morton::morton_id<CFG_DEPTH, unsigned(RANK)>;
using patch_layout_t_ = amr::ndt::patches::patch_layout<layout_t, std::size_t{ CFG_H }>;
#if CFG_EQ == 0
using cell_t = amr::cell::AdvectionCell;
#else
using cell_t = typename S_::euler;
#endif
using physics_t = amr::ndt::solver::physics_system<patch_index_t, patch_layout_t_, S_::lengths>;

// the part of the ndtree interface vtk_print uses
struct tree_view
{
    using patch_layout_t = patch_layout_t_;
    using fields_t       = typename cell_t::deconstructed_types_map_t;
    template <typename Map>
    using patch_t = amr::ndt::patches::patch<typename Map::type, patch_layout_t>;
    static constexpr std::size_t flat = patch_layout_t::flat_size();

    std::vector<patch_index_t> ids;
    std::vector<double>        data; // [field][patch][flat]

    [[nodiscard]] auto size() const noexcept -> std::size_t { return ids.size(); }
    [[nodiscard]] auto get_node_index_at(std::size_t i) const noexcept -> patch_index_t { return ids[i]; }
    template <typename Map>
    [[nodiscard]] auto get_patch(std::size_t i) const -> patch_t<Map> const&
    {
        return *reinterpret_cast<patch_t<Map> const*>(data.data() + (Map::index() * ids.size() + i) * flat);
    }
};
} // namespace

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::uint64_t n = 0;
    in.read(reinterpret_cast<char*>(&n), 8);
    std::vector<std::uint64_t> raw(n);
    in.read(reinterpret_cast<char*>(raw.data()), static_cast<std::streamsize>(8 * n));
    tree_view t;
    for (auto r : raw) t.ids.emplace_back(r);
    t.data.resize(std::tuple_size_v<tree_view::fields_t> * n * tree_view::flat);
    in.read(reinterpret_cast<char*>(t.data.data()), static_cast<std::streamsize>(8 * t.data.size()));
    if (!in) return 3;
    amr::ndt::print::vtk_print<physics_t> printer("check");
    printer.print(t, ".vtk");
    return 0;
}
