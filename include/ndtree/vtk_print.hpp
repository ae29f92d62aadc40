// amr::ndt::print::vtk_print<PhysicsSystem> — legacy-VTK (BINARY, big-endian) dump of the leaf
// patches, one cell per interior cell, so that the reference's examples keep writing their
// vtk_output/<name><ext> files when built against this tree.
//
// Same class, constructor and print(tree, extension) call and the same file layout as the
// reference's include/ndtree/vtk_print.hpp:16-276: header lines, POINTS (2^rank corner points
// per cell, in the reference's corner order), CELLS, CELL_TYPES (9 = quad, 11 = voxel), CELL_DATA
// with cell_index, is_halo and one SCALARS array per field in interior iteration order (last
// layout dim fastest).  Written differently: each section is assembled big-endian in one memory
// block and handed to the stream in a single write; the patch data comes from the tree's host
// mirror, which the tree refreshes from the device pool on first access (one bulk download).
#ifndef AMRB_NDTREE_VTK_PRINT_HPP
#define AMRB_NDTREE_VTK_PRINT_HPP
#include "patch_utils.hpp"
#include "utility/logging.hpp"

#include <array>
#include <bit>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <vector>

namespace amr::ndt::print
{
namespace detail
{
// growing block of big-endian values
class be_block
{
public:
    explicit be_block(std::size_t bytes) { m_bytes.reserve(bytes); }
    template <typename T>
    auto put(T v) -> void
    {
        static_assert(std::is_trivially_copyable_v<T>);
        std::array<char, sizeof(T)> raw;
        std::memcpy(raw.data(), &v, sizeof(T));
        if constexpr (std::endian::native == std::endian::little)
            for (std::size_t k = sizeof(T); k-- > 0;) m_bytes.push_back(raw[k]);
        else
            m_bytes.insert(m_bytes.end(), raw.begin(), raw.end());
    }
    auto flush(std::ostream& os) -> void
    {
        os.write(m_bytes.data(), static_cast<std::streamsize>(m_bytes.size()));
        m_bytes.clear();
    }

private:
    std::vector<char> m_bytes;
};

template <typename T>
constexpr auto vtk_type_name() -> const char*
{
    if constexpr (std::is_same_v<T, float>)
        return "float";
    else if constexpr (std::is_same_v<T, double>)
        return "double";
    else
    {
        static_assert(std::is_same_v<T, int>, "vtk_print: unsupported scalar type");
        return "int";
    }
}
} // namespace detail

template <typename Physics_System>
struct vtk_print
{
    using physics_system_t = Physics_System;

    explicit vtk_print(std::string_view base_filename) : m_base_filename{ base_filename }
    {
        const std::filesystem::path dir = "vtk_output";
        std::filesystem::create_directory(dir);
        std::cout << "VTK files will be saved to: " << std::filesystem::absolute(dir) << std::endl;
    }

    template <typename Tree>
    auto print(Tree const& tree, std::string filename_extension) const -> void
    {
        const std::string path = "vtk_output/" + m_base_filename + filename_extension;
        std::ofstream     file(path, std::ios::binary);
        if (!file.is_open()) throw std::runtime_error("Cannot open file: " + path);
        file << "# vtk DataFile Version 3.0\n"
             << "AMR Tree Structure\n"
             << "BINARY\n"
             << "DATASET UNSTRUCTURED_GRID\n";
        write_grid(file, tree);
    }

private:
    template <typename Tree>
    static auto write_grid(std::ofstream& file, Tree const& tree) -> void
    {
        using layout_t   = typename Tree::patch_layout_t;
        using padded_t   = typename layout_t::padded_layout_t;
        using fields_t   = typename Tree::fields_t;
        using real_t     = typename std::tuple_element_t<0, fields_t>::type;
        constexpr auto R = static_cast<std::size_t>(layout_t::rank());
        static_assert(R == 2 || R == 3, "vtk_print handles 2D and 3D trees");
        constexpr std::size_t corners   = std::size_t{ 1 } << R;
        constexpr std::size_t per_patch = layout_t::data_size();
        constexpr auto        h         = layout_t::halo_width();
        constexpr auto        data      = layout_t::data_layout_t::sizes();
        constexpr auto        pstride   = padded_t::strides();

        const std::size_t n_patches = tree.size();
        const std::size_t n_cells   = n_patches * per_patch;

        // interior cells of a patch in iteration order (last layout dim fastest): padded linear
        // index and the offset of the cell in cell units per PHYSICAL axis (axis i <-> layout
        // dim R-1-i)
        std::vector<std::size_t>                linear(per_patch);
        std::vector<std::array<std::size_t, R>> offs(per_patch);
        for (std::size_t c = 0; c != per_patch; ++c)
        {
            std::size_t rest = c, lin = 0;
            for (std::size_t k = R; k-- > 0;)
            {
                const std::size_t i = rest % data[k];
                rest /= data[k];
                lin += (i + h) * pstride[k];
                offs[c][R - 1 - k] = i;
            }
            linear[c] = lin;
        }

        // ---- POINTS: quad corners (x0 y0)(x1 y0)(x1 y1)(x0 y1); voxel corners with x fastest
        file << "POINTS " << n_cells * corners << ' ' << detail::vtk_type_name<real_t>() << '\n';
        {
            detail::be_block blk(per_patch * corners * 3 * sizeof(real_t));
            for (std::size_t p = 0; p != n_patches; ++p)
            {
                const auto id     = tree.get_node_index_at(p);
                const auto origin = physics_system_t::patch_coord(id);
                const auto dx     = physics_system_t::cell_sizes(id);
                for (std::size_t c = 0; c != per_patch; ++c)
                {
                    double lo[3] = { 0.0, 0.0, 0.0 }, hi[3] = { 0.0, 0.0, 0.0 };
                    for (std::size_t a = 0; a != R; ++a)
                    {
                        lo[a] = origin[a] + static_cast<double>(offs[c][a]) * dx[a];
                        hi[a] = lo[a] + dx[a];
                    }
                    auto corner = [&](bool ux, bool uy, bool uz) {
                        blk.put(static_cast<real_t>(ux ? hi[0] : lo[0]));
                        blk.put(static_cast<real_t>(uy ? hi[1] : lo[1]));
                        blk.put(static_cast<real_t>(R == 3 ? (uz ? hi[2] : lo[2]) : 0.0));
                    };
                    if constexpr (R == 2)
                    {
                        corner(false, false, false);
                        corner(true, false, false);
                        corner(true, true, false);
                        corner(false, true, false);
                    }
                    else
                    {
                        for (unsigned k = 0; k != 8; ++k) corner(k & 1u, k & 2u, k & 4u);
                    }
                }
                blk.flush(file);
            }
        }

        // ---- CELLS / CELL_TYPES / cell_index
        file << "CELLS " << n_cells << ' ' << n_cells * (corners + 1) << '\n';
        {
            detail::be_block blk((corners + 1) * sizeof(int) * per_patch);
            for (std::size_t c = 0; c != n_cells; ++c)
            {
                blk.put(static_cast<int>(corners));
                for (std::size_t j = 0; j != corners; ++j) blk.put(static_cast<int>(c * corners + j));
                if ((c + 1) % per_patch == 0) blk.flush(file);
            }
            blk.flush(file);
        }
        file << "CELL_TYPES " << n_cells << '\n';
        {
            detail::be_block blk(n_cells * sizeof(int));
            for (std::size_t c = 0; c != n_cells; ++c) blk.put(static_cast<int>(R == 2 ? 9 : 11));
            blk.flush(file);
        }
        file << "CELL_DATA " << n_cells << '\n';
        file << "SCALARS cell_index int 1\nLOOKUP_TABLE default\n";
        {
            detail::be_block blk(n_cells * sizeof(int));
            for (std::size_t c = 0; c != n_cells; ++c) blk.put(static_cast<int>(c));
            blk.flush(file);
        }
        // the reference evaluates is_halo_cell on the first data_size PADDED linear indices of
        // every patch (vtk_print.hpp:226-237); reproduced as is so the files compare equal
        file << "SCALARS is_halo int 1\nLOOKUP_TABLE default\n";
        {
            detail::be_block blk(n_cells * sizeof(int));
            for (std::size_t p = 0; p != n_patches; ++p)
                for (std::size_t j = 0; j != per_patch; ++j)
                    blk.put(static_cast<int>(utils::patches::is_halo_cell<layout_t>(j)));
            blk.flush(file);
        }

        // ---- one SCALARS array per field, interior cells only
        [&]<std::size_t... I>(std::index_sequence<I...>) {
            (write_field<std::tuple_element_t<I, fields_t>>(file, tree, linear), ...);
        }(std::make_index_sequence<std::tuple_size_v<fields_t>>{});
    }

    template <typename Map, typename Tree>
    static auto write_field(std::ofstream& file, Tree const& tree, std::vector<std::size_t> const& linear)
        -> void
    {
        using value_t = typename Map::type;
        file << "SCALARS " << Map::name() << ' ' << detail::vtk_type_name<value_t>() << " 1\n"
             << "LOOKUP_TABLE default\n";
        detail::be_block blk(linear.size() * sizeof(value_t));
        for (std::size_t p = 0; p != tree.size(); ++p)
        {
            auto const& patch = tree.template get_patch<Map>(p);
            for (const std::size_t l : linear) blk.put(static_cast<value_t>(patch[l]));
            blk.flush(file);
        }
    }

    std::string m_base_filename;
};
} // namespace amr::ndt::print
#endif
