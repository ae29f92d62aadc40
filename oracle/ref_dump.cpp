// TEST INFRASTRUCTURE — not product code.
//
// Driver that runs the UNMODIFIED reference (bgce-cse/gpu-amr headers under
// /root/reference/include, compiled where they lie; nothing is copied) through a
// scripted sequence of tree reconstructions / halo exchanges / solver steps and
// dumps what it did: leaf ids, per-(patch,direction) neighbor tables, padded
// patch data and the dt sequence.  The dumps become the golden fixtures under
// tests/golden/ (see oracle/gen_golden.py) that pin oracle/amr_oracle.c and,
// through it, the CUDA path.
//
// Built by oracle/Makefile into oracle/_ref/ref_dump_<cfg> (git-ignored).
// Reference entry points exercised:
//   ndtree::reconstruct_tree        include/ndtree/ndtree.hpp:1249-1271
//   ndtree::halo_exchange_update    include/ndtree/ndtree.hpp:1862-1877 (CPU: 1581-1596)
//   ndtree::get_neighbor_at / neighbor_linear_index   ndtree.hpp:681-725
//   amr_solver::initialize/advance  include/solver/amr_solver.hpp:105-153
//
// Config is compile-time (the reference is a template library):
//   -DCFG_RANK=2|3 -DCFG_S=<cells per dim> -DCFG_H=<halo> -DCFG_DEPTH=<morton depth>
//   -DCFG_EQ=0 (advection) | 1 (euler)  -DCFG_L=<domain length>
//
// Script (argv[1], one op per line):
//   A                      reconstruct_tree(refine every leaf)
//   H seed pr pc minl maxl reconstruct_tree(hash flags): u=splitmix64(id^seed)%1000;
//                          u<pr && level<maxl -> Refine; pr<=u<pr+pc && level>minl -> Coarsen
//   B r minl maxl c0 c1 [c2]  reconstruct_tree(ball): refine if |patch centre/L - c| < r
//                          and level<maxl; coarsen if outside 2r and level>minl
//   R id id ...            reconstruct_tree(refine exactly the listed leaf ids)
//   K id id ...            reconstruct_tree(flag exactly the listed leaf ids Coarsen)
//   X                      halo_exchange_update()
//   P                      fill every cell of every padded patch, field f, with the exact
//                          integer code  (f*2^20 + patch)*2^12 + cell   (index probe)
//   I                      solver.initialize(IC) (pulse / gaussian, see ic())
//   S n                    n x advance(); dt of every step recorded
//   T n                    time n x advance() and print one JSON line (updates/s)
//   V name                 vtk_print(tree) -> vtk_output/<name>.vtk in the working directory
//   D tag                  dump ids, neighbor tables, padded data, under "tag/"
//
// With -DAMR_ENABLE_CUDA_AMR=1 (oracle/Makefile target _ref/ref_cuda_bench_*: this file + the
// reference's own src/cuda/*.cu compiled for sm_100) the same script drives the reference's CUDA
// backend, with the sync calls the reference's benchmark makes
// (benchmark/bench_fvm_solver_integration.b.cpp:181-197): before I the mirror is pushed / pulled
// around every halo exchange, from I on the state lives on the device and D pulls it back.  That is
// the GPU-vs-GPU baseline of SURVEY 8d ("the kernel to beat on the same box").
// Output (argv[2]): flat sequence of records
//   u32 name_len | name | u32 dtype(0=f64,1=i64,2=i32,3=i8,4=u64) | u32 ndim | u64 shape[ndim] | raw data

#include "containers/static_layout.hpp"
#include "containers/static_shape.hpp"
#include "containers/static_vector.hpp"
#include "morton/morton_id.hpp"
#include "ndtree/intergrid_operator.hpp"
#include "ndtree/ndtree.hpp"
#include "ndtree/patch_layout.hpp"
#include "ndtree/vtk_print.hpp"
#include "solver/AdvectionPhysics.hpp"
#include "solver/EulerPhysics.hpp"
#include "solver/amr_solver.hpp"
#include "solver/cell_types.hpp"
#include "solver/physics_system.hpp"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <variant>
#include <vector>

#ifndef CFG_RANK
#    define CFG_RANK 2
#endif
#ifndef CFG_S
#    define CFG_S 8
#endif
#ifndef CFG_H
#    define CFG_H 1
#endif
#ifndef CFG_DEPTH
#    define CFG_DEPTH 7
#endif
#ifndef CFG_EQ
#    define CFG_EQ 0
#endif
#ifndef CFG_L
#    define CFG_L 1.0
#endif

namespace
{

constexpr int         RANK  = CFG_RANK;
constexpr std::size_t S     = CFG_S;
constexpr std::size_t HALO  = CFG_H;
constexpr unsigned    DEPTH = CFG_DEPTH;
constexpr double      L     = CFG_L;

template <int R>
struct cfg;

template <>
struct cfg<2>
{
    using shape_t                                  = amr::containers::static_shape<S, S>;
    static constexpr std::array<double, 2> lengths = { L, L };
#if CFG_EQ == 0
    using cell_t = amr::cell::AdvectionCell;
    using eq_t   = AdvectionPhysics<2>;
#else
    using cell_t = amr::cell::EulerCell2D;
    using eq_t   = EulerPhysics<2>;
#endif
};

template <>
struct cfg<3>
{
    using shape_t                                  = amr::containers::static_shape<S, S, S>;
    static constexpr std::array<double, 3> lengths = { L, L, L };
#if CFG_EQ == 0
    using cell_t = amr::cell::AdvectionCell;
    using eq_t   = AdvectionPhysics<3>;
#else
    using cell_t = amr::cell::EulerCell3D;
    using eq_t   = EulerPhysics<3>;
#endif
};

using C              = cfg<RANK>;
using layout_t       = amr::containers::static_layout<typename C::shape_t>;
using patch_index_t  = amr::ndt::morton::morton_id<DEPTH, unsigned(RANK)>;
using patch_layout_t = amr::ndt::patches::patch_layout<layout_t, HALO>;
using intergrid_t    = amr::ndt::intergrid_operator::linear_interpolator<patch_layout_t>;
using tree_t =
    amr::ndt::tree::ndtree<typename C::cell_t, patch_index_t, patch_layout_t, intergrid_t>;
using physics_t =
    amr::ndt::solver::physics_system<patch_index_t, patch_layout_t, C::lengths>;
using eq_t     = typename C::eq_t;
using solver_t = amr_solver<tree_t, physics_t, eq_t, RANK>;
using status_t = typename tree_t::refine_status_t;
using dir_t    = typename tree_t::patch_direction_t;
using nbr_pv_t = typename tree_t::neighbor_patch_index_variant_t;
using nbr_lv_t = typename tree_t::neighbor_linear_index_variant_t;

constexpr int         NVAR = eq_t::NVAR;
constexpr std::size_t FLAT = patch_layout_t::flat_size();
constexpr int         NDIR = 2 * RANK;
constexpr int         KF   = 1 << (RANK - 1);

std::uint64_t splitmix64(std::uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct writer
{
    std::ofstream out;

    void rec(std::string const& name, std::uint32_t dtype,
             std::vector<std::uint64_t> const& shape, void const* data, std::size_t bytes)
    {
        const std::uint32_t nl = static_cast<std::uint32_t>(name.size());
        out.write(reinterpret_cast<char const*>(&nl), 4);
        out.write(name.data(), nl);
        out.write(reinterpret_cast<char const*>(&dtype), 4);
        const std::uint32_t nd = static_cast<std::uint32_t>(shape.size());
        out.write(reinterpret_cast<char const*>(&nd), 4);
        out.write(reinterpret_cast<char const*>(shape.data()), 8 * nd);
        out.write(reinterpret_cast<char const*>(data), static_cast<std::streamsize>(bytes));
    }
};

template <std::size_t I>
using tag_t = std::tuple_element_t<I, typename eq_t::FieldTags>;

template <std::size_t... Is>
void gather_fields(tree_t& tree, std::vector<double>& buf, std::index_sequence<Is...>)
{
    const auto P = tree.size();
    buf.resize(static_cast<std::size_t>(NVAR) * P * FLAT);
    (
        [&]
        {
            for (std::size_t p = 0; p != P; ++p)
            {
                auto& patch = tree.template get_patch<tag_t<Is>>(p);
                for (std::size_t i = 0; i != FLAT; ++i)
                {
                    buf[(Is * P + p) * FLAT + i] = patch[i];
                }
            }
        }(),
        ...
    );
}

template <std::size_t... Is>
void probe_fill(tree_t& tree, std::index_sequence<Is...>)
{
    const auto P = tree.size();
    (
        [&]
        {
            for (std::size_t p = 0; p != P; ++p)
            {
                auto& patch = tree.template get_patch<tag_t<Is>>(p);
                for (std::size_t i = 0; i != FLAT; ++i)
                {
                    patch[i] = static_cast<double>(
                        ((static_cast<std::uint64_t>(Is) << 20) + p) * 4096ull + i
                    );
                }
            }
        }(),
        ...
    );
}

void dump(writer& w, std::string const& tag, tree_t& tree)
{
    const std::uint64_t        P = tree.size();
    std::vector<std::uint64_t> ids(P);
    std::vector<std::int8_t>   rel(P * NDIR, 0);
    std::vector<std::int32_t>  nbr(P * NDIR * KF, -1);
    std::vector<std::int8_t>   quad(P * NDIR * RANK, 0);
    for (std::uint64_t p = 0; p != P; ++p)
    {
        ids[p] = tree.get_node_index_at(p).id();
        for (auto d = dir_t::first(); d != dir_t::sentinel(); d.advance())
        {
            const auto di  = static_cast<std::size_t>(d.index());
            const auto lin = tree.neighbor_linear_index(tree.get_neighbor_at(p, d));
            std::visit(
                [&](auto const& n)
                {
                    using T = std::decay_t<decltype(n)>;
                    if constexpr (std::is_same_v<T, typename nbr_lv_t::same>)
                    {
                        rel[p * NDIR + di]        = 1;
                        nbr[(p * NDIR + di) * KF] = static_cast<std::int32_t>(n.id);
                    }
                    else if constexpr (std::is_same_v<T, typename nbr_lv_t::finer>)
                    {
                        rel[p * NDIR + di] = 2;
                        for (int k = 0; k != KF; ++k)
                            nbr[(p * NDIR + di) * KF + k] =
                                static_cast<std::int32_t>(n.ids[k]);
                    }
                    else if constexpr (std::is_same_v<T, typename nbr_lv_t::coarser>)
                    {
                        rel[p * NDIR + di]        = 3;
                        nbr[(p * NDIR + di) * KF] = static_cast<std::int32_t>(n.id);
                        for (int r = 0; r != RANK; ++r)
                            quad[(p * NDIR + di) * RANK + r] =
                                static_cast<std::int8_t>(n.contact_quadrant[r]);
                    }
                },
                lin.data
            );
        }
    }
    std::vector<double> data;
    gather_fields(tree, data, std::make_index_sequence<NVAR>{});
    w.rec(tag + "/ids", 4, { P }, ids.data(), 8 * P);
    w.rec(tag + "/rel", 3, { P, NDIR }, rel.data(), rel.size());
    w.rec(tag + "/nbr", 2, { P, NDIR, KF }, nbr.data(), 4 * nbr.size());
    w.rec(tag + "/quad", 3, { P, NDIR, RANK }, quad.data(), quad.size());
    w.rec(tag + "/data", 0, { NVAR, P, FLAT }, data.data(), 8 * data.size());
}

auto ic(std::array<double, RANK> const& x) -> amr::containers::static_vector<double, NVAR>
{
    amr::containers::static_vector<double, NVAR> prim;
#if CFG_EQ == 0
    // examples/fvm_solver_advection.e.cpp:57-63 (scaled by L)
    double r2 = 0;
    for (int d = 0; d != RANK; ++d) r2 += (x[d] - 0.2 * L) * (x[d] - 0.2 * L);
    prim[0] = std::exp(-r2 / (0.005 * L * L));
#else
    // benchmark/bench_fvm_solver_integration.b.cpp:146-178, 3D.b.cpp:149-186
    double r2 = 0;
    for (int d = 0; d != RANK; ++d) r2 += (x[d] - 0.5 * L) * (x[d] - 0.5 * L);
    const double pert = 10.0 * std::exp(-r2 / (0.01 * L * L));
    prim[0]           = 0.5 + pert * 0.2;
    for (int d = 0; d != RANK; ++d) prim[1 + d] = 0.0;
    prim[RANK + 1] = 1.0 + pert;
#endif
    return prim;
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        std::fprintf(stderr, "usage: %s script out [capacity] [gamma] [cfl]\n", argv[0]);
        return 2;
    }
    const std::size_t capacity = argc > 3 ? std::stoul(argv[3]) : 20000;
    const double      gamma    = argc > 4 ? std::stod(argv[4]) : 1.4;
    const double      cfl      = argc > 5 ? std::stod(argv[5]) : 0.3;

    solver_t solver(capacity, gamma, cfl);
    auto&    tree = solver.get_tree();
    writer   w{ std::ofstream(argv[2], std::ios::binary) };

    {
        const std::int64_t meta[8] = { RANK, (std::int64_t)S, (std::int64_t)HALO, DEPTH,
                                       CFG_EQ, NVAR, (std::int64_t)FLAT, 0 };
        w.rec("meta", 1, { 8 }, meta, sizeof(meta));
        const double fmeta[3] = { L, gamma, cfl };
        w.rec("fmeta", 0, { 3 }, fmeta, sizeof(fmeta));
    }

    std::ifstream       script(argv[1]);
    std::string         line;
    std::vector<double> dts;
#ifdef AMR_ENABLE_CUDA_AMR
    bool device_live = false; // the device holds the authoritative state (after I)
#endif
    while (std::getline(script, line))
    {
        std::istringstream is(line);
        std::string        op;
        if (!(is >> op) || op[0] == '#') continue;
        if (op == "A")
        {
            tree.reconstruct_tree([](patch_index_t const&) { return status_t::Refine; });
        }
        else if (op == "H")
        {
            std::uint64_t seed;
            unsigned      pr, pc;
            int           minl, maxl;
            is >> seed >> pr >> pc >> minl >> maxl;
            tree.reconstruct_tree(
                [=](patch_index_t const& id)
                {
                    const auto u   = splitmix64(id.id() ^ (seed * 0x100000001B3ull)) % 1000u;
                    const int  lvl = static_cast<int>(id.level());
                    if (u < pr && lvl < maxl) return status_t::Refine;
                    if (u >= pr && u < pr + pc && lvl > minl) return status_t::Coarsen;
                    return status_t::Stable;
                }
            );
        }
        else if (op == "B")
        {
            double r;
            int    minl, maxl;
            double c[3] = { 0, 0, 0 };
            is >> r >> minl >> maxl;
            for (int d = 0; d != RANK; ++d) is >> c[d];
            tree.reconstruct_tree(
                [=](patch_index_t const& id)
                {
                    const auto org = physics_t::patch_coord(id);
                    const auto sz  = physics_t::patch_sizes(id);
                    double     r2  = 0;
                    for (int d = 0; d != RANK; ++d)
                    {
                        const double x = (org[d] + 0.5 * sz[d]) / L - c[d];
                        r2 += x * x;
                    }
                    const int lvl = static_cast<int>(id.level());
                    if (r2 < r * r && lvl < maxl) return status_t::Refine;
                    if (r2 > 4 * r * r && lvl > minl) return status_t::Coarsen;
                    return status_t::Stable;
                }
            );
        }
        else if (op == "R" || op == "K")
        {
            std::vector<std::uint64_t> list;
            std::uint64_t              v;
            while (is >> v) list.push_back(v);
            const auto flag = op == "R" ? status_t::Refine : status_t::Coarsen;
            tree.reconstruct_tree(
                [&](patch_index_t const& id)
                {
                    for (auto const l : list)
                        if (l == id.id()) return flag;
                    return status_t::Stable;
                }
            );
        }
        else if (op == "X")
        {
#ifdef AMR_ENABLE_CUDA_AMR
            if (!device_live)
            {
                tree.sync_current_to_device();
                tree.halo_exchange_update();
                tree.sync_current_from_device();
            }
            else
#endif
                tree.halo_exchange_update();
        }
        else if (op == "P")
        {
            probe_fill(tree, std::make_index_sequence<NVAR>{});
        }
        else if (op == "I")
        {
            solver.initialize(ic);
#ifdef AMR_ENABLE_CUDA_AMR
            tree.sync_current_to_device();
            tree.build_patch_levels_on_device();
            device_live = true;
#endif
        }
        else if (op == "S")
        {
            int n;
            is >> n;
            for (int i = 0; i != n; ++i) dts.push_back(solver.advance());
        }
        else if (op == "T")
        {
            // timed leg for bench.py's reference arm: n x advance(), wall clock, the
            // reference's own metric (benchmark/bench_fvm_solver_integration.b.cpp:241-256)
            int n;
            is >> n;
            const auto        patches = tree.size();
            const auto        t0      = std::chrono::steady_clock::now();
            std::size_t       done    = 0;
            double            sum_dt  = 0;
#ifdef AMR_ENABLE_CUDA_AMR
            {
                // one batch of n steps, as the reference's benchmark loop issues them
                // (b.cpp:218-247); finish_advance_batch waits for the fence recorded after the
                // batch's last launch on the default stream
                std::size_t executed = 0;
                solver.advance_batch_async(static_cast<std::size_t>(n));
                sum_dt += solver.finish_advance_batch(&executed);
                done = executed;
            }
#else
            for (int i = 0; i != n; ++i)
            {
                sum_dt += solver.advance();
                ++done;
            }
#endif
            const std::chrono::duration<double> el = std::chrono::steady_clock::now() - t0;
            const double updates = static_cast<double>(done) * static_cast<double>(patches) *
                                   static_cast<double>(patch_layout_t::data_layout_t::flat_size());
            std::printf(
                "{\"patches\": %zu, \"cells\": %zu, \"steps\": %zu, \"seconds\": %.6f, "
                "\"updates_per_s\": %.6e, \"sum_dt\": %.17g}\n",
                patches, patches * patch_layout_t::data_layout_t::flat_size(), done, el.count(),
                updates / el.count(), sum_dt
            );
            std::fflush(stdout);
        }
        else if (op == "V")
        {
            // the reference's own VTK writer over the current tree -> vtk_output/<name>.vtk (cwd)
            std::string name;
            is >> name;
            amr::ndt::print::vtk_print<physics_t> printer(name);
            printer.print(tree, ".vtk");
        }
        else if (op == "D")
        {
            std::string tag;
            is >> tag;
#ifdef AMR_ENABLE_CUDA_AMR
            if (device_live) tree.sync_current_from_device();
#endif
            dump(w, tag, tree);
            w.rec(tag + "/dts", 0, { dts.size() }, dts.data(), 8 * dts.size());
        }
        else
        {
            std::fprintf(stderr, "unknown op '%s'\n", op.c_str());
            return 3;
        }
    }
    return 0;
}
