// amr::ndt::intergrid_operator::linear_interpolator — the only transfer operator the device path
// implements: piecewise-constant injection coarse->fine and 2^rank mean fine->coarse
// (include/ndtree/intergrid_operator.hpp:18-106 of the reference).  The arithmetic itself lives in
// the CUDA kernels (halo_source, plan_kernel); this type selects it.
#ifndef AMRB_NDTREE_INTERGRID_OPERATOR_HPP
#define AMRB_NDTREE_INTERGRID_OPERATOR_HPP
#include "patch_layout.hpp"

namespace amr::ndt::intergrid_operator
{
template <typename Patch_Layout>
struct linear_interpolator
{
    using patch_layout_t = Patch_Layout;
    using index_t        = typename Patch_Layout::index_t;
    static constexpr bool device_native = true;
};
} // namespace amr::ndt::intergrid_operator
#endif
