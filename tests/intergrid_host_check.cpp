// Host check of amr::ndt::intergrid_operator::linear_interpolator (this repo's include/): injection, block
// injection and child mean on plain arrays; prints the results for tests/test_intergrid_operator.py.
#include "ndtree/intergrid_operator.hpp"
#include "ndtree/patch_layout.hpp"
#include "ndtree/patch_utils.hpp"
#include "containers/static_layout.hpp"
#include "containers/static_shape.hpp"

#include <array>
#include <cstdio>

int main()
{
    using shape_t  = amr::containers::static_shape<4, 4>;
    using layout_t = amr::containers::static_layout<shape_t>;
    using patch_layout_t = amr::ndt::patches::patch_layout<layout_t, 1>;
    using op      = amr::ndt::intergrid_operator::linear_interpolator<patch_layout_t>;
    using index_t = op::index_t;

    std::array<double, 8> coarse{ 1.5, -2.0, 3.25, 0.0, 7.0, 8.0, 9.0, 10.0 };
    std::array<double, 8> fine{};
    op::interpolation(fine, index_t{ 5 }, index_t{ 2 }, coarse, index_t{ 1 });
    op::template interpolation<4>(fine, std::array<index_t, 4>{ 0, 1, 2, 3 }, coarse, index_t{ 2 });
    std::array<double, 2> back{};
    op::template restriction<4>(back, index_t{ 0 }, coarse, std::array<index_t, 4>{ 4, 5, 6, 7 });
    op::template restriction<4>(back, index_t{ 1 }, fine, std::array<index_t, 4>{ 0, 1, 2, 5 });
    // children of the coarse cell whose lowest fine corner is padded cell (1, 1) of a 4 x 4 / halo 1 patch (6 x 6
    // padded): (1,1) (1,2) (2,1) (2,2) -> 7 8 13 14
    constexpr auto kids = amr::ndt::utils::patches::detail::hypercube_offset<patch_layout_t, 2>(index_t{ 7 });
    static_assert(kids.size() == 4 && kids[0] == 7 && kids[1] == 8 && kids[2] == 13 && kids[3] == 14);
    for (double v : fine) std::printf("%.17g ", v);
    std::printf("| %.17g %.17g\n", back[0], back[1]);
    return 0;
}
