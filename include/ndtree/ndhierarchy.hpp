// The reference's alternative hierarchical-prefix index (include/ndtree/ndhierarchy.hpp) is unused by
// the FVM drivers (out of scope); this header only keeps their #include lines valid.
#ifndef AMRB_NDTREE_NDHIERARCHY_HPP
#define AMRB_NDTREE_NDHIERARCHY_HPP
#endif
