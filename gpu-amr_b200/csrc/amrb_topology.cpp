// Host topology of the Morton-ordered leaf patch store (C-ABI section 7 of gpuamr_b200.h).
//
// Set-based formulation: the tree is nothing but the ascending array of leaf ids
// (id = morton(x,y[,z]) << 6 | level, include/morton/morton_id.hpp:21-229 / 232-450 of the
// reference).  Neighbor relations, contact quadrants and finer-neighbor orderings are derived
// from the leaf set by key lookup instead of being maintained incrementally through
// fragment/recombine as the reference does (ndtree/neighbor.hpp:419-572,
// ndtree.hpp:942-1125); both give the same tables because linear index == rank of the leaf id
// and the relations depend only on the set of leaves.  The refine/coarsen selection rules
// (eligibility, 2:1 ripple, coarsening veto) follow ndtree.hpp:886-940 and 1127-1240.
#include "../../include/gpuamr_b200.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace amrb
{
extern thread_local std::string g_error;
}

namespace
{

amrb_status tfail(amrb_status code, const char* what)
{
    amrb::g_error = what;
    return code;
}

constexpr int kLevelBits = 6;

// open-addressing id -> linear index map
struct IdMap
{
    std::vector<uint64_t> keys;
    std::vector<int32_t>  vals;
    uint64_t              mask = 0;
    static constexpr uint64_t kEmpty = ~0ull;

    static uint64_t mix(uint64_t x)
    {
        x ^= x >> 31;
        x *= 0x9E3779B97F4A7C15ull;
        x ^= x >> 29;
        return x;
    }
    void build(const std::vector<uint64_t>& ids)
    {
        size_t cap = 16;
        while (cap < ids.size() * 2 + 2) cap <<= 1;
        keys.assign(cap, kEmpty);
        vals.assign(cap, -1);
        mask = cap - 1;
        for (size_t i = 0; i < ids.size(); ++i)
        {
            uint64_t s = mix(ids[i]) & mask;
            while (keys[s] != kEmpty) s = (s + 1) & mask;
            keys[s] = ids[i];
            vals[s] = (int32_t)i;
        }
    }
    int32_t find(uint64_t id) const
    {
        uint64_t s = mix(id) & mask;
        while (keys[s] != kEmpty)
        {
            if (keys[s] == id) return vals[s];
            s = (s + 1) & mask;
        }
        return -1;
    }
};

struct Nbr
{
    int8_t  rel;
    int32_t idx[4];
    int8_t  quad[3];
};

} // namespace

struct amrb_tree
{
    int                   rank = 2, depth = 1;
    std::vector<uint64_t> ids;
    // id -> linear index.  The hash map costs a full pass over the leaves to build (0.1 s at 2e6
    // leaves), so it is only built when a call is going to do many lookups (table build, a pass
    // that splits a large share of the leaves); otherwise a lookup is a binary search in the
    // ascending id array.
    mutable IdMap         map;
    mutable bool          map_valid = false;

    void invalidate_map() { map_valid = false; }
    void ensure_map() const
    {
        if (!map_valid)
        {
            map.build(ids);
            map_valid = true;
        }
    }
    int32_t find(uint64_t id) const
    {
        if (map_valid) return map.find(id);
        const auto it = std::lower_bound(ids.begin(), ids.end(), id);
        return (it != ids.end() && *it == id) ? (int32_t)(it - ids.begin()) : -1;
    }
    std::vector<int8_t>   plan_kind, plan_child;
    std::vector<int32_t>  plan_src;

    uint32_t span() const { return 1u << depth; }

    uint64_t encode(const uint32_t* c, int level) const { return amrb_morton_encode(rank, c, level); }
    void     decode(uint64_t id, uint32_t* c, int& level) const { amrb_morton_decode(rank, id, c, &level); }

    // neighbor of leaf `id` across direction d, derived from the leaf set
    Nbr neighbor(uint64_t id, int d) const
    {
        Nbr out{};
        out.rel = AMRB_REL_NONE;
        for (int k = 0; k < 4; ++k) out.idx[k] = -1;
        uint32_t c[3] = { 0, 0, 0 };
        int      lvl  = 0;
        decode(id, c, lvl);
        const int      dim = d >> 1, pos = d & 1;
        const int      ax  = rank - 1 - dim; // layout dim k <-> morton axis rank-1-k
        const uint32_t e   = 1u << (depth - lvl);
        uint32_t       n[3] = { c[0], c[1], c[2] };
        n[ax] = (pos ? c[ax] + e : c[ax] + span() - e) & (span() - 1); // periodic (ndtree.hpp:412-421)

        int32_t li = find(encode(n, lvl));
        if (li >= 0)
        {
            out.rel    = AMRB_REL_SAME;
            out.idx[0] = li;
            return out;
        }
        if (lvl > 0)
        {
            const uint32_t E = e << 1;
            uint32_t       cn[3];
            for (int a = 0; a < 3; ++a) cn[a] = n[a] & ~(E - 1);
            li = find(encode(cn, lvl - 1));
            if (li >= 0)
            {
                out.rel    = AMRB_REL_COARSER;
                out.idx[0] = li;
                // contact quadrant (ndtree/neighbor.hpp:340-364): normal component 0 for a
                // positive direction, 1 for a negative one; tangential = my half inside the
                // coarse neighbor's extent
                for (int k = 0; k < rank; ++k)
                    out.quad[k] = (k == dim) ? (int8_t)(pos ? 0 : 1)
                                             : (int8_t)((c[rank - 1 - k] / e) & 1u);
                return out;
            }
        }
        if (lvl < depth)
        {
            const uint32_t hh = e >> 1;
            const int      kf = 1 << (rank - 1);
            bool           ok = true;
            for (int j = 0; j < kf && ok; ++j)
            {
                // finer-neighbor order: non-normal layout dims ascending, lowest = bit 0
                // (ndtree/neighbor.hpp:316-337)
                uint32_t cc[3] = { n[0], n[1], n[2] };
                cc[ax] += pos ? 0 : hh;
                int bit = 0;
                for (int k = 0; k < rank; ++k)
                {
                    if (k == dim) continue;
                    if ((j >> bit) & 1) cc[rank - 1 - k] += hh;
                    ++bit;
                }
                li = find(encode(cc, lvl + 1));
                if (li < 0)
                    ok = false;
                else
                    out.idx[j] = li;
            }
            if (ok)
            {
                out.rel = AMRB_REL_FINER;
                return out;
            }
            for (int k = 0; k < 4; ++k) out.idx[k] = -1;
        }
        return out;
    }
};

extern "C" {

// bit interleave by shift-and-mask (21 bits per coordinate): x in bit 0 of every group
static inline uint64_t spread2(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
static inline uint32_t gather2(uint64_t v)
{
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
    v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
    v = (v | (v >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)v;
}
static inline uint64_t spread3(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
static inline uint32_t gather3(uint64_t v)
{
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10c30c30c30c30c3ull;
    v = (v | (v >> 4)) & 0x100f00f00f00f00full;
    v = (v | (v >> 8)) & 0x1f0000ff0000ffull;
    v = (v | (v >> 16)) & 0x1f00000000ffffull;
    v = (v | (v >> 32)) & 0x1fffffull;
    return (uint32_t)v;
}

uint64_t amrb_morton_encode(int rank, const uint32_t* c, int level)
{
    const uint64_t m = (rank == 2) ? (spread2(c[0]) | (spread2(c[1]) << 1))
                                   : (spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2));
    return (m << kLevelBits) | (uint64_t)level;
}

void amrb_morton_decode(int rank, uint64_t id, uint32_t* c, int* level)
{
    if (level) *level = (int)(id & ((1u << kLevelBits) - 1));
    const uint64_t m = id >> kLevelBits;
    if (rank == 2)
    {
        c[0] = gather2(m);
        c[1] = gather2(m >> 1);
    }
    else
    {
        c[0] = gather3(m);
        c[1] = gather3(m >> 1);
        c[2] = gather3(m >> 2);
    }
}

amrb_status amrb_tree_create(int rank, int depth, amrb_tree** out)
{
    if (!out) return tfail(AMRB_ERR_ARGUMENT, "null out");
    if (rank != 2 && rank != 3) return tfail(AMRB_ERR_ARGUMENT, "rank must be 2 or 3");
    if (depth < 1 || depth > 19) return tfail(AMRB_ERR_ARGUMENT, "depth out of range");
    amrb_tree* t = new amrb_tree();
    t->rank      = rank;
    t->depth     = depth;
    t->ids.assign(1, 0ull); // the periodic root (ndtree.hpp:411-421)
    t->invalidate_map();
    *out = t;
    return AMRB_OK;
}

amrb_status amrb_tree_destroy(amrb_tree* t)
{
    delete t;
    return AMRB_OK;
}

size_t          amrb_tree_size(const amrb_tree* t) { return t ? t->ids.size() : 0; }
const uint64_t* amrb_tree_ids(const amrb_tree* t) { return t ? t->ids.data() : nullptr; }

amrb_status amrb_tree_tables(const amrb_tree* t, int32_t* levels, int8_t* rel, int32_t* nbr,
                             int8_t* quad)
{
    if (!t || !levels || !rel || !nbr || !quad) return tfail(AMRB_ERR_ARGUMENT, "null argument");
    const int    R = t->rank, ND = 2 * R, KF = 1 << (R - 1);
    const size_t n = t->ids.size();
    t->ensure_map(); // 2R lookups per leaf
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i)
    {
        levels[i] = (int32_t)(t->ids[i] & 63u);
        for (int d = 0; d < ND; ++d)
        {
            const Nbr    nb = t->neighbor(t->ids[i], d);
            const size_t o  = (size_t)i * ND + d;
            rel[o]          = nb.rel;
            for (int k = 0; k < KF; ++k) nbr[o * KF + k] = nb.idx[k];
            for (int k = 0; k < R; ++k) quad[o * R + k] = (nb.rel == AMRB_REL_COARSER) ? nb.quad[k] : 0;
        }
    }
    return AMRB_OK;
}

amrb_status amrb_tree_reconstruct(amrb_tree* t, const int8_t* flags, size_t capacity, int* changed)
{
    if (!t || !flags) return tfail(AMRB_ERR_ARGUMENT, "null argument");
    if (changed) *changed = 0;
    const int    R = t->rank, ND = 2 * R, FAN = 1 << R, KF = 1 << (R - 1);
    const size_t n = t->ids.size();

    // ---- eligibility (ndtree.hpp:904-940, 1888-1917)
    std::vector<uint8_t> refine(n, 0);      // leaf is split
    std::vector<int32_t> work;              // ripple work list (leaf indices)
    for (size_t i = 0; i < n; ++i)
        if (flags[i] == AMRB_REFINE && (int)(t->ids[i] & 63u) < t->depth)
        {
            refine[i] = 1;
            work.push_back((int32_t)i);
        }
    // a parent is recombined when all of its 2^R children are leaves flagged Coarsen; children
    // of one parent are contiguous in Morton order, first child = aligned anchor
    std::vector<int32_t> coarsen_first;
    for (size_t i = 0; i + FAN <= n; ++i)
    {
        if (flags[i] != AMRB_COARSEN) continue;
        uint32_t c[3] = { 0, 0, 0 };
        int      lvl  = 0;
        t->decode(t->ids[i], c, lvl);
        if (lvl == 0) continue;
        const uint32_t E = 1u << (t->depth - lvl + 1);
        bool           first = true;
        for (int a = 0; a < R; ++a) first = first && ((c[a] & (E - 1)) == 0);
        if (!first) continue;
        bool ok = true;
        for (int j = 0; j < FAN && ok; ++j)
        {
            uint32_t cc[3] = { c[0], c[1], c[2] };
            for (int a = 0; a < R; ++a)
                if ((j >> a) & 1) cc[a] += E >> 1;
            ok = t->ids[i + j] == t->encode(cc, lvl) && flags[i + j] == AMRB_COARSEN;
        }
        if (ok) coarsen_first.push_back((int32_t)i);
    }
    if (work.empty() && coarsen_first.empty()) return AMRB_OK;
    // ripple + veto look up 2R neighbors per splitting leaf / per child of a merging family
    if ((work.size() + coarsen_first.size() * FAN) * (size_t)ND > n / 4) t->ensure_map();

    // ---- 2:1 ripple: a coarser neighbor of a splitting leaf splits too (ndtree.hpp:1127-1166)
    for (size_t w = 0; w < work.size(); ++w)
    {
        const uint64_t id = t->ids[work[w]];
        for (int d = 0; d < ND; ++d)
        {
            const Nbr nb = t->neighbor(id, d);
            if (nb.rel == AMRB_REL_COARSER && !refine[nb.idx[0]])
            {
                refine[nb.idx[0]] = 1;
                work.push_back(nb.idx[0]);
            }
        }
    }
    // ---- coarsening veto (ndtree.hpp:1173-1239): a boundary child whose outward neighbor is
    // finer, or same-level and about to split, blocks the recombination
    std::vector<uint8_t> merge_first(n, 0);
    for (int32_t first : coarsen_first)
    {
        bool veto = false;
        for (int j = 0; j < FAN && !veto; ++j)
        {
            const uint64_t id = t->ids[first + j];
            for (int d = 0; d < ND && !veto; ++d)
            {
                const int dim = d >> 1, pos = d & 1, ax = R - 1 - dim;
                if (((j >> ax) & 1) != pos) continue; // inward direction: sibling
                const Nbr nb = t->neighbor(id, d);
                if (nb.rel == AMRB_REL_FINER) veto = true;
                if (nb.rel == AMRB_REL_SAME && refine[nb.idx[0]]) veto = true;
            }
        }
        if (!veto) merge_first[first] = 1;
    }
    (void)KF;

    // ---- new leaf set in ascending id order + transfer plan
    std::vector<uint64_t> ids;
    std::vector<int8_t>   kind, child;
    std::vector<int32_t>  src;
    ids.reserve(n + work.size() * (FAN - 1));
    bool any = false;
    for (size_t i = 0; i < n;)
    {
        if (merge_first[i])
        {
            uint32_t c[3] = { 0, 0, 0 };
            int      lvl  = 0;
            t->decode(t->ids[i], c, lvl);
            ids.push_back(t->encode(c, lvl - 1));
            kind.push_back(2);
            src.push_back((int32_t)i);
            child.push_back(0);
            i += FAN;
            any = true;
            continue;
        }
        if (refine[i])
        {
            uint32_t c[3] = { 0, 0, 0 };
            int      lvl  = 0;
            t->decode(t->ids[i], c, lvl);
            const uint32_t hh = 1u << (t->depth - lvl - 1);
            for (int j = 0; j < FAN; ++j)
            {
                // child_of(parent, j): coordinate axis a advances by bit a of j
                // (morton_id.hpp:142-148, 64-83); ascending j == ascending id
                uint32_t cc[3] = { c[0], c[1], c[2] };
                for (int a = 0; a < R; ++a)
                    if ((j >> a) & 1) cc[a] += hh;
                ids.push_back(t->encode(cc, lvl + 1));
                kind.push_back(1);
                src.push_back((int32_t)i);
                child.push_back((int8_t)j);
            }
            ++i;
            any = true;
            continue;
        }
        ids.push_back(t->ids[i]);
        kind.push_back(0);
        src.push_back((int32_t)i);
        child.push_back(0);
        ++i;
    }
    if (!any) return AMRB_OK;
    if (capacity && ids.size() > capacity)
        return tfail(AMRB_ERR_CAPACITY, "reconstruct would exceed the patch capacity");
    t->ids.swap(ids);
    t->invalidate_map();
    t->plan_kind.swap(kind);
    t->plan_src.swap(src);
    t->plan_child.swap(child);
    if (changed) *changed = 1;
    return AMRB_OK;
}

// adopt a leaf set computed elsewhere (amrb_pool_reconstruct_device): ascending ids
amrb_status amrb_tree_assign(amrb_tree* t, const uint64_t* ids, size_t n)
{
    if (!t || !ids || n == 0) return tfail(AMRB_ERR_ARGUMENT, "null argument");
    for (size_t i = 1; i < n; ++i)
        if (!(ids[i - 1] < ids[i])) return tfail(AMRB_ERR_ARGUMENT, "leaf ids must be strictly ascending");
    t->ids.assign(ids, ids + n);
    t->invalidate_map();
    t->plan_kind.clear();
    t->plan_src.clear();
    t->plan_child.clear();
    return AMRB_OK;
}

size_t amrb_tree_plan_size(const amrb_tree* t) { return t ? t->plan_kind.size() : 0; }

amrb_status amrb_tree_plan(const amrb_tree* t, int8_t* kind, int32_t* src, int8_t* child)
{
    if (!t || !kind || !src || !child) return tfail(AMRB_ERR_ARGUMENT, "null argument");
    const size_t n = t->plan_kind.size();
    std::memcpy(kind, t->plan_kind.data(), n);
    std::memcpy(src, t->plan_src.data(), n * sizeof(int32_t));
    std::memcpy(child, t->plan_child.data(), n);
    return AMRB_OK;
}

} // extern "C"
