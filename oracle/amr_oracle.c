/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See amr_oracle.h for the contract.
 *
 * Plain-C restatement of the reference algorithm; every block cites the reference
 * file:line it follows.  Parity status: PINNED against reference dumps
 * (tests/golden/, tests/test_oracle_golden.py).
 *
 * Conventions (all verified against the dumps):
 *   - layout dim k (0 = slowest) <-> Morton axis R-1-k <-> bit R-1-k of a child number;
 *   - direction index d: dim = d/2, positive = d&1 (ndtree/neighbor.hpp:103-255);
 *   - finer-neighbor slot k = sum over non-normal layout dims ascending, lowest dim = bit 0.
 */
#include "amr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#    include <omp.h>
#endif

#define MAXR 3
#define MAXD 6
#define MAXK 4

typedef struct
{
    int8_t   rel;
    int8_t   quad[MAXR];
    uint64_t id[MAXK];
} nbr_t;

struct orc_tree
{
    int    rank, depth, halo, nvar, ndir, kf, fan;
    int    size[MAXR], psize[MAXR], pstride[MAXR];
    size_t flat, dflat, n, cap;
    uint64_t* ids;
    int8_t*   status;
    nbr_t*    nbrs; /* [cap][MAXD] */
    double**  cur;
    double**  nxt;
    /* id -> linear index, open addressing (ndtree.hpp:277 m_index_map) */
    uint64_t* hkey;
    int64_t*  hval; /* -1 empty, -2 tombstone */
    size_t    hmask;
    /* work lists (ndtree.hpp:2155-2156) */
    uint64_t* to_refine;
    size_t    n_refine;
    uint64_t* to_coarsen;
    size_t    n_coarsen;
};

/* ------------------------------------------------------------------ morton */
/* morton/morton_id.hpp:118-140 (2D), 3D analogue: id = interleave(coords) << 6 | level */
uint64_t orc_morton_encode(int rank, const uint32_t* c, int level)
{
    uint64_t m = 0;
    for (int b = 0; b < 20; ++b)
        for (int a = 0; a < rank; ++a)
            m |= (uint64_t)((c[a] >> b) & 1u) << (rank * b + a);
    return (m << 6) | (uint64_t)(level & 0x3F);
}

void orc_morton_decode(int rank, uint64_t id, uint32_t* c, int* level)
{
    *level     = (int)(id & 0x3F);
    uint64_t m = id >> 6;
    for (int a = 0; a < rank; ++a) c[a] = 0;
    for (int b = 0; b < 20; ++b)
        for (int a = 0; a < rank; ++a)
            c[a] |= (uint32_t)((m >> (rank * b + a)) & 1u) << b;
}

static int id_level(uint64_t id)
{
    return (int)(id & 0x3F);
}

/* morton_id.hpp:142-148 child_of = offset(id+1, off); :64-83 offset */
static uint64_t child_of(const orc_tree* t, uint64_t parent, int off)
{
    uint32_t c[MAXR];
    int      lvl;
    orc_morton_decode(t->rank, parent + 1, c, &lvl);
    const uint32_t delta = 1u << (t->depth - lvl);
    for (int a = 0; a < t->rank; ++a) c[a] += (uint32_t)((off >> a) & 1) * delta;
    return orc_morton_encode(t->rank, c, lvl);
}

/* morton_id.hpp:103-116 */
static uint64_t parent_of(const orc_tree* t, uint64_t id)
{
    uint32_t c[MAXR];
    int      lvl;
    orc_morton_decode(t->rank, id, c, &lvl);
    const uint32_t off = 1u << (t->depth - lvl);
    for (int a = 0; a < t->rank; ++a) c[a] &= ~off;
    return orc_morton_encode(t->rank, c, lvl - 1);
}

/* ------------------------------------------------------------------ hash map */
static size_t hslot(const orc_tree* t, uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return (size_t)k & t->hmask;
}

static int64_t map_find(const orc_tree* t, uint64_t id)
{
    for (size_t s = hslot(t, id);; s = (s + 1) & t->hmask)
    {
        if (t->hval[s] == -1) return -1;
        if (t->hval[s] >= 0 && t->hkey[s] == id) return t->hval[s];
    }
}

static void map_set(orc_tree* t, uint64_t id, int64_t v)
{
    size_t  s;
    int64_t tomb = -1;
    for (s = hslot(t, id);; s = (s + 1) & t->hmask)
    {
        if (t->hval[s] == -1) break;
        if (t->hval[s] == -2)
        {
            if (tomb < 0) tomb = (int64_t)s;
            continue;
        }
        if (t->hkey[s] == id)
        {
            t->hval[s] = v;
            return;
        }
    }
    if (tomb >= 0) s = (size_t)tomb;
    t->hkey[s] = id;
    t->hval[s] = v;
}

static void map_erase(orc_tree* t, uint64_t id)
{
    for (size_t s = hslot(t, id);; s = (s + 1) & t->hmask)
    {
        if (t->hval[s] == -1) return;
        if (t->hval[s] >= 0 && t->hkey[s] == id)
        {
            t->hval[s] = -2;
            return;
        }
    }
}

static void map_clear(orc_tree* t)
{
    for (size_t s = 0; s <= t->hmask; ++s) t->hval[s] = -1;
}

/* ------------------------------------------------------------------ lifecycle */
/* ndtree.hpp:1304-1316 append (capacity is checked here; the reference does not, N7) */
static int append(orc_tree* t, uint64_t id, const nbr_t* nb)
{
    if (t->n >= t->cap) return -1;
    t->ids[t->n]    = id;
    t->status[t->n] = ORC_STABLE;
    memcpy(&t->nbrs[t->n * MAXD], nb, sizeof(nbr_t) * MAXD);
    map_set(t, id, (int64_t)t->n);
    ++t->n;
    return 0;
}

orc_tree* orc_tree_create(int rank, int depth, const int* size, int halo, int nvar, size_t cap)
{
    if (rank < 2 || rank > 3 || nvar < 1) return NULL;
    orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
    t->rank     = rank;
    t->depth    = depth;
    t->halo     = halo;
    t->nvar     = nvar;
    t->ndir     = 2 * rank;
    t->kf       = 1 << (rank - 1);
    t->fan      = 1 << rank;
    t->flat     = 1;
    t->dflat    = 1;
    for (int k = 0; k < rank; ++k)
    {
        t->size[k]  = size[k];
        t->psize[k] = size[k] + 2 * halo;
        t->flat *= (size_t)t->psize[k];
        t->dflat *= (size_t)size[k];
    }
    /* containers/static_layout.hpp:28-37 row-major strides, last dim fastest */
    t->pstride[rank - 1] = 1;
    for (int k = rank - 1; k-- > 0;) t->pstride[k] = t->pstride[k + 1] * t->psize[k + 1];
    t->cap    = cap;
    t->ids    = (uint64_t*)calloc(cap, sizeof(uint64_t));
    t->status = (int8_t*)calloc(cap, 1);
    t->nbrs   = (nbr_t*)calloc(cap * MAXD, sizeof(nbr_t));
    t->cur    = (double**)calloc((size_t)nvar, sizeof(double*));
    t->nxt    = (double**)calloc((size_t)nvar, sizeof(double*));
    for (int f = 0; f < nvar; ++f)
    {
        /* zero-initialised (N5/N7: the reference leaves malloc garbage in never-written cells) */
        t->cur[f] = (double*)calloc(cap * t->flat, sizeof(double));
        t->nxt[f] = (double*)calloc(cap * t->flat, sizeof(double));
    }
    size_t h = 64;
    while (h < 4 * cap) h <<= 1;
    t->hmask = h - 1;
    t->hkey  = (uint64_t*)calloc(h, sizeof(uint64_t));
    t->hval  = (int64_t*)malloc(h * sizeof(int64_t));
    map_clear(t);
    t->to_refine  = (uint64_t*)calloc(cap, sizeof(uint64_t));
    t->to_coarsen = (uint64_t*)calloc(cap, sizeof(uint64_t));
    /* ndtree.hpp:411-421 periodic root: same{root} in every direction */
    nbr_t root[MAXD];
    memset(root, 0, sizeof(root));
    for (int d = 0; d < t->ndir; ++d)
    {
        root[d].rel   = ORC_SAME;
        root[d].id[0] = 0;
    }
    append(t, 0, root);
    return t;
}

void orc_tree_destroy(orc_tree* t)
{
    if (!t) return;
    for (int f = 0; f < t->nvar; ++f)
    {
        free(t->cur[f]);
        free(t->nxt[f]);
    }
    free(t->cur);
    free(t->nxt);
    free(t->ids);
    free(t->status);
    free(t->nbrs);
    free(t->hkey);
    free(t->hval);
    free(t->to_refine);
    free(t->to_coarsen);
    free(t);
}

size_t orc_tree_size(const orc_tree* t)
{
    return t->n;
}
size_t orc_tree_flat_size(const orc_tree* t)
{
    return t->flat;
}
const uint64_t* orc_tree_ids(const orc_tree* t)
{
    return t->ids;
}
double* orc_tree_field(orc_tree* t, int f)
{
    return t->cur[f];
}
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ neighbor algebra */
/* child number i -> hypercube multi-index, row-major over (2,..,2): coords[k] = bit R-1-k */
static void child_coords(const orc_tree* t, int i, int* c)
{
    for (int k = 0; k < t->rank; ++k) c[k] = (i >> (t->rank - 1 - k)) & 1;
}
static int child_number(const orc_tree* t, const int* c)
{
    int i = 0;
    for (int k = 0; k < t->rank; ++k) i |= c[k] << (t->rank - 1 - k);
    return i;
}

/* neighbor.hpp:316-337 compute_fine_boundary_linear_index */
static int fine_boundary_index(const orc_tree* t, const int* c, int dim)
{
    int idx = 0, mul = 1;
    for (int k = 0; k < t->rank; ++k)
    {
        if (k == dim) continue;
        idx += c[k] * mul;
        mul *= 2;
    }
    return idx;
}

/* neighbor.hpp:397-417 compute_boundary_children */
static void boundary_children(const orc_tree* t, int d, int* out)
{
    const int dim = d / 2, pos = d & 1;
    for (int i = 0; i < t->fan; ++i)
    {
        int c[MAXR];
        child_coords(t, i, c);
        if (c[dim] != (pos ? 1 : 0)) continue; /* neighbor.hpp:291-313 relation map */
        out[fine_boundary_index(t, c, dim)] = i;
    }
}

/* neighbor.hpp:340-364 compute_contact_quadrant */
static void contact_quadrant(const orc_tree* t, int idx, int d, int8_t* q)
{
    const int dim = d / 2, pos = d & 1;
    for (int k = 0; k < t->rank; ++k)
    {
        if (k == dim)
            q[k] = pos ? 0 : 1;
        else
        {
            q[k] = (int8_t)(idx % 2);
            idx /= 2;
        }
    }
}

/* neighbor.hpp:419-493 compute_child_neighbors */
static void child_neighbors(const orc_tree* t, uint64_t parent, const nbr_t* pn, int child,
                            nbr_t* out)
{
    int c[MAXR];
    child_coords(t, child, c);
    memset(out, 0, sizeof(nbr_t) * MAXD);
    for (int d = 0; d < t->ndir; ++d)
    {
        const int dim = d / 2, pos = d & 1;
        const int sibling = pos ? (c[dim] != 1) : (c[dim] != 0);
        if (sibling || parent == 0)
        {
            /* neighbor.hpp:383-395 get_sibling_offset (periodic at the root) */
            int s[MAXR];
            memcpy(s, c, sizeof(s));
            s[dim]       = (s[dim] + 1) % 2;
            out[d].rel   = ORC_SAME;
            out[d].id[0] = child_of(t, parent, child_number(t, s));
            continue;
        }
        switch (pn[d].rel)
        {
            case ORC_SAME: /* :454-464, quadrant = compute_neighbor_coarse_block_coords :366-379 */
                out[d].rel   = ORC_COARSER;
                out[d].id[0] = pn[d].id[0];
                for (int k = 0; k < t->rank; ++k) out[d].quad[k] = (int8_t)c[k];
                out[d].quad[dim] = pos ? 0 : 1;
                break;
            case ORC_FINER: /* :465-475 */
                out[d].rel   = ORC_SAME;
                out[d].id[0] = pn[d].id[fine_boundary_index(t, c, dim)];
                break;
            default: out[d].rel = ORC_NONE; break;
        }
    }
}

/* neighbor.hpp:495-572 compute_parent_neighbors */
static void parent_neighbors(const orc_tree* t, const nbr_t* cn /* [fan][MAXD] */, nbr_t* out)
{
    memset(out, 0, sizeof(nbr_t) * MAXD);
    for (int d = 0; d < t->ndir; ++d)
    {
        int bc[MAXK];
        boundary_children(t, d, bc);
        const nbr_t* first = &cn[bc[0] * MAXD + d];
        switch (first->rel)
        {
            case ORC_SAME:
                out[d].rel = ORC_FINER;
                for (int i = 0; i < t->kf; ++i)
                {
                    const nbr_t* c = &cn[bc[i] * MAXD + d];
                    out[d].id[i]   = (c->rel == ORC_SAME) ? c->id[0] : 0;
                }
                break;
            case ORC_COARSER:
                out[d].rel   = ORC_SAME;
                out[d].id[0] = first->id[0];
                break;
            default: out[d].rel = ORC_NONE; break;
        }
    }
}

/* ------------------------------------------------------------------ patch transfer */
static size_t lin(const orc_tree* t, const int* idx)
{
    size_t l = 0;
    for (int k = 0; k < t->rank; ++k) l += (size_t)idx[k] * (size_t)t->pstride[k];
    return l;
}

/* ndtree/patch_utils.hpp:203-234 hypercube_offset<L,2>: {0,1}^R, last dim fastest */
static void hypercube(const orc_tree* t, size_t base, size_t* out)
{
    for (int n = 0; n < t->fan; ++n)
    {
        size_t o = base;
        for (int k = 0; k < t->rank; ++k)
            o += (size_t)((n >> (t->rank - 1 - k)) & 1) * (size_t)t->pstride[k];
        out[n] = o;
    }
}

/* ndtree.hpp:1463-1497 patch_mapping_impl, :1499-1526 restrict, :1528-1555 interpolate */
static void transfer(orc_tree* t, size_t coarse, size_t fine_start, int restrict_)
{
    int idx[MAXR];
    for (int k = 0; k < t->rank; ++k) idx[k] = t->halo;
    for (;;)
    {
        int fp[MAXR], base[MAXR];
        for (int k = 0; k < t->rank; ++k)
        {
            const int section = t->size[k] / 2;
            fp[k]             = (idx[k] - t->halo) / section;
            base[k]           = ((idx[k] - t->halo) % section) * 2 + t->halo;
        }
        const size_t fine_patch = fine_start + (size_t)child_number(t, fp);
        size_t       offs[8];
        hypercube(t, lin(t, base), offs);
        const size_t ci = lin(t, idx);
        for (int f = 0; f < t->nvar; ++f)
        {
            double* c  = t->cur[f] + coarse * t->flat;
            double* fn = t->cur[f] + fine_patch * t->flat;
            if (restrict_)
            {
                /* intergrid_operator.hpp:92-106: sum in enumeration order, then / N */
                double sum = 0.0;
                for (int n = 0; n < t->fan; ++n) sum += fn[offs[n]];
                c[ci] = sum / (double)t->fan;
            }
            else
            {
                for (int n = 0; n < t->fan; ++n) fn[offs[n]] = c[ci]; /* :41-76 injection */
            }
        }
        int k = t->rank - 1;
        for (; k >= 0; --k)
        {
            if (++idx[k] < t->halo + t->size[k]) break;
            idx[k] = t->halo;
        }
        if (k < 0) break;
    }
}

/* ------------------------------------------------------------------ reconstruct */
/* ndtree.hpp:942-1037 */
static void symmetry_after_refinement(orc_tree* t, uint64_t parent)
{
    nbr_t pn[MAXD];
    memcpy(pn, &t->nbrs[(size_t)map_find(t, parent) * MAXD], sizeof(pn));
    for (int d = 0; d < t->ndir; ++d)
    {
        const int od = d ^ 1;
        int       bc[MAXK];
        boundary_children(t, d, bc);
        if (pn[d].rel == ORC_SAME)
        {
            nbr_t nn;
            memset(&nn, 0, sizeof(nn));
            nn.rel = ORC_FINER;
            for (int i = 0; i < t->kf; ++i) nn.id[i] = child_of(t, parent, bc[i]);
            t->nbrs[(size_t)map_find(t, pn[d].id[0]) * MAXD + od] = nn;
        }
        else if (pn[d].rel == ORC_FINER)
        {
            for (int i = 0; i < t->kf; ++i)
            {
                nbr_t nn;
                memset(&nn, 0, sizeof(nn));
                nn.rel   = ORC_SAME;
                nn.id[0] = child_of(t, parent, bc[i]);
                t->nbrs[(size_t)map_find(t, pn[d].id[i]) * MAXD + od] = nn;
            }
        }
    }
}

/* ndtree.hpp:727-778 fragment(node) */
static int fragment_one(orc_tree* t, uint64_t node)
{
    const size_t from     = (size_t)map_find(t, node);
    const size_t start_to = t->n;
    for (int i = 0; i < t->fan; ++i)
    {
        nbr_t cn[MAXD];
        child_neighbors(t, node, &t->nbrs[from * MAXD], i, cn);
        if (append(t, child_of(t, node, i), cn)) return -1;
    }
    symmetry_after_refinement(t, node);
    transfer(t, from, start_to, 0);
    map_erase(t, node);
    return 0;
}

/* ndtree.hpp:1039-1125 */
static void symmetry_after_recombining(orc_tree* t, uint64_t parent, const nbr_t* pn)
{
    for (int d = 0; d < t->ndir; ++d)
    {
        const int od = d ^ 1;
        if (pn[d].rel == ORC_SAME)
        {
            const int64_t li = map_find(t, pn[d].id[0]);
            if (li >= 0)
            {
                nbr_t nn;
                memset(&nn, 0, sizeof(nn));
                nn.rel                             = ORC_SAME;
                nn.id[0]                           = parent;
                t->nbrs[(size_t)li * MAXD + od] = nn;
            }
        }
        else if (pn[d].rel == ORC_FINER)
        {
            for (int i = 0; i < t->kf; ++i)
            {
                nbr_t nn;
                memset(&nn, 0, sizeof(nn));
                nn.rel   = ORC_COARSER;
                nn.id[0] = parent;
                contact_quadrant(t, i, od, nn.quad);
                t->nbrs[(size_t)map_find(t, pn[d].id[i]) * MAXD + od] = nn;
            }
        }
    }
}

/* ndtree.hpp:780-834 recombine(parent) */
static int recombine_one(orc_tree* t, uint64_t parent)
{
    const size_t start = (size_t)map_find(t, child_of(t, parent, 0));
    nbr_t        cn[8 * MAXD];
    for (int i = 0; i < t->fan; ++i)
    {
        const uint64_t c = child_of(t, parent, i);
        memcpy(&cn[i * MAXD], &t->nbrs[(size_t)map_find(t, c) * MAXD], sizeof(nbr_t) * MAXD);
        map_erase(t, c);
    }
    nbr_t pn[MAXD];
    parent_neighbors(t, cn, pn);
    const size_t to = t->n;
    if (append(t, parent, pn)) return -1;
    transfer(t, to, start, 1);
    symmetry_after_recombining(t, parent, pn);
    return 0;
}

static int cmp_u64(const void* a, const void* b)
{
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y);
}

/* ndtree.hpp:904-940 apply_refine_coarsen (+ :1888-1917 eligibility) */
static void apply_refine_coarsen(orc_tree* t)
{
    t->n_refine = t->n_coarsen = 0;
    size_t    np               = 0;
    uint64_t* parents          = (uint64_t*)malloc((t->n + 1) * sizeof(uint64_t));
    for (size_t i = 0; i < t->n; ++i)
        if (t->ids[i] != 0) parents[np++] = parent_of(t, t->ids[i]);
    qsort(parents, np, sizeof(uint64_t), cmp_u64);
    size_t nu = 0;
    for (size_t i = 0; i < np; ++i)
        if (i == 0 || parents[i] != parents[i - 1]) parents[nu++] = parents[i];
    for (size_t i = 0; i < t->n; ++i)
        if (t->status[i] == ORC_REFINE && id_level(t->ids[i]) < t->depth)
            t->to_refine[t->n_refine++] = t->ids[i];
    for (size_t i = 0; i < nu; ++i)
    {
        int ok = 1;
        for (int c = 0; c < t->fan && ok; ++c)
        {
            const int64_t li = map_find(t, child_of(t, parents[i], c));
            ok               = li >= 0 && t->status[li] == ORC_COARSEN;
        }
        if (ok) t->to_coarsen[t->n_coarsen++] = parents[i];
    }
    free(parents);
}

static int in_refine(const orc_tree* t, uint64_t id)
{
    for (size_t i = 0; i < t->n_refine; ++i)
        if (t->to_refine[i] == id) return 1;
    return 0;
}

/* ndtree.hpp:1127-1240 balancing */
static int balancing(orc_tree* t)
{
    for (size_t i = 0; i < t->n_refine; ++i)
    {
        const nbr_t* nb = &t->nbrs[(size_t)map_find(t, t->to_refine[i]) * MAXD];
        for (int d = 0; d < t->ndir; ++d)
            if (nb[d].rel == ORC_COARSER && !in_refine(t, nb[d].id[0]))
            {
                if (t->n_refine >= t->cap) return -1;
                t->to_refine[t->n_refine++] = nb[d].id[0];
            }
    }
    /* :1167-1171 sort by level descending.  std::ranges::sort is not stable; the order inside a
     * level does not change the resulting leaf set (A1), so an insertion sort is used here. */
    for (size_t i = 1; i < t->n_refine; ++i)
    {
        const uint64_t v = t->to_refine[i];
        size_t         j = i;
        while (j > 0 && id_level(t->to_refine[j - 1]) < id_level(v))
        {
            t->to_refine[j] = t->to_refine[j - 1];
            --j;
        }
        t->to_refine[j] = v;
    }
    /* :1173-1239 coarsening veto */
    size_t keep = 0;
    for (size_t i = 0; i < t->n_coarsen; ++i)
    {
        const uint64_t parent = t->to_coarsen[i];
        int            veto   = 0;
        for (int d = 0; d < t->ndir && !veto; ++d)
        {
            int bc[MAXK];
            boundary_children(t, d, bc);
            for (int j = 0; j < t->kf && !veto; ++j)
            {
                const nbr_t* nb =
                    &t->nbrs[(size_t)map_find(t, child_of(t, parent, bc[j])) * MAXD + d];
                if (nb->rel == ORC_FINER) veto = 1;
                if (nb->rel == ORC_SAME && in_refine(t, nb->id[0])) veto = 1;
            }
        }
        if (!veto) t->to_coarsen[keep++] = parent;
    }
    t->n_coarsen = keep;
    return 0;
}

typedef struct
{
    uint64_t id;
    size_t   src;
} perm_t;
static int cmp_perm(const void* a, const void* b)
{
    const uint64_t x = ((const perm_t*)a)->id, y = ((const perm_t*)b)->id;
    return x < y ? -1 : (x > y);
}

/* ndtree.hpp:1919-1948 compact + :1335-1414 sort_buffers, expressed as one gather:
 * live slots in ascending id order; only CURRENT buffers are moved (next is overwritten
 * by the following step). */
static void compact_and_sort(orc_tree* t)
{
    perm_t* perm = (perm_t*)malloc(t->n * sizeof(perm_t));
    size_t  m    = 0;
    for (size_t i = 0; i < t->n; ++i)
        if (map_find(t, t->ids[i]) == (int64_t)i)
        {
            perm[m].id  = t->ids[i];
            perm[m].src = i;
            ++m;
        }
    qsort(perm, m, sizeof(perm_t), cmp_perm);
    uint64_t* ids    = (uint64_t*)malloc(m * sizeof(uint64_t));
    int8_t*   status = (int8_t*)malloc(m);
    nbr_t*    nbrs   = (nbr_t*)malloc(m * MAXD * sizeof(nbr_t));
    for (size_t i = 0; i < m; ++i)
    {
        ids[i]    = perm[i].id;
        status[i] = t->status[perm[i].src];
        memcpy(&nbrs[i * MAXD], &t->nbrs[perm[i].src * MAXD], sizeof(nbr_t) * MAXD);
    }
    memcpy(t->ids, ids, m * sizeof(uint64_t));
    memcpy(t->status, status, m);
    memcpy(t->nbrs, nbrs, m * MAXD * sizeof(nbr_t));
    double* tmp = (double*)malloc(m * t->flat * sizeof(double));
    for (int f = 0; f < t->nvar; ++f)
    {
        for (size_t i = 0; i < m; ++i)
            memcpy(tmp + i * t->flat, t->cur[f] + perm[i].src * t->flat, t->flat * sizeof(double));
        memcpy(t->cur[f], tmp, m * t->flat * sizeof(double));
    }
    free(tmp);
    free(ids);
    free(status);
    free(nbrs);
    free(perm);
    t->n = m;
    map_clear(t);
    for (size_t i = 0; i < m; ++i) map_set(t, t->ids[i], (int64_t)i);
}

/* ndtree.hpp:1249-1271 reconstruct_tree */
int orc_tree_reconstruct(orc_tree* t, const int8_t* flags)
{
    memcpy(t->status, flags, t->n);
    apply_refine_coarsen(t);
    if (t->n_refine == 0 && t->n_coarsen == 0) return 0;
    if (balancing(t)) return -1;
    if (t->n_refine == 0 && t->n_coarsen == 0) return 0;
    for (size_t i = t->n_refine; i > 0; --i) /* :836-859 back to front = coarsest first */
        if (fragment_one(t, t->to_refine[i - 1])) return -1;
    for (size_t i = 0; i < t->n_coarsen; ++i) /* :861-884 */
        if (recombine_one(t, t->to_coarsen[i])) return -1;
    compact_and_sort(t);
    return 1;
}

/* ndtree.hpp:691-725 neighbor_linear_index + :1606-1660 rebuild_halo_exchange_metadata */
void orc_tree_tables(const orc_tree* t, int8_t* rel, int32_t* nbr, int8_t* quad)
{
    for (size_t p = 0; p < t->n; ++p)
        for (int d = 0; d < t->ndir; ++d)
        {
            const nbr_t* nb = &t->nbrs[p * MAXD + d];
            const size_t o  = p * (size_t)t->ndir + (size_t)d;
            rel[o]          = nb->rel;
            for (int k = 0; k < t->kf; ++k) nbr[o * (size_t)t->kf + (size_t)k] = -1;
            for (int k = 0; k < t->rank; ++k) quad[o * (size_t)t->rank + (size_t)k] = 0;
            if (nb->rel == ORC_SAME || nb->rel == ORC_COARSER)
                nbr[o * (size_t)t->kf] = (int32_t)map_find(t, nb->id[0]);
            if (nb->rel == ORC_FINER)
                for (int k = 0; k < t->kf; ++k)
                    nbr[o * (size_t)t->kf + (size_t)k] = (int32_t)map_find(t, nb->id[k]);
            if (nb->rel == ORC_COARSER)
                for (int k = 0; k < t->rank; ++k)
                    quad[o * (size_t)t->rank + (size_t)k] = nb->quad[k];
        }
}

/* ------------------------------------------------------------------ halo exchange */
/* patch_utils.hpp:63-251 driver; operators :303-441; slab = patch_layout.hpp:72-110 */
static void halo_patch(orc_tree* t, size_t p)
{
    const int h = t->halo, R = t->rank;
    for (int d = 0; d < t->ndir; ++d)
    {
        const int    dim = d / 2, pos = d & 1;
        const nbr_t* nb  = &t->nbrs[p * MAXD + d];
        if (nb->rel == ORC_NONE) continue; /* boundary_t: no-op (:303-313) */
        size_t src[MAXK];
        src[0] = 0;
        if (nb->rel == ORC_FINER)
            for (int k = 0; k < t->kf; ++k) src[k] = (size_t)map_find(t, nb->id[k]);
        else
            src[0] = (size_t)map_find(t, nb->id[0]);
        int lo[MAXR], hi[MAXR], idx[MAXR];
        for (int k = 0; k < R; ++k)
        {
            lo[k] = (k == dim) ? (pos ? t->psize[k] - h : 0) : h;
            hi[k] = (k == dim) ? (pos ? t->psize[k] : h) : t->psize[k] - h;
            idx[k] = lo[k];
        }
        for (;;)
        {
            const size_t to = lin(t, idx);
            int          from[MAXR];
            memcpy(from, idx, sizeof(from));
            from[dim] += pos ? -t->size[dim] : t->size[dim];
            if (nb->rel == ORC_SAME)
            {
                /* same_t :315-332 */
                const size_t fi = lin(t, from);
                for (int f = 0; f < t->nvar; ++f)
                    t->cur[f][p * t->flat + to] = t->cur[f][src[0] * t->flat + fi];
            }
            else if (nb->rel == ORC_FINER)
            {
                /* finer_t :334-386 (note `% s_sizes[dim]`, :358) */
                int fine_patch = 0, stride = 1, base[MAXR];
                for (int k = 0; k < R; ++k)
                {
                    base[k] = ((from[k] - h) * 2) % t->size[dim] + h;
                    if (k != dim)
                    {
                        fine_patch += ((idx[k] - h) / (t->size[k] / 2)) * stride;
                        stride *= 2;
                    }
                }
                size_t offs[8];
                hypercube(t, lin(t, base), offs);
                for (int f = 0; f < t->nvar; ++f)
                {
                    const double* fn  = t->cur[f] + src[fine_patch] * t->flat;
                    double        sum = 0.0;
                    for (int n = 0; n < t->fan; ++n) sum += fn[offs[n]];
                    t->cur[f][p * t->flat + to] = sum / (double)t->fan;
                }
            }
            else
            {
                /* coarser_t :388-441 -> linear_interpolator::interpolation = injection */
                for (int k = 0; k < R; ++k)
                    from[k] = h + nb->quad[k] * (t->size[k] / 2) + (from[k] - h) / 2;
                const size_t fi = lin(t, from);
                for (int f = 0; f < t->nvar; ++f)
                    t->cur[f][p * t->flat + to] = t->cur[f][src[0] * t->flat + fi];
            }
            int k = R - 1;
            for (; k >= 0; --k)
            {
                if (++idx[k] < hi[k]) break;
                idx[k] = lo[k];
            }
            if (k < 0) break;
        }
    }
}

/* ndtree.hpp:1581-1596 */
void orc_halo_exchange(orc_tree* t)
{
#pragma omp parallel for schedule(static)
    for (long p = 0; p < (long)t->n; ++p) halo_patch(t, (size_t)p);
}

/* ------------------------------------------------------------------ solver */
/* solver/physics_system.hpp:58-85 cell_sizes: physical dim i <-> layout dim R-1-i */
static void cell_sizes(const orc_tree* t, uint64_t id, const double* L, double* dx)
{
    const int      lvl = id_level(id);
    const uint32_t pm  = 1u << (t->depth - lvl);
    for (int i = 0; i < t->rank; ++i)
    {
        const double patch = L[i] * (double)pm / (double)(1u << t->depth);
        dx[i]              = patch / (double)t->size[t->rank - 1 - i];
    }
}

/* solver/AdvectionPhysics.hpp:45-66 */
static void flux_advection(const double* UL, const double* UR, double* fl, int dir)
{
    static const double vel[3] = { 1.0, 0.5, 0.0 };
    const double        fL = UL[0] * vel[dir], fR = UR[0] * vel[dir];
    const double        smax = fabs(vel[dir]);
    fl[0]                    = 0.5 * (fL + fR) - 0.5 * smax * (UR[0] - UL[0]);
}

/* solver/EulerPhysics.hpp:74-129 */
static void flux_euler(int DIM, const double* UL, const double* UR, double* fl, int dir, double g)
{
    const double irL = 1.0 / UL[0];
    double       KL  = 0.0;
    for (int d = 0; d < DIM; ++d) KL += UL[1 + d] * UL[1 + d];
    KL *= 0.5 * irL;
    const double pL = (g - 1.0) * (UL[DIM + 1] - KL);
    const double aL = sqrt(g * pL * irL);
    const double uL = UL[1 + dir] * irL;
    const double irR = 1.0 / UR[0];
    double       KR  = 0.0;
    for (int d = 0; d < DIM; ++d) KR += UR[1 + d] * UR[1 + d];
    KR *= 0.5 * irR;
    const double pR   = (g - 1.0) * (UR[DIM + 1] - KR);
    const double aR   = sqrt(g * pR * irR);
    const double uR   = UR[1 + dir] * irR;
    const double smax = fmax(fabs(uL) + aL, fabs(uR) + aR);
    fl[0]             = 0.5 * (UL[1 + dir] + UR[1 + dir] - smax * (UR[0] - UL[0]));
    for (int d = 0; d < DIM; ++d)
    {
        double fLm = UL[1 + d] * uL, fRm = UR[1 + d] * uR;
        if (d == dir)
        {
            fLm += pL;
            fRm += pR;
        }
        fl[1 + d] = 0.5 * (fLm + fRm - smax * (UR[1 + d] - UL[1 + d]));
    }
    const double fLE = uL * (UL[DIM + 1] + pL), fRE = uR * (UR[DIM + 1] + pR);
    fl[DIM + 1]      = 0.5 * (fLE + fRE - smax * (UR[DIM + 1] - UL[DIM + 1]));
}

/* AdvectionPhysics.hpp:74-85, EulerPhysics.hpp:137-161 */
static double max_speed(const orc_tree* t, int eq, size_t cell, int dir, double g)
{
    if (eq == ORC_EQ_ADVECTION)
    {
        static const double vel[3] = { 1.0, 0.5, 0.0 };
        return fabs(vel[dir]);
    }
    const int    DIM = t->rank;
    const double ir  = 1.0 / t->cur[0][cell];
    double       K   = 0.0;
    for (int d = 0; d < DIM; ++d) K += t->cur[1 + d][cell] * t->cur[1 + d][cell];
    K *= 0.5 * ir;
    const double p = (g - 1.0) * (t->cur[DIM + 1][cell] - K);
    const double a = sqrt(g * p * ir);
    return fabs(t->cur[1 + dir][cell] * ir) + a;
}

/* amr_solver.hpp:355-413 */
double orc_compute_dt(const orc_tree* t, int eq, double gamma, double cfl, const double* L)
{
    double    dt = DBL_MAX;
    const int R  = t->rank;
#pragma omp parallel for schedule(static) reduction(min : dt)
    for (long p = 0; p < (long)t->n; ++p)
    {
        double dx[MAXR];
        cell_sizes(t, t->ids[p], L, dx);
        int idx[MAXR];
        for (int k = 0; k < R; ++k) idx[k] = t->halo;
        for (;;)
        {
            const size_t cell = (size_t)p * t->flat + lin(t, idx);
            for (int d = 0; d < R; ++d)
            {
                const double s = max_speed(t, eq, cell, d, gamma);
                if (s > 1e-12 && dx[d] / s < dt) dt = dx[d] / s;
            }
            int k = R - 1;
            for (; k >= 0; --k)
            {
                if (++idx[k] < t->halo + t->size[k]) break;
                idx[k] = t->halo;
            }
            if (k < 0) break;
        }
    }
    return cfl * dt;
}

/* amr_solver.hpp:265-353 */
void orc_time_step(orc_tree* t, int eq, double gamma, double dt, const double* L)
{
    const int R = t->rank, NV = t->nvar;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < (long)t->n; ++p)
    {
        double dx[MAXR];
        cell_sizes(t, t->ids[p], L, dx);
        int idx[MAXR];
        for (int k = 0; k < R; ++k) idx[k] = t->halo;
        for (;;)
        {
            const size_t cell = (size_t)p * t->flat + lin(t, idx);
            double       Uc[5], UL[5], UR[5], fL[5], fR[5], upd[5] = { 0, 0, 0, 0, 0 };
            for (int k = 0; k < NV; ++k) Uc[k] = t->cur[k][cell];
            for (int d = 0; d < R; ++d)
            {
                /* N1: solver direction d <-> layout dim R-1-d (:267-271, :317-319) */
                const size_t stride = (size_t)t->pstride[R - 1 - d];
                for (int k = 0; k < NV; ++k)
                {
                    UL[k] = t->cur[k][cell - stride];
                    UR[k] = t->cur[k][cell + stride];
                }
                if (eq == ORC_EQ_ADVECTION)
                {
                    flux_advection(UL, Uc, fL, d);
                    flux_advection(Uc, UR, fR, d);
                }
                else
                {
                    flux_euler(R, UL, Uc, fL, d, gamma);
                    flux_euler(R, Uc, UR, fR, d, gamma);
                }
                const double dt_over_dx = dt / dx[d];
                for (int k = 0; k < NV; ++k) upd[k] -= dt_over_dx * (fR[k] - fL[k]);
            }
            for (int k = 0; k < NV; ++k) t->nxt[k][cell] = t->cur[k][cell] + upd[k];
            int k = R - 1;
            for (; k >= 0; --k)
            {
                if (++idx[k] < t->halo + t->size[k]) break;
                idx[k] = t->halo;
            }
            if (k < 0) break;
        }
    }
    /* ndtree.hpp:1558-1579 swap_buffers, then :1862-1877 halo_exchange_update */
    double** tmp = t->cur;
    t->cur       = t->nxt;
    t->nxt       = tmp;
    orc_halo_exchange(t);
}

/* amr_solver.hpp:221-241 */
double orc_advance_batch(orc_tree* t, int eq, double gamma, double cfl, const double* L,
                         size_t steps, double remaining, size_t* executed, double* dts)
{
    double acc = 0.0;
    size_t cnt = 0;
    for (size_t s = 0; s < steps; ++s)
    {
        if (remaining <= 0.0) break;
        const double dt      = orc_compute_dt(t, eq, gamma, cfl, L);
        const double step_dt = dt < remaining ? dt : remaining;
        if (step_dt <= 0.0) break;
        orc_time_step(t, eq, gamma, step_dt, L);
        if (dts) dts[cnt] = step_dt;
        acc += step_dt;
        remaining -= step_dt;
        ++cnt;
    }
    if (executed) *executed = cnt;
    return acc;
}
