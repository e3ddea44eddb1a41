/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement, in plain C, of the
 * reference's nearest-neighbour path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (avoid-mpc_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_knn.py checks every function here
 * against the reference's own KDTreeTwo<double>/nanoflann compiled from
 * /root/reference (oracle/_ref/libampc_ref_kdtree.so) and against the golden
 * vectors minted from it (tests/golden/knn_*.npz).
 *
 * Reference (paths under roswrapper/ros/src/avoid_mpc/):
 *   include/kd_tree_two.h:11-51     PointCloudTwo: float xyz in 16-byte records, read as double
 *   include/kd_tree_two.h:88-106    Initialize: drop points whose x is NaN, then buildIndex()
 *   include/kd_tree_two.h:108-133   SearchForNearest (+ "size == k -> 0 results" quirk)
 *   include/nanoflann_two.hpp:590-599   L2_Simple_Adaptor::evalMetric  (sum over x,y,z of diff*diff)
 *   include/nanoflann_two.hpp:219-246   KNNResultSet::addPoint         (insertion sort, strict >)
 *   include/nanoflann_two.hpp:1055-1106 divideTree, :1197-1294 middleSplit_/planeSplit
 *   include/nanoflann_two.hpp:1563-1586 findNeighbors, :1729-1793 searchLevel
 *
 * Compile with -ffp-contract=off: dist2 is ((dx*dx + dy*dy) + dz*dz) with each
 * product and sum rounded separately (what evalMetric does without FMA fusion).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LEAF_MAX 10 /* kd_tree_two.h:68 */

static inline const float *pt_at(const void *xyz, int64_t stride, int64_t i) {
    return (const float *)((const char *)xyz + i * stride);
}

/* evalMetric, nanoflann_two.hpp:590-599, with kdtree_get_pt (kd_tree_two.h:34-41). */
static inline double dist2_ref(const double q[3], const float *p) {
    double result = 0.0;
    const double d0 = q[0] - (double)p[0];
    result += d0 * d0;
    const double d1 = q[1] - (double)p[1];
    result += d1 * d1;
    const double d2 = q[2] - (double)p[2];
    result += d2 * d2;
    return result;
}

/* kd_tree_two.h:99-101: copy the records whose x is not NaN, preserving order.
 * out16 receives 16-byte records; returns the surviving count. */
int64_t knn_oracle_filter_nan(const void *xyz, int64_t n, int64_t stride, void *out16) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pt_at(xyz, stride, i);
        if (!(p[0] != p[0])) {
            float *o = (float *)((char *)out16 + m * 16);
            o[0] = p[0];
            o[1] = p[1];
            o[2] = p[2];
            o[3] = 1.0f;
            ++m;
        }
    }
    return m;
}

/* SearchForNearest's result count rule, kd_tree_two.h:117-124. */
static int num_results_rule(int64_t npts, int k) {
    if (npts < k)
        return (int)npts;
    if (npts > k)
        return k;
    return 0; /* npts == k: the reference returns nothing */
}

/* Exhaustive exact k-NN in the canonical order (dist2 ascending, then index
 * ascending).  Equal to the kd-tree result whenever the k+1 nearest have
 * distinct dist2 (nanoflann's order among exact ties is traversal dependent).
 * The cloud must already be NaN-filtered.  Returns the number of results. */
int knn_oracle_bruteforce(const void *xyz, int64_t n, int64_t stride, const double q[3], int k,
                          int32_t *idx_out, double *dist2_out) {
    int count = 0;
    if (n <= 0 || k <= 0)
        return 0;
    for (int64_t i = 0; i < n; ++i) {
        const double d = dist2_ref(q, pt_at(xyz, stride, i));
        /* insert (d, i) keeping (dist2, idx) lexicographic order; i ascends so a
         * tie always goes after the entries already present. */
        int j = count;
        if (count == k) {
            if (!(d < dist2_out[k - 1]))
                continue;
            j = k - 1;
        }
        while (j > 0 && dist2_out[j - 1] > d) {
            dist2_out[j] = dist2_out[j - 1];
            idx_out[j] = idx_out[j - 1];
            --j;
        }
        dist2_out[j] = d;
        idx_out[j] = (int32_t)i;
        if (count < k)
            ++count;
    }
    return num_results_rule(n, k) == 0 ? 0 : (count < k ? count : k);
}

/* 1 if the k+1 nearest of q have pairwise distinct dist2 (so canonical order ==
 * nanoflann order and index parity is well defined). */
int knn_oracle_is_tie_free(const void *xyz, int64_t n, int64_t stride, const double q[3], int k) {
    int kk = k + 1;
    if (n < kk)
        kk = (int)n;
    if (kk <= 1)
        return 1;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)kk);
    double *d2 = (double *)malloc(sizeof(double) * (size_t)kk);
    /* n > kk is not required here: use the raw insertion (no quirk) */
    int count = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double d = dist2_ref(q, pt_at(xyz, stride, i));
        int j = count;
        if (count == kk) {
            if (!(d < d2[kk - 1])) {
                if (d == d2[kk - 1]) { /* a tie at the boundary */
                    free(idx);
                    free(d2);
                    return 0;
                }
                continue;
            }
            j = kk - 1;
        }
        while (j > 0 && d2[j - 1] > d) {
            d2[j] = d2[j - 1];
            idx[j] = idx[j - 1];
            --j;
        }
        d2[j] = d;
        idx[j] = (int32_t)i;
        if (count < kk)
            ++count;
    }
    int ok = 1;
    for (int j = 1; j < count; ++j)
        if (d2[j] == d2[j - 1])
            ok = 0;
    free(idx);
    free(d2);
    return ok;
}

/* ------------------------------------------------------------------------ */
/* kd-tree restatement (nanoflann semantics, own data layout).               */

typedef struct {
    int32_t child1, child2; /* -1/-1 for a leaf */
    int32_t left, right;    /* leaf: [left,right) into vacc */
    int32_t divfeat;
    double divlow, divhigh;
} kd_node;

typedef struct knn_tree {
    const char *xyz; /* borrowed: 16-byte records, NaN-filtered */
    int64_t n;
    uint32_t *vacc;
    kd_node *nodes;
    int64_t n_nodes, cap_nodes;
    double root_lo[3], root_hi[3];
    int32_t root;
} knn_tree;

static inline double get_pt(const knn_tree *t, uint32_t idx, int dim) {
    return (double)((const float *)(t->xyz + (int64_t)idx * 16))[dim];
}

static int32_t new_node(knn_tree *t) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? 2 * t->cap_nodes : 1024;
        t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
    }
    return (int32_t)t->n_nodes++;
}

/* planeSplit, nanoflann_two.hpp:1249-1294 */
static void plane_split(knn_tree *t, int64_t ind, int64_t count, int cutfeat, double cutval,
                        int64_t *lim1, int64_t *lim2) {
    uint32_t *v = t->vacc + ind;
    int64_t left = 0, right = count - 1;
    for (;;) {
        while (left <= right && get_pt(t, v[left], cutfeat) < cutval)
            ++left;
        while (right && left <= right && get_pt(t, v[right], cutfeat) >= cutval)
            --right;
        if (left > right || !right)
            break;
        uint32_t tmp = v[left];
        v[left] = v[right];
        v[right] = tmp;
        ++left;
        --right;
    }
    *lim1 = left;
    right = count - 1;
    for (;;) {
        while (left <= right && get_pt(t, v[left], cutfeat) <= cutval)
            ++left;
        while (right && left <= right && get_pt(t, v[right], cutfeat) > cutval)
            --right;
        if (left > right || !right)
            break;
        uint32_t tmp = v[left];
        v[left] = v[right];
        v[right] = tmp;
        ++left;
        --right;
    }
    *lim2 = left;
}

/* divideTree (nanoflann_two.hpp:1055-1106) + middleSplit_ (:1197-1247) */
static int32_t divide_tree(knn_tree *t, int64_t left, int64_t right, double lo[3], double hi[3]) {
    const int32_t id = new_node(t);
    if (right - left <= LEAF_MAX) {
        t->nodes[id].child1 = t->nodes[id].child2 = -1;
        t->nodes[id].left = (int32_t)left;
        t->nodes[id].right = (int32_t)right;
        for (int i = 0; i < 3; ++i)
            lo[i] = hi[i] = get_pt(t, t->vacc[left], i);
        for (int64_t k = left + 1; k < right; ++k)
            for (int i = 0; i < 3; ++i) {
                const double val = get_pt(t, t->vacc[k], i);
                if (lo[i] > val)
                    lo[i] = val;
                if (hi[i] < val)
                    hi[i] = val;
            }
        return id;
    }
    const int64_t count = right - left;
    const double EPS = 0.00001;
    double max_span = hi[0] - lo[0];
    for (int i = 1; i < 3; ++i)
        if (hi[i] - lo[i] > max_span)
            max_span = hi[i] - lo[i];
    double max_spread = -1, min_elem = 0, max_elem = 0;
    int cutfeat = 0;
    for (int i = 0; i < 3; ++i) {
        if (hi[i] - lo[i] > (1 - EPS) * max_span) {
            double mn = get_pt(t, t->vacc[left], i), mx = mn;
            for (int64_t k = 1; k < count; ++k) {
                const double val = get_pt(t, t->vacc[left + k], i);
                if (val < mn)
                    mn = val;
                if (val > mx)
                    mx = val;
            }
            if (mx - mn > max_spread) {
                cutfeat = i;
                max_spread = mx - mn;
                min_elem = mn;
                max_elem = mx;
            }
        }
    }
    const double split_val = (lo[cutfeat] + hi[cutfeat]) / 2;
    double cutval = split_val;
    if (split_val < min_elem)
        cutval = min_elem;
    else if (split_val > max_elem)
        cutval = max_elem;
    int64_t lim1, lim2, index;
    plane_split(t, left, count, cutfeat, cutval, &lim1, &lim2);
    if (lim1 > count / 2)
        index = lim1;
    else if (lim2 < count / 2)
        index = lim2;
    else
        index = count / 2;

    double llo[3], lhi[3], rlo[3], rhi[3];
    memcpy(llo, lo, sizeof llo);
    memcpy(lhi, hi, sizeof lhi);
    memcpy(rlo, lo, sizeof rlo);
    memcpy(rhi, hi, sizeof rhi);
    lhi[cutfeat] = cutval;
    rlo[cutfeat] = cutval;
    const int32_t c1 = divide_tree(t, left, left + index, llo, lhi);
    const int32_t c2 = divide_tree(t, left + index, right, rlo, rhi);
    kd_node *nd = &t->nodes[id]; /* re-fetch: nodes may have been realloc'd */
    nd->child1 = c1;
    nd->child2 = c2;
    nd->divfeat = cutfeat;
    nd->divlow = lhi[cutfeat];
    nd->divhigh = rlo[cutfeat];
    for (int i = 0; i < 3; ++i) {
        lo[i] = llo[i] < rlo[i] ? llo[i] : rlo[i];
        hi[i] = lhi[i] > rhi[i] ? lhi[i] : rhi[i];
    }
    return id;
}

/* buildIndex (nanoflann_two.hpp:1518-1541).  xyz16 must stay alive and be
 * NaN-filtered (knn_oracle_filter_nan). */
knn_tree *knn_oracle_tree_build(const void *xyz16, int64_t n) {
    knn_tree *t = (knn_tree *)calloc(1, sizeof(knn_tree));
    t->xyz = (const char *)xyz16;
    t->n = n;
    t->root = -1;
    if (n == 0)
        return t;
    t->vacc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i)
        t->vacc[i] = (uint32_t)i;
    for (int i = 0; i < 3; ++i)
        t->root_lo[i] = t->root_hi[i] = get_pt(t, 0, i);
    for (int64_t k = 1; k < n; ++k)
        for (int i = 0; i < 3; ++i) {
            const double val = get_pt(t, (uint32_t)k, i);
            if (val < t->root_lo[i])
                t->root_lo[i] = val;
            if (val > t->root_hi[i])
                t->root_hi[i] = val;
        }
    t->root = divide_tree(t, 0, n, t->root_lo, t->root_hi);
    return t;
}

void knn_oracle_tree_free(knn_tree *t) {
    if (!t)
        return;
    free(t->vacc);
    free(t->nodes);
    free(t);
}

typedef struct {
    int32_t *idx;
    double *d;
    int cap, count;
} result_set;

/* KNNResultSet::addPoint, nanoflann_two.hpp:219-246 */
static void add_point(result_set *r, double dist, int32_t index) {
    int i;
    for (i = r->count; i > 0; --i) {
        if (r->d[i - 1] > dist) {
            if (i < r->cap) {
                r->d[i] = r->d[i - 1];
                r->idx[i] = r->idx[i - 1];
            }
        } else
            break;
    }
    if (i < r->cap) {
        r->d[i] = dist;
        r->idx[i] = index;
    }
    if (r->count < r->cap)
        r->count++;
}

/* searchLevel, nanoflann_two.hpp:1729-1793 (eps = 0) */
static void search_level(const knn_tree *t, result_set *r, const double q[3], int32_t node,
                         double mindist, double dists[3]) {
    const kd_node *nd = &t->nodes[node];
    if (nd->child1 < 0 && nd->child2 < 0) {
        const double worst = r->d[r->cap - 1];
        for (int32_t i = nd->left; i < nd->right; ++i) {
            const uint32_t acc = t->vacc[i];
            const double dist = dist2_ref(q, (const float *)(t->xyz + (int64_t)acc * 16));
            if (dist < worst)
                add_point(r, dist, (int32_t)acc);
        }
        return;
    }
    const int idx = nd->divfeat;
    const double val = q[idx];
    const double diff1 = val - nd->divlow;
    const double diff2 = val - nd->divhigh;
    int32_t best, other;
    double cut;
    if (diff1 + diff2 < 0) {
        best = nd->child1;
        other = nd->child2;
        cut = (val - nd->divhigh) * (val - nd->divhigh);
    } else {
        best = nd->child2;
        other = nd->child1;
        cut = (val - nd->divlow) * (val - nd->divlow);
    }
    search_level(t, r, q, best, mindist, dists);
    const double dst = dists[idx];
    mindist = mindist + cut - dst;
    dists[idx] = cut;
    if (mindist * 1.0f <= r->d[r->cap - 1])
        search_level(t, r, q, other, mindist, dists);
    dists[idx] = dst;
}

/* SearchForNearest (kd_tree_two.h:108-133) over the restated tree.
 * Returns the number of results (idx/dist2 ascending, nanoflann tie order). */
int knn_oracle_tree_search(const knn_tree *t, const double q[3], int k, int32_t *idx_out,
                           double *dist2_out) {
    if (t->n <= 0 || k <= 0)
        return 0;
    result_set r = {idx_out, dist2_out, k, 0};
    dist2_out[k - 1] = DBL_MAX;
    double dists[3] = {0, 0, 0}, dist = 0;
    for (int i = 0; i < 3; ++i) { /* computeInitialDistances, :1296-1314 */
        if (q[i] < t->root_lo[i]) {
            dists[i] = (q[i] - t->root_lo[i]) * (q[i] - t->root_lo[i]);
            dist += dists[i];
        }
        if (q[i] > t->root_hi[i]) {
            dists[i] = (q[i] - t->root_hi[i]) * (q[i] - t->root_hi[i]);
            dist += dists[i];
        }
    }
    search_level(t, &r, q, t->root, dist, dists);
    return num_results_rule(t->n, k);
}

void knn_oracle_tree_search_batch(const knn_tree *t, const double *q, int Q, int k,
                                  int32_t *idx_out, double *dist2_out, int32_t *count_out) {
    for (int i = 0; i < Q; ++i)
        count_out[i] = knn_oracle_tree_search(t, q + 3 * i, k, idx_out + (size_t)i * k,
                                              dist2_out + (size_t)i * k);
}
