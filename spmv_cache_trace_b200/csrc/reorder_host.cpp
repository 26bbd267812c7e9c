// reorder_host.cpp -- the reference's matrix reorderings, on the host.
//
// Follows find_new_order_RCM (reference matrix/matrix-market-reorder.cpp:60-170) decision for decision so
// that "<path>__RCM" loads the same permuted matrix as the reference does (matrix-market.cpp:786-802,
// Matrix::permute :309-333): directed adjacency in file order with the diagonal left out and
// duplicates kept, component start = the untaken node of smallest degree (lowest index among equals),
// breadth-first levels with every batch of newly seen neighbours ordered by std::sort on degree alone
// (the same library call as the reference's, so equal degrees fall the same way), the whole order
// reversed at the end.  What differs is cost: the reference rescans all nodes for every component
// (quadratic on graphs with many isolated vertices); here the nodes are ranked once and a cursor moves
// forward through them.
//
// "__GP<n>" is the reference's graph-partitioning order (find_new_order_GP, :183-278): a K-way partition of the matrix
// graph by METIS_PartGraphKway, then the rows grouped by part.  METIS is third-party code that neither the reference's
// default build nor this image has; without it the reference returns the identity (:172-180), and so does
// spmvb200_mm_order_gp by default.  The grouping step (:246-266) is restated exactly (order_from_parts); for the
// partition itself this file has its own K-way partitioner (mm_partition_kway: level-structure cut + boundary
// refinement, unit vertex weights, the reference's 1.05 imbalance bound), switched in for "__GP<n>" by the global
// option "mm.gp_partitioner" -- a different partition than METIS would produce, so orders differ from a METIS build's.
#include "mm_host.hpp"

#include <algorithm>
#include <atomic>
#include <new>
#include <numeric>
#include <stdexcept>
#include <string>

// No exception crosses the C ABI: host-side allocations (std::vector, std::string, new) may throw.
#define SPMV_ABI_CATCH                                                                                  \
    catch (const std::bad_alloc &) { return ::spmvb200::fail(SPMVB200_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception & e) { return ::spmvb200::fail(SPMVB200_ERR_INVALID, e.what()); }

#include "../../include/spmv_b200.h"

#define SPMV_TRY_HOST(call)            \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

namespace spmvb200 {
int fail(int code, const std::string & msg);

int mm_order_rcm(const spmvb200_mm_s * m, int32_t * new_order)
{
    if (m->format != 0) return fail(SPMVB200_ERR_INVALID, "Expected matrix in coordinate format");
    if (m->rows != m->columns) return fail(SPMVB200_ERR_INVALID, "Expected a square matrix");
    if (m->field != 0) return fail(SPMVB200_ERR_INVALID, "Expected matrix with real values");
    const int32_t n = m->rows;
    const int64_t ne = m->num_entries;
    std::vector<int32_t> degrees((size_t)n, 0);
    for (int64_t e = 0; e < ne; e++)
        if (m->i[e] != m->j[e]) degrees[(size_t)m->i[e] - 1]++;
    std::vector<int64_t> first((size_t)n + 1, 0);
    for (int32_t v = 0; v < n; v++) first[(size_t)v + 1] = first[v] + degrees[v];
    std::vector<int32_t> adjacency((size_t)first[n]);
    {
        std::vector<int64_t> fill(first.begin(), first.end() - 1);
        for (int64_t e = 0; e < ne; e++)
            if (m->i[e] != m->j[e]) adjacency[(size_t)fill[(size_t)m->i[e] - 1]++] = m->j[e] - 1;
    }
    // nodes by (degree, index): the reference's scan `degrees[i] < min_degree` keeps the first minimum
    std::vector<int32_t> ranked((size_t)n);
    std::iota(ranked.begin(), ranked.end(), 0);
    std::stable_sort(ranked.begin(), ranked.end(), [&](int32_t a, int32_t b) { return degrees[a] < degrees[b]; });
    size_t cursor = 0;

    std::vector<int32_t> R;
    R.reserve((size_t)n);
    std::vector<char> not_taken((size_t)n, 1), not_visited((size_t)n, 1);
    std::vector<int32_t> queue, batch;
    queue.reserve((size_t)n);
    auto take = [&](int32_t v) {
        R.push_back(v);
        not_taken[v] = 0;
        not_visited[v] = 0;
        batch.clear();
        for (int64_t p = first[v]; p < first[(size_t)v + 1]; p++) {
            const int32_t k = adjacency[(size_t)p];
            if (not_visited[k]) {
                batch.push_back(k);
                not_visited[k] = 0;
            }
        }
        if (batch.size() > 1)
            std::sort(batch.begin(), batch.end(), [&degrees](int i1, int i2) { return degrees[i1] < degrees[i2]; });
        queue.insert(queue.end(), batch.begin(), batch.end());
    };
    while ((int32_t)R.size() < n) {
        while (cursor < (size_t)n && !not_taken[ranked[cursor]]) cursor++;
        const int32_t start = ranked[cursor];
        if (degrees[start] >= n)  // the reference's scan starts from min_degree = n and would find no node
            return fail(SPMVB200_ERR_INVALID, "RCM: a row holds more off-diagonal entries than the matrix has rows");
        queue.clear();
        take(start);
        for (size_t q = 0; q < queue.size(); q++)
            if (not_taken[queue[q]]) take(queue[q]);
    }
    std::reverse(R.begin(), R.end());
    for (int32_t k = 0; k < n; k++) new_order[R[(size_t)k]] = k;
    return 0;
}


// ---- graph partitioning ----------------------------------------------------------------------------------------------

static std::atomic<int64_t> g_gp_partitioner{0};
void set_gp_partitioner(int64_t v) { g_gp_partitioner.store(v); }
int64_t gp_partitioner() { return g_gp_partitioner.load(); }

// The second half of find_new_order_GP (matrix-market-reorder.cpp:246-266): count the rows of every part, turn the
// counts into offsets, give the rows of a part consecutive new indices in ascending old index; new_order[old] = new.
int order_from_parts(int32_t n, int32_t nparts, const int32_t * part, int32_t * new_order)
{
    if (n < 0 || nparts < 1) return fail(SPMVB200_ERR_INVALID, "order_from_parts: bad size");
    std::vector<int32_t> offset((size_t)nparts + 1, 0);
    for (int32_t v = 0; v < n; v++) {
        if (part[v] < 0 || part[v] >= nparts) return fail(SPMVB200_ERR_INVALID, "order_from_parts: part number outside [0, nparts)");
        offset[(size_t)part[v] + 1]++;
    }
    for (int32_t p = 1; p <= nparts; p++) offset[(size_t)p] += offset[(size_t)p - 1];
    for (int32_t v = 0; v < n; v++) new_order[v] = offset[(size_t)part[v]]++;
    return 0;
}

namespace {

// The matrix graph as METIS wants it: undirected, simple (every off-diagonal entry is an edge in both directions,
// duplicates merged).  The reference hands METIS the directed pattern as stored (:203-227), which is the same graph for
// the structurally symmetric matrices it is meant for.
struct Graph {
    int32_t n = 0;
    std::vector<int64_t> first;
    std::vector<int32_t> adj;
};

Graph build_graph(const spmvb200_mm_s * m)
{
    Graph g;
    g.n = m->rows;
    const int64_t ne = m->num_entries;
    std::vector<int64_t> deg((size_t)g.n + 1, 0);
    for (int64_t e = 0; e < ne; e++)
        if (m->i[(size_t)e] != m->j[(size_t)e]) { deg[(size_t)m->i[(size_t)e]]++; deg[(size_t)m->j[(size_t)e]]++; }
    for (int32_t v = 0; v < g.n; v++) deg[(size_t)v + 1] += deg[(size_t)v];
    std::vector<int32_t> raw((size_t)deg[(size_t)g.n]);
    {
        std::vector<int64_t> fill(deg.begin(), deg.end() - 1);
        for (int64_t e = 0; e < ne; e++) {
            const int32_t a = m->i[(size_t)e] - 1, b = m->j[(size_t)e] - 1;
            if (a == b) continue;
            raw[(size_t)fill[(size_t)a]++] = b;
            raw[(size_t)fill[(size_t)b]++] = a;
        }
    }
    g.first.assign((size_t)g.n + 1, 0);
    g.adj.reserve(raw.size() / 2 + 16);
    for (int32_t v = 0; v < g.n; v++) {
        auto lo = raw.begin() + deg[(size_t)v], hi = raw.begin() + deg[(size_t)v + 1];
        std::sort(lo, hi);
        hi = std::unique(lo, hi);
        g.adj.insert(g.adj.end(), lo, hi);
        g.first[(size_t)v + 1] = (int64_t)g.adj.size();
    }
    return g;
}

// Breadth-first order of the component of `start` inside region `rid` (neighbours in ascending index); returns the
// last vertex reached.
int32_t bfs(const Graph & g, const std::vector<int32_t> & region, int32_t rid, int32_t start, std::vector<int32_t> & mark,
            int32_t stamp, std::vector<int32_t> & out)
{
    const size_t head0 = out.size();
    out.push_back(start);
    mark[(size_t)start] = stamp;
    for (size_t h = head0; h < out.size(); h++) {
        const int32_t v = out[h];
        for (int64_t p = g.first[(size_t)v]; p < g.first[(size_t)v + 1]; p++) {
            const int32_t w = g.adj[(size_t)p];
            if (region[(size_t)w] == rid && mark[(size_t)w] != stamp) { mark[(size_t)w] = stamp; out.push_back(w); }
        }
    }
    return out.back();
}

struct Bisector {
    const Graph & g;
    int32_t * part;
    std::vector<int32_t> region, mark, scratch;
    int32_t next_region = 0, stamp = 0;

    // `verts` (ascending) get parts [part0, part0 + k): level structure of the induced subgraph from a pseudo-peripheral
    // vertex of every component, cut in proportion k/2 : k - k/2, both sides recursively.
    void run(std::vector<int32_t> & verts, int32_t k, int32_t part0)
    {
        if (k <= 1 || verts.size() <= 1) {
            for (int32_t v : verts) part[v] = part0;
            return;
        }
        const int32_t rid = ++next_region;
        for (int32_t v : verts) region[(size_t)v] = rid;
        std::vector<int32_t> order;
        order.reserve(verts.size());
        for (int32_t v0 : verts) {
            if (region[(size_t)v0] != rid) continue;  // placed already (its region was flipped below)
            scratch.clear();
            const int32_t far1 = bfs(g, region, rid, v0, mark, ++stamp, scratch);
            scratch.clear();
            const int32_t far2 = bfs(g, region, rid, far1, mark, ++stamp, scratch);
            const size_t before = order.size();
            bfs(g, region, rid, far2, mark, ++stamp, order);
            for (size_t t = before; t < order.size(); t++) region[(size_t)order[t]] = -rid;  // placed
        }
        const int32_t k1 = k / 2;
        const size_t cut = (size_t)((int64_t)order.size() * k1 / k);
        std::vector<int32_t> left(order.begin(), order.begin() + (ptrdiff_t)cut), right(order.begin() + (ptrdiff_t)cut, order.end());
        std::vector<int32_t>().swap(order);
        std::vector<int32_t>().swap(verts);
        std::sort(left.begin(), left.end());
        std::sort(right.begin(), right.end());
        run(left, k1, part0);
        run(right, k - k1, part0 + k1);
    }
};

}  // namespace

// K-way partition with unit vertex weights; every part holds at most max(ceil(n/k), floor(ub*n/k)) vertices.
//   1. recursive bisection by level structures: the vertices of a region are ordered breadth-first from a
//      pseudo-peripheral vertex of every component (the far end of a sweep from the component's lowest vertex, then the
//      far end of a sweep from there), the order is cut in proportion k/2 : k - k/2 -- edges only join neighbouring
//      levels, so the cut is one level wide -- and both sides are bisected again until k parts exist;
//   2. boundary refinement: sweeps over the vertices; a vertex moves to the neighbouring part that holds most of its
//      neighbours when that lowers the cut (or keeps it and evens the sizes out) and the target has room; until a
//      sweep moves nothing (at most 16 sweeps).  Every move lowers (cut, imbalance) lexicographically: it terminates.
// Deterministic.  edgecut = undirected edges whose ends lie in different parts.
int mm_partition_kway(const spmvb200_mm_s * m, int32_t nparts, double ub, int32_t * part, int64_t * edgecut)
{
    if (m->format != 0) return fail(SPMVB200_ERR_INVALID, "Expected matrix in coordinate format");
    if (m->rows != m->columns) return fail(SPMVB200_ERR_INVALID, "Expected a square matrix");
    if (m->field != 0) return fail(SPMVB200_ERR_INVALID, "Expected matrix with real values");
    if (nparts < 1) return fail(SPMVB200_ERR_INVALID, "partition: nparts must be positive");
    if (!(ub >= 1.0)) return fail(SPMVB200_ERR_INVALID, "partition: the imbalance bound must be at least 1");
    const Graph g = build_graph(m);
    const int32_t n = g.n, k = nparts;
    if (edgecut) *edgecut = 0;
    if (n == 0) return 0;
    // 1. recursive bisection
    std::vector<int32_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    {
        Bisector b{g, part, std::vector<int32_t>((size_t)n, 0), std::vector<int32_t>((size_t)n, 0), {}, 0, 0};
        b.scratch.reserve((size_t)n);
        std::vector<int32_t> all(order);
        b.run(all, k, 0);
    }
    // the refinement sweeps visit the vertices part by part
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t c) { return part[a] < part[c]; });
    // 2. boundary refinement
    const int64_t cap = std::max<int64_t>(((int64_t)n + k - 1) / k, (int64_t)(ub * (double)n / (double)k));
    std::vector<int64_t> size((size_t)k, 0);
    for (int32_t v = 0; v < n; v++) size[(size_t)part[v]]++;
    std::vector<int32_t> cnt((size_t)k, 0), touched;
    for (int sweep = 0; sweep < 16; sweep++) {
        int64_t moves = 0;
        for (int32_t t = 0; t < n; t++) {
            const int32_t v = order[(size_t)t], p = part[v];
            touched.clear();
            for (int64_t q = g.first[(size_t)v]; q < g.first[(size_t)v + 1]; q++) {
                const int32_t pw = part[g.adj[(size_t)q]];
                if (cnt[(size_t)pw]++ == 0) touched.push_back(pw);
            }
            int32_t best = -1;
            for (int32_t q : touched)
                if (q != p && (best < 0 || cnt[(size_t)q] > cnt[(size_t)best] || (cnt[(size_t)q] == cnt[(size_t)best] && q < best))) best = q;
            if (best >= 0 && size[(size_t)p] > 1 && size[(size_t)best] < cap) {
                const int32_t gain = cnt[(size_t)best] - cnt[(size_t)p];
                if (gain > 0 || (gain == 0 && size[(size_t)best] + 1 < size[(size_t)p])) {
                    part[v] = best;
                    size[(size_t)p]--;
                    size[(size_t)best]++;
                    moves++;
                }
            }
            for (int32_t q : touched) cnt[(size_t)q] = 0;
        }
        if (!moves) break;
    }
    if (edgecut) {
        int64_t cut = 0;
        for (int32_t v = 0; v < n; v++)
            for (int64_t q = g.first[(size_t)v]; q < g.first[(size_t)v + 1]; q++)
                if (g.adj[(size_t)q] > v && part[g.adj[(size_t)q]] != part[v]) cut++;
        *edgecut = cut;
    }
    return 0;
}

// find_new_order_GP (matrix-market-reorder.cpp:183-278) with the partitioner above in METIS's place.
int mm_order_gp_kway(const spmvb200_mm_s * m, int32_t nparts, int32_t * new_order)
{
    if (nparts <= 1) nparts = 16;  // :232-233
    std::vector<int32_t> part((size_t)std::max(m->rows, 1));
    SPMV_TRY_HOST(mm_partition_kway(m, nparts, 1.05, part.data(), nullptr));  // ubvec = 1.05, :200
    return order_from_parts(m->rows, nparts, part.data(), new_order);
}

int mm_permute(spmvb200_mm_s * m, const int32_t * new_order)
{
    if (m->format != 0) return fail(SPMVB200_ERR_INVALID, "Expected matrix in coordinate format");
    if (m->field != 0) return fail(SPMVB200_ERR_INVALID, "Expected matrix with real values");
    if (m->rows != m->columns) return fail(SPMVB200_ERR_INVALID, "The dimension of the matrix doesn't match");
    for (int32_t v = 0; v < m->rows; v++)
        if (new_order[v] < 0 || new_order[v] >= m->rows) return fail(SPMVB200_ERR_INVALID, "permutation entry outside the matrix");
    for (int64_t e = 0; e < m->num_entries; e++) {  // matrix-market.cpp:329-332
        m->i[(size_t)e] = new_order[(size_t)m->i[(size_t)e] - 1] + 1;
        m->j[(size_t)e] = new_order[(size_t)m->j[(size_t)e] - 1] + 1;
    }
    return 0;
}

}  // namespace spmvb200

using namespace spmvb200;

extern "C" {

int spmvb200_mm_order_rcm(spmvb200_mm_t mm, int32_t * new_order)
try {
    if (!mm || !new_order) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_order_rcm(mm, new_order);
}
SPMV_ABI_CATCH

int spmvb200_mm_order_gp(spmvb200_mm_t mm, int32_t nparts, int32_t * new_order)
try {
    if (!mm || !new_order) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (gp_partitioner()) return mm_order_gp_kway(mm, nparts, new_order);
    for (int32_t v = 0; v < mm->rows; v++) new_order[v] = v;  // the reference without METIS
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_mm_order_gp_kway(spmvb200_mm_t mm, int32_t nparts, int32_t * new_order)
try {
    if (!mm || !new_order) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_order_gp_kway(mm, nparts, new_order);
}
SPMV_ABI_CATCH

int spmvb200_mm_partition_kway(spmvb200_mm_t mm, int32_t nparts, int32_t ub_permille, int32_t * part, int64_t * edgecut)
try {
    if (!mm || (!part && mm->rows > 0)) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_partition_kway(mm, nparts, (double)ub_permille / 1000.0, part, edgecut);
}
SPMV_ABI_CATCH

int spmvb200_order_from_parts(int32_t n, int32_t nparts, const int32_t * part, int32_t * new_order)
try {
    if (n > 0 && (!part || !new_order)) return fail(SPMVB200_ERR_INVALID, "null argument");
    return order_from_parts(n, nparts, part, new_order);
}
SPMV_ABI_CATCH

int spmvb200_mm_permute(spmvb200_mm_t mm, const int32_t * new_order)
try {
    if (!mm || !new_order) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_permute(mm, new_order);
}
SPMV_ABI_CATCH

}  // extern "C"
