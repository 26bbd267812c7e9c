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
// forward through them.  "__GP<n>" is the reference's graph-partitioning order, which without METIS
// is the identity (matrix-market-reorder.cpp:172-180): accepted, nothing permuted.
#include "mm_host.hpp"

#include <algorithm>
#include <new>
#include <numeric>
#include <stdexcept>
#include <string>

// No exception crosses the C ABI: host-side allocations (std::vector, std::string, new) may throw.
#define SPMV_ABI_CATCH                                                                                  \
    catch (const std::bad_alloc &) { return ::spmvb200::fail(SPMVB200_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception & e) { return ::spmvb200::fail(SPMVB200_ERR_INVALID, e.what()); }

#include "../../include/spmv_b200.h"

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
    (void)nparts;
    if (!mm || !new_order) return fail(SPMVB200_ERR_INVALID, "null argument");
    for (int32_t v = 0; v < mm->rows; v++) new_order[v] = v;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_mm_permute(spmvb200_mm_t mm, const int32_t * new_order)
try {
    if (!mm || !new_order) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_permute(mm, new_order);
}
SPMV_ABI_CATCH

}  // extern "C"
