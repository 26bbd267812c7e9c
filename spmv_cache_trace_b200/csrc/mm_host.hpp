// mm_host.hpp -- host-side Matrix Market matrix behind spmvb200_mm_t.
//
// Stands in for matrix_market::Matrix of the reference (matrix/matrix-market.hpp:78-136) as far as
// the SpMV path needs it: size, the header's field/symmetry/format, and the coordinate entries as
// three parallel arrays (1-based i, j; a = values_real() semantics).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

struct spmvb200_mm_s {
    int32_t rows = 0, columns = 0, num_entries = 0;
    int32_t format = 0;    // 0 coordinate, 1 array
    int32_t field = 0;     // 0 real, 1 complex, 2 integer, 3 pattern
    int32_t symmetry = 0;  // 0 general, 1 symmetric, 2 skew-symmetric, 3 hermitian
    std::vector<int32_t> i, j;
    std::vector<double> a;
};

namespace spmvb200 {

int mm_parse_text(const char * text, size_t len, spmvb200_mm_s ** out);
int mm_load_path(const char * path, spmvb200_mm_s ** out);
int mm_from_entries(int32_t rows, int32_t columns, int32_t n, const int32_t * i, const int32_t * j,
                    const double * a, spmvb200_mm_s ** out);
int mm_row_lengths(const spmvb200_mm_s * mm, int32_t * lengths);
int mm_sort(spmvb200_mm_s * mm, bool row_major);
// reorder_host.cpp
int mm_order_rcm(const spmvb200_mm_s * mm, int32_t * new_order);
int mm_permute(spmvb200_mm_s * mm, const int32_t * new_order);
int order_from_parts(int32_t n, int32_t nparts, const int32_t * part, int32_t * new_order);
int mm_partition_kway(const spmvb200_mm_s * mm, int32_t nparts, double ub, int32_t * part, int64_t * edgecut);
int mm_order_gp_kway(const spmvb200_mm_s * mm, int32_t nparts, int32_t * new_order);
void set_gp_partitioner(int64_t v);  // global option "mm.gp_partitioner"
int64_t gp_partitioner();

}  // namespace spmvb200
