/*
 * spmv_oracle.h -- CPU oracle for the SpMV hot path of jamtrott/spmv-cache-trace.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under spmv_cache_trace_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Every function is a plain-C restatement of one reference function; the
 * reference file:line it follows is cited on the declaration.  Parity status:
 * PINNED -- checked against the reference's own known-answer tests
 * (test/test_{csr,coo,ell,hybrid}-matrix.cpp), the poisson2D golden fixture
 * (test/poisson2D.hpp) and against the reference library itself compiled from
 * /root/reference (oracle/_ref, see oracle/Makefile and tests/test_oracle_vs_ref.py).
 *
 * All index types are int32_t like the reference (matrix/csr-matrix.hpp:15-16,
 * coo-matrix.hpp:16-17, ell-matrix.hpp:15-16, hybrid-matrix.hpp:17-18,
 * matrix-market.hpp:13-14).
 */
#ifndef SPMV_ORACLE_H
#define SPMV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OK = 0, ORC_ERR_PARSE = 1, ORC_ERR_FORMAT = 2, ORC_ERR_OVERFLOW = 3, ORC_ERR_NOMEM = 4 };

enum { ORC_FIELD_REAL = 0, ORC_FIELD_COMPLEX = 1, ORC_FIELD_INTEGER = 2, ORC_FIELD_PATTERN = 3 };
enum { ORC_FORMAT_COORDINATE = 0, ORC_FORMAT_ARRAY = 1 };
enum { ORC_SYM_GENERAL = 0, ORC_SYM_SYMMETRIC = 1, ORC_SYM_SKEW = 2, ORC_SYM_HERMITIAN = 3 };

/* Last error text of the calling thread ("" if none). */
const char *orc_last_error(void);

/* ---- Matrix Market (matrix/matrix-market.cpp:346-555) -------------------- */

/* Entries as the reference holds them after parsing: 1-based (i,j), value =
 * values_real() semantics (matrix-market.cpp:243-277): real -> a, complex ->
 * real part, integer -> (double) a, pattern -> 1.0.  No symmetry expansion
 * (the header's symmetry is parsed, :396-414, and then ignored). */
typedef struct {
    int32_t rows, columns, num_entries;
    int32_t format, field, symmetry;
    int32_t num_comments;
    int32_t *i, *j;
    double *a;
} orc_mm;

/* fromStream (matrix-market.cpp:530-555): header line, '%' comment lines,
 * size line, then num_entries whitespace-separated records. */
int orc_mm_parse(const char *text, size_t len, orc_mm *out);
void orc_mm_free(orc_mm *m);

/* Matrix::row_lengths / max_row_length (matrix-market.cpp:279-307). */
void orc_row_lengths(int32_t rows, int32_t n, const int32_t *i, int32_t *len);
int32_t orc_max_row_length(int32_t rows, int32_t n, const int32_t *i);

/* sort_matrix_row_major (matrix-market.cpp:897-929): order by (i, j).  The
 * reference uses std::sort (unstable) so duplicates have unspecified order;
 * this restatement is stable.  Inputs without duplicate (i,j) are identical. */
void orc_sort_row_major(int32_t n, int32_t *i, int32_t *j, double *a);
/* sort_matrix_column_major (matrix-market.cpp:863-895): order by (j, i). */
void orc_sort_column_major(int32_t n, int32_t *i, int32_t *j, double *a);

/* ---- CSR (matrix/csr-matrix.{hpp,cpp}, csr-matrix-spmv.cpp) -------------- */
typedef struct {
    int32_t rows, columns, num_entries, row_alignment;
    int32_t *row_ptr;      /* rows + 1 */
    int32_t *column_index; /* row_ptr[rows] (includes alignment padding) */
    double *value;         /* row_ptr[rows] */
} orc_csr;

/* from_matrix_market_row_aligned (csr-matrix.cpp:193-243); row_alignment 1 is
 * from_matrix_market (:187-191).  Input: unsorted 1-based entries. */
int orc_csr_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a,
                         int32_t row_alignment, orc_csr *out);
void orc_csr_free(orc_csr *m);
/* Matrix::size() = value_size + index_size (csr-matrix.cpp:46-60). */
size_t orc_csr_size(const orc_csr *m);
/* csr_spmv_inner_loop + csr_spmv (csr-matrix-spmv.cpp:21-33, 63-76): y += A x,
 * each row summed left to right into z = 0.0 and then added to y[i]. */
void orc_csr_spmv(const orc_csr *A, const double *x, double *y);
/* Same arithmetic, rows split over OpenMP threads like csr_matrix::spmv
 * (csr-matrix-spmv.cpp:148-167, chunk = ceil(rows/T)); for the CPU baseline. */
void orc_csr_spmv_omp(const orc_csr *A, const double *x, double *y, int num_threads);
/* Matrix::spmv_rows_per_thread / spmv_nonzeros_per_thread (csr-matrix.cpp:77-95):
 * the reference row partition. */
int32_t orc_csr_rows_per_thread(int32_t rows, int thread, int num_threads);
int32_t orc_csr_start_row(int32_t rows, int thread, int num_threads);
int32_t orc_csr_nonzeros_per_thread(const int32_t *row_ptr, int32_t rows, int thread, int num_threads);

/* ---- COO (matrix/coo-matrix.{hpp,cpp}) ------------------------------------ */
typedef struct {
    int32_t rows, columns, num_entries;
    int32_t *row_index, *column_index; /* 0-based, FILE ORDER (coo-matrix.cpp:226-239) */
    double *value;
} orc_coo;

int orc_coo_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a, orc_coo *out);
void orc_coo_free(orc_coo *m);
size_t orc_coo_size(const orc_coo *m); /* coo-matrix.cpp:49-63 */
/* coo_spmv (coo-matrix.cpp:248-285) with chunk = ceil(nnz/T) (:313-323).
 * T == 1: serial scatter.  T > 1: thread t = (k / chunk) % T accumulates into
 * workspace[t*rows + r]; then y[i] += sum_t workspace[t*rows+i], t ascending.
 * The workspace is NOT cleared (neither does the reference); pass it zeroed
 * for y += A x.  Executed sequentially -- the static schedule makes the
 * OpenMP result independent of timing, so this is the same arithmetic. */
void orc_coo_spmv(int num_threads, const orc_coo *A, const double *x, double *y,
                  double *workspace, int32_t chunk_size);
/* coo_spmv_atomic (coo-matrix.cpp:287-309).  With T > 1 the reference's order
 * of atomic updates is timing dependent; the oracle uses entry order. */
void orc_coo_spmv_atomic(const orc_coo *A, const double *x, double *y);
void orc_coo_spmv_omp(const orc_coo *A, const double *x, double *y, double *workspace, int num_threads);

/* ---- ELLPACK (matrix/ell-matrix.{hpp,cpp}) -------------------------------- */
typedef struct {
    int32_t rows, columns, num_entries, row_length;
    int32_t skip_padding;
    int32_t *column_index; /* rows*row_length, ROW-MAJOR (k = i*row_length + l, ell-matrix.cpp:254) */
    double *value;
} orc_ell;

/* from_matrix_market (ell-matrix.cpp:190-238): W = max row length; pad value
 * 0.0; pad column = column of the last real entry consumed so far
 * (column_indices[k-1]-1, :229) or INT32_MAX with skip_padding.  The reference
 * reads column_indices[-1] when the first row is empty (undefined behaviour);
 * the oracle writes 0 there and flags it via *first_row_empty. */
int orc_ell_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a,
                         int32_t skip_padding, orc_ell *out, int32_t *first_row_empty);
void orc_ell_free(orc_ell *m);
size_t orc_ell_size(const orc_ell *m); /* ell-matrix.cpp:52-65 */
/* ell_spmv_inner_loop / _skip_padding (ell-matrix.cpp:243-258, 275-292). */
void orc_ell_spmv(const orc_ell *A, const double *x, double *y);
void orc_ell_spmv_omp(const orc_ell *A, const double *x, double *y, int num_threads);

/* ---- Hybrid ELL+COO (matrix/hybrid-matrix.{hpp,cpp}) ---------------------- */
typedef struct {
    int32_t rows, columns, num_entries;
    int32_t ell_row_length, num_ell_entries, ell_skip_padding;
    int32_t *ell_column_index; /* rows*ell_row_length row-major */
    double *ell_value;
    int32_t num_coo_entries;
    int32_t *coo_row_index, *coo_column_index; /* row-major sorted order */
    double *coo_value;
} orc_hyb;

/* ELL width rule (hybrid-matrix.cpp:329-344): smallest L such that
 * #rows(len <= L) >= floor(2*rows/3). */
int32_t orc_hyb_ell_row_length(int32_t rows, const int32_t *row_lengths);
/* from_matrix_market (hybrid-matrix.cpp:316-417). */
int orc_hyb_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a,
                         int32_t skip_padding, orc_hyb *out);
void orc_hyb_free(orc_hyb *m);
/* Bytes of all stored arrays: 12*ell_slots + 16*coo_nnz.  NOTE the reference's
 * index_size() forgets coo_row_index (hybrid-matrix.cpp:80-86);
 * orc_hyb_size_reference() reproduces that, orc_hyb_size() does not. */
size_t orc_hyb_size(const orc_hyb *m);
size_t orc_hyb_size_reference(const orc_hyb *m);
/* hybrid_matrix::spmv (hybrid-matrix.cpp:535-567): ELL pass (:422-452) then
 * COO workspace pass (:491-528); both with chunk = ceil(rows/T) (:543-545). */
void orc_hyb_spmv(int num_threads, const orc_hyb *A, const double *x, double *y, double *workspace);
void orc_hyb_spmv_omp(const orc_hyb *A, const double *x, double *y, double *workspace, int num_threads);

/* ---- Tolerance bound named by BASELINE.json ------------------------------- */
/* bound[i] = sum_j |a_ij * x_j| over the stored entries of row i (CSR). */
void orc_csr_abs_rowsum(const orc_csr *A, const double *x, double *bound);

/* ---- Partitioners of the multi-GPU mode (oracle for OUR design) - */
/* P_ref: the reference rule, start_p = min(rows, p*ceil(rows/P)) (csr-matrix.cpp:77-83). */
void orc_partition_rows_ref(int64_t rows, int P, int64_t *starts /* P+1 */);
/* P_nnz: start_p = first row r with row_ptr[r] >= floor(p*nnz/P); start_0 = 0, start_P = rows. */
void orc_partition_rows_nnz(int64_t rows, const int64_t *row_ptr, int P, int64_t *starts /* P+1 */);

/* matrix-market-reorder.cpp:246-266: rows grouped by part; new_order[old] = new. */
int orc_order_from_parts(int32_t nvtxs, int32_t nparts, const int32_t *part, int32_t *new_order);
void orc_partition_rows_weighted(int64_t rows, const int64_t *row_ptr, int P, int64_t row_weight_q10, int64_t *starts /* P+1 */);

/* ---- R-MAT edges: C twin of OUR device generator (generators.cu), for the full-size parity tests ----------- */
void orc_rmat_keys(int scale, uint64_t seed, double a, double b, double c, uint64_t e0, uint64_t e1, uint64_t *keys);
void orc_rmat_unpack(int64_t n, const uint64_t *keys, int32_t *col, double *val);

#ifdef __cplusplus
}
#endif
#endif
