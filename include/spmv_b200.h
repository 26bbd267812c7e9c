/*
 * spmv_b200.h -- C ABI of the B200-native SpMV engine (libspmvb200.so).
 *
 * Drop-in boundary for the hot path of jamtrott/spmv-cache-trace: fp64
 * y += A*x for the COO, CSR, ELLPACK and hybrid (ELL+COO) formats.  Every
 * entry point names the reference interface it stands in for (file:line into
 * the reference tree).  Plain C types only: pointers, sizes, opaque handles.
 * No exceptions cross this boundary: every function returns 0 on success and
 * a non-zero spmvb200_status otherwise; spmvb200_last_error() gives the text.
 *
 * There is no CPU fallback.  Every compute entry point fails with
 * SPMVB200_ERR_CUDA when no CUDA device is usable.
 *
 * INTEGRATION.md shows the reference-side adapter (Kernel subclasses) that
 * binds these functions.
 */
#ifndef SPMV_B200_H
#define SPMV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPMVB200_VERSION 210 /* 0.2.0: + spmvb200_comm_*, spmvb200_dist_*, spmvb200_exchange_plan, spmvb200_partition_rows_weighted;
                                0.2.1: + spmvb200_mm_partition_kway, spmvb200_mm_order_gp_kway, spmvb200_order_from_parts */

typedef enum {
    SPMVB200_OK = 0,
    SPMVB200_ERR_INVALID = 1,  /* bad argument / handle                                   */
    SPMVB200_ERR_PARSE = 2,    /* Matrix Market syntax  (matrix::matrix_error in the ref) */
    SPMVB200_ERR_IO = 3,       /* file / gzip / tar     (matrix_error / std::system_error) */
    SPMVB200_ERR_OVERFLOW = 4, /* rows*row_length etc. (ell-matrix.cpp:201-205)           */
    SPMVB200_ERR_CUDA = 5,     /* CUDA runtime failure, or no device                      */
    SPMVB200_ERR_NOMEM = 6,
    SPMVB200_ERR_UNSUPPORTED = 7
} spmvb200_status;

typedef enum {
    SPMVB200_CSR = 0,
    SPMVB200_COO = 1,
    SPMVB200_ELL = 2,
    SPMVB200_HYB = 3
} spmvb200_format;

/* COO execution mode.  SEGMENTED: the builder keeps a row-sorted device copy
 * (stable, so the column order inside a row is the file order) and the kernel
 * is a segmented reduction -- the counterpart of coo_matrix::spmv
 * (coo-matrix.cpp:313-335).  ATOMIC: file order kept, one fp64 reduction per
 * entry -- the counterpart of coo_matrix::spmv_atomic (coo-matrix.cpp:337-358). */
typedef enum { SPMVB200_COO_SEGMENTED = 0, SPMVB200_COO_ATOMIC = 1 } spmvb200_coo_mode;

typedef enum {
    SPMVB200_STENCIL_2D5 = 0, /* 2D 5-point Poisson:  4 on the diagonal, -1 neighbours  */
    SPMVB200_STENCIL_3D7 = 1, /* 3D 7-point Poisson:  6 on the diagonal, -1 neighbours  */
    SPMVB200_STENCIL_3D27 = 2 /* 3D 27-point:        26 on the diagonal, -1 neighbours  */
} spmvb200_stencil;

typedef struct spmvb200_mm_s *spmvb200_mm_t;         /* host Matrix Market matrix  */
typedef struct spmvb200_matrix_s *spmvb200_matrix_t; /* device matrix + its x and y */

/* What Kernel::print emits (csr-spmv.cpp:97-112, hybrid-spmv.cpp:113-131) plus
 * the device-side facts.  matrix_size follows the reference's Matrix::size():
 * stored index + value bytes in the reference layout (csr-matrix.cpp:46-60,
 * coo-matrix.cpp:49-63, ell-matrix.cpp:52-65); for hybrid it is
 * 12*num_ell_entries + 16*num_coo_entries, i.e. WITH the coo_row_index bytes
 * that hybrid-matrix.cpp:80-86 forgets. */
typedef struct {
    int32_t format;          /* spmvb200_format                                  */
    int32_t coo_mode;        /* spmvb200_coo_mode (COO / HYB tail)               */
    int64_t rows, columns;
    int64_t num_entries;     /* true non-zeros ("nonzeros")                      */
    int64_t stored_entries;  /* CSR: row_ptr[rows]; ELL: rows*W; COO: nnz        */
    int64_t row_alignment;   /* CSR                                              */
    int64_t ell_row_length;  /* ELL / HYB                                        */
    int64_t num_ell_entries; /* HYB: rows*W                                      */
    int64_t num_coo_entries; /* HYB                                              */
    int32_t skip_padding;    /* ELL / HYB: padding column = INT32_MAX sentinel   */
    int32_t offsets_64bit;   /* device row_ptr is int64 (stored_entries >= 2^32) */
    int64_t matrix_size;     /* bytes, reference layout (see above)              */
    int64_t x_size, y_size;  /* 8*columns, 8*rows                                */
    int64_t device_bytes;    /* bytes actually allocated on the GPU              */
    int64_t row_offset;      /* first global row held (row-partitioned mode)     */
} spmvb200_info;

/* ---- errors, devices ----------------------------------------------------- */

/* Text of the last failure on the calling thread ("" if none).
 * Ref: what() of matrix::matrix_error (matrix-error.hpp:10-15) / kernel_error
 * (kernels/kernel.hpp:11-16). */
const char *spmvb200_last_error(void);
int spmvb200_version(void);
int spmvb200_device_count(int *count);
int spmvb200_set_device(int device);
/* sm_count, l2_bytes, total memory; name[] gets at most name_cap-1 chars. */
int spmvb200_device_props(int device, char *name, size_t name_cap, int *sm_count,
                          int64_t *l2_bytes, int64_t *mem_bytes, int *cc_major, int *cc_minor);
/* Total number of kernels this library has launched in this process. */
int64_t spmvb200_launch_count(void);
/* Process-wide switches.  "force_offsets64" = 1 makes every CSR built afterwards keep int64 row
 * offsets on the device (normally only when stored_entries >= 2^32) so that path can be tested.
 * "coo.col_block_log2": the builders of SEGMENTED COO matrices and of the hybrid tail may store the
 * row-sorted entries partitioned by column block (blocks of 2^k columns, rows ascending inside a block)
 * so that a block's slice of x stays in L2; 0 (default) = automatic (only when the x referenced is larger
 * than 0.75 L2, the gathers are not local already -- more than three column blocks touched per 2^20
 * consecutive entries -- and the extra sweeps over y cost less than the gather misses), -1 = never, k > 0 = always, with 2^k
 * columns.  Exports restore the row-major order.  spmvb200_get_option(m, "coo.col_block_log2") tells
 * what was applied to a matrix.
 * "mm.gp_partitioner" = 1: the graph-partitioning order ("__GP<n>", spmvb200_mm_order_gp) uses the library's own
 * K-way partitioner instead of being the identity (0, default: the reference's default build, which has no METIS). */
int spmvb200_set_global_option(const char *key, int64_t value);

/* ---- host side: Matrix Market --------------------------------------------- */

/* matrix_market::fromStream (matrix/matrix-market.cpp:530-555): header line,
 * '%' comment lines, size line, whitespace-separated coordinate records
 * (real, complex -> real part, integer, pattern -> 1.0; matrix-market.cpp:243-277).
 * Symmetry is parsed but entries are NOT expanded, like the reference. */
int spmvb200_mm_parse(const char *text, size_t len, spmvb200_mm_t *out);
/* matrix_market::load_matrix (matrix-market.cpp:777-861): .mtx, .gz, .tar.gz
 * and .tgz (member <name>/<name>.mtx, :755-757).  A "__RCM" path suffix loads
 * the file without the suffix and applies the reverse Cuthill-McKee order
 * (:786-802); "__GP<n>" is accepted and, like the reference built without
 * METIS, permutes nothing -- unless the global option "mm.gp_partitioner" is set
 * (spmvb200_mm_order_gp below). */
int spmvb200_mm_load(const char *path, spmvb200_mm_t *out);
/* matrix_market::Matrix(Header, Comments, Size, vector<CoordinateEntryReal>)
 * (matrix-market.hpp:81-84); i, j are 1-based. */
int spmvb200_mm_from_entries(int32_t rows, int32_t columns, int32_t num_entries,
                             const int32_t *i, const int32_t *j, const double *a, spmvb200_mm_t *out);
/* Matrix::rows/columns/num_entries/field/symmetry (matrix-market.hpp:104-113);
 * field: 0 real 1 complex 2 integer 3 pattern; symmetry: 0 general 1 symmetric
 * 2 skew-symmetric 3 hermitian; format: 0 coordinate 1 array. */
int spmvb200_mm_info(spmvb200_mm_t mm, int32_t *rows, int32_t *columns, int32_t *num_entries,
                     int32_t *field, int32_t *symmetry, int32_t *format);
/* Matrix::row_indices / column_indices / values_real (matrix-market.cpp:171-277);
 * borrowed pointers, valid until spmvb200_mm_free. */
int spmvb200_mm_entries(spmvb200_mm_t mm, const int32_t **i, const int32_t **j, const double **a);
/* Matrix::max_row_length / row_lengths (matrix-market.cpp:279-307). */
int spmvb200_mm_max_row_length(spmvb200_mm_t mm, int32_t *out);
int spmvb200_mm_row_lengths(spmvb200_mm_t mm, int32_t *lengths /* rows */);
/* sort_matrix_row_major / sort_matrix_column_major (matrix-market.cpp:863-929),
 * in place and stable. */
int spmvb200_mm_sort_row_major(spmvb200_mm_t mm);
int spmvb200_mm_sort_column_major(spmvb200_mm_t mm);
/* find_new_order_RCM (matrix/matrix-market-reorder.cpp:60-170): new_order[old index] = new index,
 * `rows` entries; square real coordinate matrices only. */
int spmvb200_mm_order_rcm(spmvb200_mm_t mm, int32_t *new_order);
/* find_new_order_GP (matrix-market-reorder.cpp:172-278).  The reference partitions the matrix graph with
 * METIS_PartGraphKway (third-party; absent from its default build and from this image) and groups the rows by
 * part; built without METIS it returns the identity (:172-180), and so does this call by default.  With the
 * global option "mm.gp_partitioner" = 1 (also honoured by the "__GP<n>" path suffix of spmvb200_mm_load) it is
 * spmvb200_mm_order_gp_kway: the library's own K-way partitioner in METIS's place -- a valid partition with the
 * same balance bound, but not the one METIS would return. */
int spmvb200_mm_order_gp(spmvb200_mm_t mm, int32_t nparts, int32_t *new_order);
int spmvb200_mm_order_gp_kway(spmvb200_mm_t mm, int32_t nparts /* <= 1: 16, :232-233 */, int32_t *new_order);
/* What METIS_PartGraphKway is called for (:236-237): part[v] in [0, nparts) for every row of a square real
 * coordinate matrix, unit vertex weights, no part larger than max(ceil(n/k), ub*n/k) rows (the reference passes
 * ub = 1.05 -> ub_permille 1050); the graph is the off-diagonal pattern made undirected.  Level-structure cut from
 * a pseudo-peripheral vertex per component, then boundary refinement; deterministic.  edgecut (may be NULL) =
 * undirected edges whose ends lie in different parts. */
int spmvb200_mm_partition_kway(spmvb200_mm_t mm, int32_t nparts, int32_t ub_permille, int32_t *part, int64_t *edgecut);
/* The second half of find_new_order_GP (:246-266), exact: rows grouped by part, parts ascending, the rows of a
 * part in ascending index order; new_order[old index] = new index. */
int spmvb200_order_from_parts(int32_t n, int32_t nparts, const int32_t *part, int32_t *new_order);
/* Matrix::permute (matrix-market.cpp:309-333): i, j <- new_order[i-1]+1, new_order[j-1]+1. */
int spmvb200_mm_permute(spmvb200_mm_t mm, const int32_t *new_order);
void spmvb200_mm_free(spmvb200_mm_t mm);

/* ---- device builders from Matrix Market entries ---------------------------- */
/* The conversions run on the GPU (sort, row pointer, padding, split) and leave
 * the matrix resident in the padded, aligned device layout the kernels want.
 * Each also allocates the kernel object's vectors, x = 1.0 and y = 0.0, as
 * Kernel::init does (csr-spmv.cpp:35-36). */

/* csr_matrix::from_matrix_market_row_aligned (matrix/csr-matrix.cpp:193-243);
 * row_alignment 1 == from_matrix_market (:187-191). */
int spmvb200_csr_from_mm(spmvb200_mm_t mm, int32_t row_alignment, spmvb200_matrix_t *out);
/* coo_matrix::from_matrix_market (matrix/coo-matrix.cpp:220-243). */
int spmvb200_coo_from_mm(spmvb200_mm_t mm, int32_t coo_mode, spmvb200_matrix_t *out);
/* ell_matrix::from_matrix_market (matrix/ell-matrix.cpp:190-238). */
int spmvb200_ell_from_mm(spmvb200_mm_t mm, int32_t skip_padding, spmvb200_matrix_t *out);
/* hybrid_matrix::from_matrix_market (matrix/hybrid-matrix.cpp:316-417). */
int spmvb200_hyb_from_mm(spmvb200_mm_t mm, int32_t skip_padding, spmvb200_matrix_t *out);

/* ---- device matrices from arrays the caller already holds ------------------ */
/* This is what a reference-side Kernel adapter calls from prepare(): it owns a
 * converted host matrix and hands its arrays over once. */

/* csr_matrix::Matrix{rows, columns, num_entries, row_ptr, column_index, value}
 * (matrix/csr-matrix.hpp:22-65).  stored = row_ptr[rows]. */
int spmvb200_csr_create(int32_t rows, int32_t columns, int32_t num_entries,
                        const int32_t *row_ptr, const int32_t *column_index, const double *value,
                        spmvb200_matrix_t *out);
/* Same with 64-bit offsets, for matrices beyond the reference's int32 limit. */
int spmvb200_csr_create64(int64_t rows, int64_t columns, int64_t num_entries,
                          const int64_t *row_ptr, const int32_t *column_index, const double *value,
                          spmvb200_matrix_t *out);
/* coo_matrix::Matrix (matrix/coo-matrix.hpp:22-70); 0-based indices. */
int spmvb200_coo_create(int32_t rows, int32_t columns, int64_t num_entries,
                        const int32_t *row_index, const int32_t *column_index, const double *value,
                        int32_t coo_mode, spmvb200_matrix_t *out);
/* ell_matrix::Matrix (matrix/ell-matrix.hpp:22-65): ROW-MAJOR arrays of
 * rows*row_length entries; re-laid out column-major on the device. */
int spmvb200_ell_create(int32_t rows, int32_t columns, int32_t num_entries, int32_t row_length,
                        const int32_t *column_index, const double *value, int32_t skip_padding,
                        spmvb200_matrix_t *out);
/* hybrid_matrix::Matrix (matrix/hybrid-matrix.hpp:24-96). */
int spmvb200_hyb_create(int32_t rows, int32_t columns, int32_t num_entries,
                        int32_t ell_row_length, const int32_t *ell_column_index, const double *ell_value,
                        int32_t ell_skip_padding, int32_t num_coo_entries,
                        const int32_t *coo_row_index, const int32_t *coo_column_index,
                        const double *coo_value, spmvb200_matrix_t *out);

/* ---- synthetic matrices generated on the device (BASELINE.json configs) ---- */

/* Rows [row_begin, row_end) of the nx*ny*nz stencil matrix (nz = 1 for 2D), in
 * row-major (x fastest) grid order, global column indices, columns sorted inside
 * each row.  format: CSR, ELL, COO or HYB.  The matrix keeps `columns` = nx*ny*nz
 * so x is full length; y holds row_end-row_begin entries. */
int spmvb200_gen_stencil(int32_t kind, int64_t nx, int64_t ny, int64_t nz,
                         int64_t row_begin, int64_t row_end, int32_t format, spmvb200_matrix_t *out);
/* R-MAT power-law matrix: 2^scale rows/columns, edge_factor*2^scale edge draws
 * with quadrant probabilities (a, b, c, 1-a-b-c), counter-based RNG keyed by
 * (seed, edge id), duplicates removed, value = hash(i, j) in (-1, 1).
 * Rows [row_begin, row_end) are kept (0, 0 = all). */
int spmvb200_gen_rmat(int32_t scale, int32_t edge_factor, uint64_t seed, double a, double b, double c,
                      int64_t row_begin, int64_t row_end, int32_t format, int32_t coo_mode,
                      spmvb200_matrix_t *out);
/* Device-side format change following the reference's conversion rules
 * (CSR source only): arg = skip_padding for ELL/HYB, coo_mode for COO. */
int spmvb200_convert(spmvb200_matrix_t src, int32_t format, int32_t arg, spmvb200_matrix_t *out);

/* ---- inspection / export (reference layout) -------------------------------- */

int spmvb200_matrix_info(spmvb200_matrix_t m, spmvb200_info *info);
/* Copy the matrix back in the reference's host layout (for bit-exact
 * conversion parity and for adapters that want the arrays).  NULL = skip. */
int spmvb200_csr_export(spmvb200_matrix_t m, int64_t *row_ptr, int32_t *column_index, double *value);
int spmvb200_coo_export(spmvb200_matrix_t m, int32_t *row_index, int32_t *column_index, double *value);
int spmvb200_ell_export(spmvb200_matrix_t m, int32_t *column_index_row_major, double *value_row_major);
int spmvb200_hyb_export(spmvb200_matrix_t m, int32_t *ell_column_index, double *ell_value,
                        int32_t *coo_row_index, int32_t *coo_column_index, double *coo_value);

/* ---- vectors (the Kernel object's x and y, csr-spmv.hpp:36-37) ------------- */

int spmvb200_set_x(spmvb200_matrix_t m, const double *x_host);
int spmvb200_set_y(spmvb200_matrix_t m, const double *y_host);
int spmvb200_get_x(spmvb200_matrix_t m, double *x_host);
int spmvb200_get_y(spmvb200_matrix_t m, double *y_host);
int spmvb200_fill_x(spmvb200_matrix_t m, double value);
int spmvb200_fill_y(spmvb200_matrix_t m, double value);
/* Device pointers of x / y, and binding of caller-owned device buffers (e.g.
 * the all-gathered x of the row-partitioned mode).  Bound buffers are not freed. */
int spmvb200_x_device(spmvb200_matrix_t m, void **ptr);
int spmvb200_y_device(spmvb200_matrix_t m, void **ptr);
int spmvb200_bind_x(spmvb200_matrix_t m, void *device_ptr);
int spmvb200_bind_y(spmvb200_matrix_t m, void *device_ptr);
/* cudaStream_t to launch on (default: a stream owned by the matrix). */
int spmvb200_set_stream(spmvb200_matrix_t m, void *cuda_stream);
/* Pinned host memory for x / y staging of spmvb200_spmv_host. */
int spmvb200_host_alloc(size_t bytes, void **ptr);
int spmvb200_host_free(void *ptr);

/* ---- the hot path ----------------------------------------------------------- */

/* Build whatever launch metadata the selected kernel needs (CSR: span table / tile table) now instead
 * of on the first spmvb200_spmv, and wait for it.  Ref: Kernel::prepare (kernels/kernel.hpp:28), which
 * profile_kernel_run calls once before the timed runs (profile-kernel.cpp:227). */
int spmvb200_prepare(spmvb200_matrix_t m);
/* y += A*x on the device, asynchronous on the matrix's stream.
 * Ref: csr_matrix::spmv (matrix/csr-matrix-spmv.cpp:148-167),
 *      coo_matrix::spmv / spmv_atomic (matrix/coo-matrix.cpp:313-358),
 *      ell_matrix::spmv (matrix/ell-matrix.cpp:311-335),
 *      hybrid_matrix::spmv (matrix/hybrid-matrix.cpp:535-567),
 * i.e. the bodies of Kernel::run (kernels/csr-spmv.cpp:64-67 etc.). */
int spmvb200_spmv(spmvb200_matrix_t m);
/* y += alpha*A*x from now on (default 1.0, which is exact and the reference's semantics).  With the
 * "beta0" option: y = alpha*A*x, which kernels that own whole rows (ELL, sliced CSR) write with plain
 * stores -- no clearing pass, no read of y.  Supported by the default kernels of every format. */
int spmvb200_set_alpha(spmvb200_matrix_t m, double alpha);
/* Wait for the matrix's stream. */
int spmvb200_sync(spmvb200_matrix_t m);
/* Host-buffer form: copies x (columns) and y (rows) to the device, runs
 * y += A*x, copies y back, synchronises.  This is the end-to-end call. */
int spmvb200_spmv_host(spmvb200_matrix_t m, const double *x_host, double *y_host);
/* `reps` launches after `warmup` untimed ones, each bracketed by CUDA events on
 * the launching stream; ms[r] = duration of launch r.  The protocol of
 * profile_kernel_run (profile-kernel.cpp:137-179) with events for the clock. */
int spmvb200_time(spmvb200_matrix_t m, int warmup, int reps, float *ms);
/* Benchmark form of the same protocol for operands that fit in L2: the n matrices are copies of
 * one workload (each with its own x and y); `steps` launches go round-robin over them on the
 * stream of ms[0] so that no launch finds its operands in L2.  total_ms = CUDA-event time around
 * the whole sequence of `steps` launches; per_launch_ms (may be NULL; steps entries) = event time
 * of each launch, taken in a second, identical pass so the events do not perturb total_ms. */
int spmvb200_time_rotating(const spmvb200_matrix_t *ms, int n, int warmup, int steps,
                           float *total_ms, float *per_launch_ms);
/* End-to-end variant: every step is spmvb200_spmv_host(ms[k], xs[k], ys[k]) (H2D x and y, kernel,
 * D2H y, synchronise), k round-robin; total_ms = CUDA-event time around all `steps` steps. */
int spmvb200_time_host_rotating(const spmvb200_matrix_t *ms, int n, const double *const *xs,
                                double *const *ys, int warmup, int steps, float *total_ms);
/* Yardstick for the isolated-launch figures: the event-pair protocol of spmvb200_time around a plain device-to-device
 * cudaMemcpyAsync of `bytes` (2 * bytes of DRAM traffic), rotating over `copies` buffer pairs (L2-cold); ms[r] per copy. */
int spmvb200_time_copy(int64_t bytes, int copies, int warmup, int reps, float *ms);
/* Options (0 = automatic unless noted):
 *   semantics   "beta0" 1: y = A*x instead of y += A*x.
 *               "independent_launches": y is updated with reductions, so two launches need ordering only
 *               when one writes what the other reads.  On the stream a matrix owns the library tracks
 *               the x/y ranges of the kernels in flight and lets a launch start while the previous one
 *               drains whenever that is provably safe (the reference protocol: x constant, y
 *               accumulated); any other call on the matrix, overlapping ranges or "beta0" restore full
 *               ordering.  On a caller-provided stream (spmvb200_set_stream) nothing is assumed unless
 *               this option is 1 = the caller promises that no kernel in flight writes this matrix's x.
 *               -1 = never overlap.
 *               "pdl" (default 1): programmatic dependent launch, used on the stream the matrix owns; on a
 *               caller-provided stream (which the caller may have tied to other streams with events) launches
 *               are plain unless "pdl" = 2; 0 = never.
 *   CSR         "csr.algo" 1 stream/direct, 2 stream/product, 3 warp-granular, 4 flat (split by non-zeros,
 *               "csr.entries" 4|8 per lane, rows from span metadata; "csr.rowptr_path" 1 = matrices with empty
 *               rows rebuild the row numbers from row_ptr instead of the row map), 5 sliced (lane per row on a slot-major copy of
 *               the entries; "csr.batch" 2|4|8 slots in flight; "csr.drop_row_major" (default 1) = free the row-major
 *               column_index/value once that copy exists, so the matrix is resident once -- they are rebuilt on
 *               demand by export, convert, row_block, column_span and the other kernels; 0 = keep both copies;
 *               "csr.index_runs": the copy's column stream is stored by DIAGONAL where a 32-row slice's entries lie on few
 *               diagonals -- one {base, mask} descriptor per slot (offset = column - row) instead of up to 32 column
 *               indices, no row_ptr read; values and summation order untouched, results bit-identical -- 0 = when that
 *               shrinks the stream to <= 3/4 (banded matrices), 1 = always, -1 = never; changing it rebuilds the copy;
 *               "csr.regs" 40: (experiment) that kernel with a 40-register budget); "csr.rmw" (sliced kernel, y += A*x): the lane that owns a row adds to y
 *               with a plain load and store instead of a reduction -- the launches are then ordered; 1 = on (default off:
 *               measured slower); "csr.probe" 1 = regular traffic
 *               (y_i += sum a_k, values streamed, no gather), 2 = irregular traffic (y_i += sum x[j_k], the
 *               gather alone): spmv_regular_traffic / spmv_irregular_traffic of the reference
 *               (csr-matrix-spmv.cpp:35-61, 119-146); "csr.algo" 0 = automatic: sliced when the mean row has
 *               >= 10 entries and the longest row <= 2x the mean, else flat; "csr.threads" 128|256 (algo 3, 4),
 *               32..256 (algo 1, 2); algo 1-3: "csr.lanes" 1|2|4|8 lanes per row; algo 1, 2: "csr.tile"
 *               256..2048, "csr.stages" 2|3, "csr.ctas_per_sm", "csr.spare_ctas" CTA slots per SM left
 *               free for a concurrent kernel.
 *   ELL         "ell.rows_per_thread" 1|2|4, "ell.block" 32..256.
 *   COO         "coo.algo" 1 shared-memory staged tiles (sorted entries), 2 register-staged, 4 entries per
 *               lane, one segmented warp scan per 128 entries (default, any entry order), 3 one reduction
 *               per entry, 4 register-staged, striped lanes; "coo.items" 2|4|8 stripes per warp (algo 4);
 *               "coo.threads" 64|128|256; "coo.stages" 2|3|4 and "coo.ctas_per_sm" (algo 1).
 *               "coo.hot" 1: hot-column kernel (row-sorted entries): the entries are cut into segments, every segment's
 *               most referenced columns ("coo.hot_slots", default 24576) are gathered from shared memory by a
 *               persistent grid ("coo.hot_threads" 256|512|1024, "coo.hot_entries" 4|8 per lane, "coo.hot_segments"
 *               per CTA); off by default: measured slower than the plain kernel.  Experiment switches of
 *               the plain kernel: "coo.xload" 0 ld.global.nc | 1 ld.global.cg | 2 nc L1::no_allocate | 3 nc L1::evict_last
 *               | 4 cp.async through shared memory | 5 texture fetch | 6 texture + read-only path, half each;
 *               "coo.carveout" preferred shared-memory carve-out in percent.
 *   host path   "host.zero_copy" (default 1): ELL: spmvb200_spmv_host lets the kernel read and write
 *               pinned host y directly; 2 = y up by DMA in "host.chunks" row chunks, results stored by
 *               the kernel; 0 = copies only ("host.chunks" > 1 pipelines them).  CSR (sliced kernel,
 *               square matrices of >= 2^20 rows, x uploaded in "host.chunks" pieces and every row chunk
 *               launched when the columns it references have arrived): 1 = automatic = 3; 3 = y_old up
 *               and y_new down by the copy engines, chunk by chunk; 2 = y_old up by DMA, y_new stored
 *               into the pinned buffer by the kernel; 4 = the kernel reads y_old from and stores y_new
 *               to the pinned buffer. */
int spmvb200_set_option(spmvb200_matrix_t m, const char *key, int64_t value);
/* Also answers the read-only keys "csr.index_runs_active" (the sliced kernel's copy is stored by diagonal) and
 * "csr.index_columns_stored" (int32 entries of its column stream; = stored entries when plain),
 * "coo.col_block_log2" (what the builder applied), "coo.hot_coverage_permille" and
 * "coo.hot_segments_built" (the hot-column tables, if built), and
 * "last_launch.overlapped" / "last_launch.pdl" (how the library ordered the last kernel of this matrix). */
int spmvb200_get_option(spmvb200_matrix_t m, const char *key, int64_t *value);
/* Name of the kernel spmvb200_spmv launches for this matrix. */
const char *spmvb200_kernel_name(spmvb200_matrix_t m);

int spmvb200_destroy(spmvb200_matrix_t m);

/* ---- row partition of the multi-GPU mode ------------------------------------ */

/* The reference rule (matrix/csr-matrix.cpp:77-83): start_p = min(rows, p*ceil(rows/P)). */
int spmvb200_partition_rows_ref(int64_t rows, int32_t parts, int64_t *starts /* parts+1 */);
/* Balanced non-zeros: start_p = first row r with row_ptr[r] >= floor(p*nnz/P). CSR only. */
int spmvb200_partition_rows_nnz(spmvb200_matrix_t m, int32_t parts, int64_t *starts /* parts+1 */);
/* Balanced COST, where a row costs its entries plus row_weight_q10/1024 entries of per-row work (its y traffic, its ELL
 * padding, one reduction per run -- what makes a block of many short rows slower than a block of few long ones with the
 * same non-zeros): start_p = first row r with 1024*row_ptr[r] + w*r >= floor(p*(1024*nnz + w*rows)/P).  w = 0 is the
 * cut of spmvb200_partition_rows_nnz (up to one row: the targets round differently); a very large w tends to equal rows.
 * CSR only.  (Measured on configs[3] at two GPUs, profiles/r02_sweep_s_c4_row_weight.log: w = 0 / 4 / 12 / 32 -> 4.90 / 4.90 /
 * 5.15 / 6.06 ms: equal non-zeros IS the balanced cut for that matrix.) */
int spmvb200_partition_rows_weighted(spmvb200_matrix_t m, int32_t parts, int64_t row_weight_q10, int64_t *starts /* parts+1 */);
/* New matrix holding rows [row_begin, row_end) of a CSR matrix (global columns). */
int spmvb200_csr_row_block(spmvb200_matrix_t m, int64_t row_begin, int64_t row_end,
                           spmvb200_matrix_t *out);
/* Two new CSR matrices with the rows of `m`: the entries whose column lies in [col_begin, col_end) and all the
 * others (A = inside + outside).  The row-partitioned mode uses it for matrices that are not banded: the part of a
 * rank's rows that references only the rank's own slice of x runs while the exchange of x is in flight. */
int spmvb200_csr_column_split(spmvb200_matrix_t m, int64_t col_begin, int64_t col_end, spmvb200_matrix_t *inside,
                              spmvb200_matrix_t *outside);
/* What a rank of the row-partitioned mode needs from the others.  For a CSR matrix whose rows own
 * columns [col_begin, col_end) of x (the analogue of the reference tagging every x[j] with the
 * thread that owns its page, matrix/csr-matrix.cpp:132-136):
 *   col_min, col_max   smallest / largest column index referenced by any row (-1 if no entries);
 *   lo_end             1 + last row that references a column <  col_begin (0 if none);
 *   hi_begin           first row that references a column >= col_end (rows if none).
 * Rows [lo_end, hi_begin) -- if that range is not empty -- reference only the rank's own columns
 * and can run while the exchange of x is still in flight. */
int spmvb200_csr_column_span(spmvb200_matrix_t m, int64_t col_begin, int64_t col_end,
                             int64_t *col_min, int64_t *col_max, int64_t *lo_end, int64_t *hi_begin);

/* ---- the row-partitioned multi-GPU mode (one rank per GPU) ---------------------- */
/* The reference is one process: thread t of its OpenMP team owns rows [t*ceil(rows/T), ...) (matrix/csr-matrix.cpp:77-95)
 * and runs Kernel::run together with the others between barriers (profile-kernel.cpp:159-161).  Here a RANK owns a row
 * block on its own GPU, and the iteration x_(k+1) = alpha*A*x_k needs one exchange of x per step.  A communicator
 * connects the ranks; two kinds exist behind the same handle:
 *   - in-process ("local"): all ranks live in one process (one host thread may drive them all, or one thread per rank,
 *     e.g. the OpenMP team of profile_kernel with a barrier between steps); x moves by direct peer copies over NVLink
 *     (cudaMemcpyPeerAsync), ordered by CUDA events between the ranks' streams.  Several ranks may share a GPU,
 *     which is how the single-GPU tests exercise the multi-rank logic.
 *   - NCCL: one process per rank (torchrun, MPI, ...): the caller carries the 128-byte id from rank 0 to the others.
 *     NCCL is loaded at run time (dlopen of libnccl.so.2); the library has no link-time dependency on it. */
typedef struct spmvb200_comm_s *spmvb200_comm_t;
typedef struct spmvb200_dist_s *spmvb200_dist_t;
#define SPMVB200_COMM_ID_BYTES 128

/* In-process communicator: comms[r] is rank r's handle, living on devices[r] (NULL: device r % device_count). */
int spmvb200_comm_create_local(int nranks, const int *devices, spmvb200_comm_t *comms /* nranks */);
/* NCCL communicator.  Rank 0 calls _unique_id and ships the bytes to the other ranks by any means; then EVERY rank
 * calls _create_nccl (collective; the calling thread's current device is the rank's GPU). */
int spmvb200_comm_unique_id(void *id /* SPMVB200_COMM_ID_BYTES */);
int spmvb200_comm_create_nccl(const void *id, int rank, int nranks, spmvb200_comm_t *out);
int spmvb200_comm_rank(spmvb200_comm_t c, int *rank, int *nranks, int *device);
/* Collectives for the plumbing around a measurement: barrier; max / sum of one double over the ranks.
 * In-process communicators: call from one thread per rank, or for rank 0..n-1 in order from one thread. */
int spmvb200_comm_barrier(spmvb200_comm_t c);
int spmvb200_comm_allreduce(spmvb200_comm_t c, double *value, int op /* 0 max, 1 sum, 2 min */);
int spmvb200_comm_destroy(spmvb200_comm_t c);

/* How x travels between steps. */
typedef enum {
    SPMVB200_EXCHANGE_AUTO = 0,      /* halo when every rank needs at most a quarter of x from the others */
    SPMVB200_EXCHANGE_ALLGATHER = 1, /* every rank receives every other slice (north star: "NCCL all-gather of x") */
    SPMVB200_EXCHANGE_HALO = 2       /* every rank receives exactly the column range its rows reference */
} spmvb200_exchange;
/* spmvb200_dist_create flags */
#define SPMVB200_DIST_CONSUME_LOCAL 1 /* the executor may destroy `local` once its row blocks exist (halves the footprint) */
#define SPMVB200_DIST_COLUMN_SPLIT 2  /* cut the block by COLUMNS (own slice of x / the rest) instead of by rows: for
                                         matrices without a band (power law); all-gather exchange */
#define SPMVB200_DIST_NO_OVERLAP 4    /* one block per rank, run after the exchange */
#define SPMVB200_DIST_PEER_COPY 8     /* NCCL communicators: move x with copy-engine pulls out of the owners' IPC-mapped
                                         buffers over NVLink instead of NCCL kernels (no SM taken from the SpMV, no
                                         rendezvous); spmvb200_dist_set_x and _destroy become collective */

/* The exchange plan, as plain arithmetic (no device, no communicator): rank `rank` of `parts` owns columns
 * [starts[rank], starts[rank+1]) of x; need_lo/need_hi[q] = the column range rank q's rows reference (hi exclusive).
 * Returns the mode chosen and up to `cap` (peer, lo, hi) triples for the sends and the receives of `rank`
 * (all-gather: none listed; recv_bytes = 8 * (n - own)). */
int spmvb200_exchange_plan(int32_t parts, const int64_t *starts, const int64_t *need_lo, const int64_t *need_hi,
                           int32_t rank, int32_t mode, int32_t cap, int32_t *chosen_mode,
                           int32_t *n_sends, int64_t *sends /* 3*cap */, int32_t *n_recvs, int64_t *recvs /* 3*cap */,
                           int64_t *recv_bytes);

/* Executor of x_(k+1) = alpha*A*x_k for this rank.  `local`: the rank's rows [starts[rank], starts[rank+1]) with GLOBAL
 * column indices (spmvb200_gen_stencil(..., row_begin, row_end, ...), spmvb200_csr_row_block, spmvb200_csr_create64), CSR;
 * ELL / COO / hybrid blocks are accepted with the all-gather exchange.  `format`: format the pieces are converted to
 * before they run (SPMVB200_CSR = leave as is; SPMVB200_HYB for BASELINE configs[3]).  Collective over the communicator.
 * The rank's rows are cut into blocks: rows that reference only the rank's own slice of x run on one stream WHILE the
 * exchange is in flight on a second (high-priority) stream; rows that need remote x run on a third stream as soon as the
 * exchange has delivered it -- concurrently with the first block, so the step costs max(interior, exchange + boundary).
 * x lives in two full-length device buffers used in ping-pong; step k reads X_k and writes its rows of alpha*A*X_k into
 * its slice of X_(k+1) with plain stores ("beta0"), so there is no y -> x copy, no clearing pass and no scaling pass. */
int spmvb200_dist_create(spmvb200_comm_t comm, spmvb200_matrix_t local, const int64_t *starts /* nranks+1 */,
                         int32_t exchange, int32_t format, int32_t flags, spmvb200_dist_t *out);
/* The rank's slice of the CURRENT x: host -> device, device -> host, device pointer. */
int spmvb200_dist_set_x(spmvb200_dist_t d, const double *x_local_host);
int spmvb200_dist_get_x(spmvb200_dist_t d, double *x_local_host);
int spmvb200_dist_x_device(spmvb200_dist_t d, void **ptr);
/* One step, asynchronous.  In-process communicators: issue step k of every rank before step k+1 of any. */
int spmvb200_dist_step(spmvb200_dist_t d, double alpha);
int spmvb200_dist_sync(spmvb200_dist_t d);
/* `steps` steps of every executor in ds[0..n), bracketed by CUDA events on each rank's streams, after `warmup` untimed
 * ones; ms[r] = device time of ds[r].  The caller takes the maximum over the ranks.  NCCL: n = 1 (the process's own
 * rank; the ranks meet at a barrier before the timed steps).  In-process: all ranks of the communicator, driven by the
 * calling thread. */
int spmvb200_dist_time(const spmvb200_dist_t *ds, int n, int warmup, int steps, double alpha, float *ms /* n */);
/* End to end: `steps` INDEPENDENT products y_i = alpha*A*x_i, the rank's slice of every x_i coming from and the rank's
 * rows of every y_i going to (pinned) host memory.  Pipelined: the upload of step i+1 and the download of step i-1 run on
 * copy streams while step i computes.  ms (may be NULL) = device time of the whole sequence.  Synchronises. */
int spmvb200_dist_run_host(spmvb200_dist_t d, int steps, const double *const *x_local_host, double *const *y_local_host,
                           double alpha, float *ms);
#define SPMVB200_DIST_PEER_PUSH 16    /* PEER_COPY plus the fused form of the halo plan: the kernel that computes the rows a
                                         neighbouring rank references stores them into that rank's x buffer as well (peer-
                                         mapped pointer, the stores travel over NVLink), so the next exchange has nothing to
                                         copy; used when every rank's sending blocks run the sliced CSR kernel */
typedef struct {
    int32_t rank, nranks;
    int32_t exchange;             /* the mode in use (spmvb200_exchange)                         */
    int32_t n_blocks;             /* kernels' row blocks / column pieces                          */
    int32_t n_sends, n_recvs;
    int64_t recv_bytes_per_step;  /* bytes of x this rank receives per step                       */
    int64_t send_bytes_per_step;
    int64_t rows, row_begin;
    int64_t num_entries;          /* non-zeros of the rank's rows                                 */
    int64_t interior_rows;        /* rows that run during the exchange                            */
    int64_t device_bytes;         /* matrices + x buffers resident on this rank's GPU             */
    int64_t launches_per_step;
    int64_t steps_done;
    int64_t halo_push;            /* 1: the halo is stored by the senders' kernels (SPMVB200_DIST_PEER_PUSH in effect) */
} spmvb200_dist_info_t;
int spmvb200_dist_info(spmvb200_dist_t d, spmvb200_dist_info_t *info);
/* Block b of the executor (0 <= b < n_blocks): its row range inside the rank, whether it needs remote x, and the
 * kernel it runs. */
int spmvb200_dist_block(spmvb200_dist_t d, int32_t b, int64_t *row_begin, int64_t *row_end, int32_t *needs_remote_x,
                        spmvb200_matrix_t *matrix);
int spmvb200_dist_destroy(spmvb200_dist_t d);

/* ---- the reference's cache model for the chosen partition (host side) -------- */

/* The fully associative LRU model of the reference (cache-simulation/lru.cpp:31-54, driver
 * cache-trace.cpp:92-161, interleaving replacement.cpp:41-95) run over the SpMV reference string
 * (csr-matrix.cpp:97-143, ell-matrix.cpp:103-143, coo-matrix.cpp:144-185) with the two things the
 * GPU path needs and the reference lacks: misses attributed to the array they hit, and an
 * arbitrary contiguous partition.  Pure host code; no CUDA call. */
typedef struct {
    int64_t cache_bytes;    /* e.g. the L2 size reported by spmvb200_device_props                 */
    int32_t line_bytes;     /* 64 = the reference's configs; 128 = L2 line; 32 = DRAM sector         */
    int32_t parts;          /* "threads" of the reference model: GPUs of the row-partitioned mode   */
    const int64_t *starts;  /* parts+1 row (COO: entry) starts; NULL = the reference rule ceil(n/P) */
    int32_t shared;         /* 1: one cache, references of the parts interleaved round-robin
                               (the reference's shared L3); 0: a private cache per part (an L2 per GPU) */
    int32_t warmup;         /* 1: a warm-up pass before the counted one (cache-trace.cpp:128-140)  */
    int32_t page_bytes;     /* > 0: x_j / y_i are owned per page by the reference's rule
                               (aligned-allocator.hpp:156-211); 0: by the part that owns index j   */
    int32_t stream_bypass;  /* 1: index/value streams miss but are not allocated (the kernels'
                               evict-first L2 policy); 0: plain LRU for every reference             */
} spmvb200_cache_config;

typedef struct {
    int64_t references;
    int64_t misses_index;         /* row_ptr (CSR) / row_index (COO)                              */
    int64_t misses_column_index;
    int64_t misses_value;
    int64_t misses_x_local, misses_x_remote; /* the x gather, by owner of x_j                      */
    int64_t misses_y_local, misses_y_remote;
    int64_t x_references, x_remote_references;
} spmvb200_cache_misses;

/* out[parts].  Host arrays in the reference's layouts (ELL: row-major). */
int spmvb200_cache_trace_csr(int64_t rows, int64_t columns, const int64_t *row_ptr, const int32_t *column_index,
                             const spmvb200_cache_config *cfg, spmvb200_cache_misses *out);
int spmvb200_cache_trace_ell(int64_t rows, int64_t columns, int64_t row_length, const int32_t *column_index_row_major,
                             const spmvb200_cache_config *cfg, spmvb200_cache_misses *out);
int spmvb200_cache_trace_coo(int64_t rows, int64_t columns, int64_t num_entries, const int32_t *row_index,
                             const int32_t *column_index, const spmvb200_cache_config *cfg, spmvb200_cache_misses *out);
/* The same for a device matrix (CSR, ELL, COO; hybrid: its ELL part then its COO part, misses summed):
 * the index arrays are copied back to the host first.  Ref: Kernel::memory_reference_string
 * (kernels/kernel.hpp:33-36) + trace_cache_misses (cache-trace.cpp:163-187). */
int spmvb200_cache_trace(spmvb200_matrix_t m, const spmvb200_cache_config *cfg, spmvb200_cache_misses *out);

#ifdef __cplusplus
}
#endif
#endif
