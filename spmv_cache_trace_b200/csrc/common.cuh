// common.cuh -- internal declarations shared by the translation units of libspmvb200.so.
//
// Device layouts (all allocations 256 B aligned by cudaMalloc; streamed arrays carry >= 8192 zeroed
// elements of slack so vector loads and bulk copies of a last partial span stay inside the allocation):
//   CSR  row_ptr uint32[rows+1] (int64 when stored >= 2^32), column_index int32[], value f64[], plus the
//        launch metadata of the selected kernel, built by spmvb200_prepare or the first launch:
//        flat_meta  {first row, -, -, -, E mask words} per 32*E stored entries (flat kernel, default);
//        slice_col / slice_val  a second copy of the entries with every 32-row slice stored slot-major
//                   (sliced kernel, long regular rows; the row-major copy may then be dropped);
//        tile_row / span_row    tile and span tables of the first-generation stream / warp kernels.
//   ELL  COLUMN-MAJOR: slot l of row i lives at l*pitch + i (pitch = rows rounded up to 32), so
//        the 32 lanes of a warp read 32 consecutive rows of one slot with 64..256-bit loads.
//        (The reference layout is row-major, k = i*row_length + l, ell-matrix.cpp:254.)
//   COO  row_index int32[], column_index int32[], value f64[]; SEGMENTED mode stores them stably sorted
//        by row and, when x is much larger than L2, partitioned by column block (coo_col_shift).
//   HYB  the ELL arrays plus the COO arrays of the tail (row-major sorted as the reference builds them,
//        then possibly column-blocked like COO).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/spmv_b200.h"

namespace spmvb200 {

// ---- error plumbing ----------------------------------------------------------------------
void set_error(const std::string & msg);
int fail(int code, const std::string & msg);
int cuda_fail(cudaError_t e, const char * what, const char * file, int line);
void count_launch(int n = 1);

#define SPMV_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) return ::spmvb200::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define SPMV_TRY(call)                 \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

// ---- tiling constants ----------------------------------------------------------------------
constexpr int kCsrThreads = 256;
constexpr int kCooItems = 7;  // odd: blocked per-thread walks over shared memory stay bank-conflict free
constexpr int kCooThreads = 256;
constexpr int kCooTile = kCooThreads * kCooItems;  // 1792 entries = 28 KiB per stage
constexpr int64_t kPadEntries = 8192;              // slack after every streamed array (>= largest tile)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

}  // namespace spmvb200

// The opaque handle of the C ABI.
struct spmvb200_matrix_s {
    int32_t format = 0, coo_mode = 0;
    int64_t rows = 0, cols = 0, nnz = 0, stored = 0, row_alignment = 1;
    int64_t ell_w = 0, n_ell = 0, n_coo = 0;
    int32_t skip_padding = 0;
    bool off64 = false;
    int64_t row_offset = 0;
    int device = 0;

    // CSR
    void * rp = nullptr;  // uint32_t* or int64_t*
    int32_t * col = nullptr;
    double * val = nullptr;
    int32_t * tile_row = nullptr;
    int64_t ntiles = 0;
    int csr_tile = 0;  // entries per tile the tile table was built for
    int csr_grid = 0;  // ... and the persistent grid size
    int64_t csr_chunk = 0;  // non-zeros per CTA
    int csr_tpc = 0;        // tiles per CTA
    int32_t * span_row = nullptr;  // warp / flat kernels: row holding the first entry of every span_size-entry span
    int span_size = 0;
    int32_t * flat_meta = nullptr;  // flat kernel: {first row, -, -, -, 128-bit row-start mask} per 128-entry span
    int flat_span = 0;              // entries per span the metadata was built for (128 or 256)
    bool flat_has_empty = false;    // some row is empty: the metadata numbers the non-empty rows, flat_rowmap translates back
    bool flat_meta_compressed = false;
    int32_t * flat_rowmap = nullptr;  // non-empty row number -> row
    int32_t * slice_col = nullptr;  // sliced kernel: column_index / value with every 32-row slice stored slot-major
    double * slice_val = nullptr;
    // index runs of the slot-major copy: slice_col is then a COMPRESSED stream -- a slice whose entries lie on few diagonals
    // stores its slots by offset (column - row) with one descriptor per slot ({base, mask}, or the base alone when every slot
    // holds all 32 rows), other slices their explicit columns; per slice a flag word (form + slot count) and the slice's
    // offset in the stream (kernels_csr_sliced.cu)
    uint32_t * slice_flags = nullptr;
    void * slice_cofs = nullptr;    // uint32_t* or int64_t*, like rp; (rows/32 + 1) entries
    void * slice_meta = nullptr;    // {flags, cofs, row_ptr[32 s], -} per slice in one 16 / 32-byte record (what the SpMV kernel reads)
    bool slice_runs = false;
    int64_t slice_ccount = 0;       // int32 entries in slice_col
    bool slice_unavailable = false;  // the copy could not be allocated: automatic selection stays with the flat kernel
    int64_t csr_maxlen = -1;       // longest row (computed on first use)

    // ELL (column-major)
    int32_t * ell_col = nullptr;
    double * ell_val = nullptr;
    int64_t ell_pitch = 0;

    // COO (whole matrix, or the hybrid tail)
    int32_t * coo_row = nullptr;
    int32_t * coo_col = nullptr;
    double * coo_val = nullptr;
    int64_t coo_n = 0;
    bool coo_sorted = false;
    int coo_col_shift = 0;  // > 0: entries are partitioned by column block of 2^shift columns, row-sorted inside a block
    // "hot column" layout of the scattered-gather kernel (coo_hot_kernel): the entries are cut into segments (never
    // across a column block); every segment has a table of its most referenced columns, whose x values the CTA that
    // runs the segment keeps in shared memory, and coo_colh = column_index with those columns replaced by
    // 0x80000000 | slot.  coo_col itself stays as it is (export, conversion and the cache model read it).
    int32_t * coo_colh = nullptr;
    int32_t * coo_hot_cols = nullptr;  // [coo_nseg][coo_hot_h] global column of every slot (unused slots: column 0)
    int64_t * coo_seg = nullptr;       // [coo_nseg + 1] entry offsets
    int coo_hot_h = 0, coo_nseg = 0;
    bool coo_hot_tried = false;        // the builder ran (and may have decided against the layout)
    double coo_hot_coverage = 0.0;     // fraction of the entries whose column is in its segment's table
    cudaTextureObject_t x_tex = 0;     // (experiment "coo.xload" = 5 / 6) x as a linear texture, for the pointer below
    const double * x_tex_ptr = nullptr;

    // pipelined host-buffer path of the sliced CSR kernel: largest column referenced by each of host_chunks row chunks
    int host_chunks = 0;
    int64_t host_rows_per_chunk = 0;
    int host_colmax[64] = {};
    // fused compute + halo push (row-partitioned executor, sliced CSR kernel in store mode): rows [push_lo, push_hi) of the
    // next launch are ALSO stored at push_y[..][row] -- a peer GPU's x buffer, mapped over NVLink
    double * push_y[2] = {nullptr, nullptr};
    int64_t push_lo[2] = {0, 0}, push_hi[2] = {0, 0};
    // optional row range for the next launches (ELL, sliced CSR): [range_begin, range_end), range_end <= 0 = all rows
    int64_t range_begin = 0, range_end = 0;
    const double * host_y_in = nullptr;    // zero-copy host-buffer path (ELL): device-visible host pointers
    double * host_y_out = nullptr;
    cudaStream_t upload_stream = nullptr;  // second stream of the pipelined host-buffer path
    cudaEvent_t ev_x = nullptr;
    cudaEvent_t ev_chunk[64] = {};

    // vectors
    double * x = nullptr;
    double * y = nullptr;
    bool own_x = false, own_y = false;

    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t device_bytes = 0;
    int sm_count = 148;

    // options (spmvb200_set_option)
    int64_t opt_csr_tile = 0;     // 0 = auto
    int64_t opt_csr_stages = 0;   // 0 = auto
    int64_t opt_csr_threads = 0;  // threads per CTA (128, 256), 0 = auto
    int64_t opt_pdl = 1;          // programmatic dependent launch
    int64_t opt_independent = 0;  // caller promises: the previous kernel in the stream does not write this x
    int64_t opt_csr_lanes = 0;    // lanes per row in direct mode (1, 2, 4, 8), 0 = auto
    int64_t opt_csr_algo = 0;     // 0 auto, 1 direct (thread forms its row's products), 2 product pass
    int64_t opt_csr_ctas = 0;     // CTAs per SM of the persistent grid, 0 = auto
    int64_t opt_csr_batch = 0;    // sliced kernel: slots in flight per lane (2, 4, 8), 0 = auto
    int64_t opt_csr_rowptr_path = 0;  // flat kernel, matrices with empty rows: 1 = rebuild row numbers from row_ptr (MASK = false path)
    int64_t opt_csr_entries = 0;  // flat kernel: entries per lane (4, 8), 0 = auto
    int64_t opt_csr_rmw = 0;      // sliced kernel, y += A*x: 1 = the owning lane updates y with a plain read-modify-write instead of a
                                  // reduction (needs the launches ordered); off by default: measured slower
    int64_t opt_csr_probe = 0;    // 1 regular traffic (values only), 2 irregular traffic (x gather only): csr-matrix-spmv.cpp:35-61
    int64_t opt_csr_regs = 0;        // (experiment) sliced kernel with index runs: register budget 40 / 48 instead of 32
    int64_t opt_csr_index_runs = 0;  // sliced kernel: 0 store index runs when they shrink the column stream to <= 3/4, 1 always, -1 never
    int64_t opt_csr_drop = 1;     // sliced kernel: free the row-major column_index/value once the slot-major copy exists (0 = keep both)
    int64_t opt_csr_spare = 0;    // CTA slots per SM left free (for a concurrent NCCL kernel)
    int64_t opt_ell_rows = 0;     // rows per thread (1, 2, 4), 0 = auto
    int64_t opt_ell_block = 0;    // threads per block, 0 = auto
    int64_t opt_coo_threads = 0;  // threads per CTA of the segmented kernel (64, 128, 256), 0 = auto
    int64_t opt_coo_stages = 0;
    int64_t opt_coo_ctas = 0;
    int64_t opt_coo_algo = 0;     // 0 auto, 1 shared-memory staged (coo_segmented_kernel), 2 register-staged (coo_warp_kernel)
    int64_t opt_coo_items = 0;    // stripes of 32 entries per warp of coo_warp_kernel (2, 4, 8), 0 = auto
    int64_t opt_coo_xload = 0;    // cache path of the x gather (ptx.cuh ld_x): 0 nc, 1 cg (L2 only), 2 nc no_allocate, 3 nc evict_last, 4 cp.async, 5 texture, 6 texture + nc
    int64_t opt_coo_carveout = -1;     // >= 0: preferred shared-memory carve-out in percent (experiment: shrinks the L1)
    int64_t opt_coo_hot = 0;      // hot-column kernel: 0 / -1 off (default: it measured slower), 1 on, 2 on if the gathers are scattered
    int64_t opt_coo_hot_slots = 0;     // table size per segment (multiple of 1024, <= 27648), 0 = auto
    int64_t opt_coo_hot_threads = 0;   // threads per CTA (512, 1024), 0 = auto
    int64_t opt_coo_hot_entries = 0;   // entries per lane (4, 8), 0 = auto
    int64_t opt_coo_hot_segs = 0;      // segments per CTA, 0 = auto
    int64_t opt_host_zero_copy = 1;  // spmvb200_spmv_host: let the ELL kernel read/write pinned host y directly
    int64_t opt_host_chunks = 0;  // row chunks of the pipelined spmvb200_spmv_host (ELL), 0 = 8, 1 = off
    int64_t opt_beta0 = 0;        // 1: y = A*x (y is cleared first) instead of y += A*x
    // per-launch decisions of plan_run() (abi.cu): launch attribute and whether the kernel may skip
    // griddepcontrol.wait because nothing in flight on its stream writes its x or reads its y
    bool run_pdl = true, run_independent = false;
    bool run_rmw = false;         // this launch updates y with plain loads and stores by the lanes that own the rows (sliced CSR)
    bool run_beta0 = false;       // this launch computes y = alpha*A*x: kernels that own whole rows store, the others clear y first
    double alpha = 1.0;           // y += alpha*A*x (spmvb200_set_alpha); 1.0 is exact, the reference's semantics
    bool aux_dirty = false;       // an auxiliary table was just (re)built on the stream: serialise the next launch
    bool dry_run = false;         // spmvb200_prepare: build the launch metadata of the selected kernel, launch nothing
    const char * kernel_name = "";
};

namespace spmvb200 {

using Matrix = spmvb200_matrix_s;

// ---- allocation helpers (track device_bytes) ------------------------------------------------
template <typename T>
inline int dev_alloc(Matrix * m, T ** p, int64_t count)
{
    *p = nullptr;
    size_t bytes = sizeof(T) * (size_t)(count > 0 ? count : 1);
    cudaError_t e = cudaMalloc((void **)p, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
    if (m) m->device_bytes += (int64_t)bytes;
    return 0;
}

// Streamed arrays get kPadEntries zeroed elements of slack so whole-tile bulk copies and vector
// loads never leave the allocation.
template <typename T>
inline int alloc_streamed(Matrix * m, T ** p, int64_t count)
{
    const int64_t cap = round_up(count, 4096) + kPadEntries;
    int rc = dev_alloc(m, p, cap);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(*p + count, 0, sizeof(T) * (size_t)(cap - count), m->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync", __FILE__, __LINE__);
    return 0;
}

// Scratch allocations that are not part of the resident matrix.
template <typename T>
struct Scratch {
    T * p = nullptr;
    Scratch() = default;
    Scratch(const Scratch &) = delete;
    Scratch & operator=(const Scratch &) = delete;
    ~Scratch() { if (p) cudaFree(p); }
    int alloc(int64_t n)
    {
        cudaError_t e = cudaMalloc((void **)&p, sizeof(T) * (size_t)(n > 0 ? n : 1));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)", __FILE__, __LINE__);
        return 0;
    }
    T * release() { T * q = p; p = nullptr; return q; }
};

inline unsigned grid_for(int64_t n, int sm = 148)
{
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)sm * 16) g = (int64_t)sm * 16;
    return (unsigned)(g < 1 ? 1 : g);
}

int fill_device(double * p, int64_t n, double v, cudaStream_t s);

// ---- in-flight tracking of library-owned streams (abi.cu) -----------------------------------
void stream_register(cudaStream_t s);    // a stream created by the library
void stream_forget(cudaStream_t s);
void stream_invalidate(cudaStream_t s);  // something other than an SpMV launch was enqueued
void stream_synced(cudaStream_t s);      // the stream was synchronised: nothing is in flight
void plan_run(Matrix * m, bool conservative);

// What the launchers pass on: {launch with the PDL attribute, kernel may skip griddepcontrol.wait}.
// A launch that follows the (re)build of an auxiliary table on the same stream is fully serialised.
struct RunMode {
    bool pdl;
    int independent;
};
// Kernels that add partial row sums with reductions need y cleared first when the launch is y = alpha*A*x.
inline int clear_y_for_beta0(Matrix * m)
{
    if (!m->run_beta0) return 0;
    m->run_beta0 = false;
    cudaError_t e = cudaMemsetAsync(m->y, 0, sizeof(double) * (size_t)m->rows, m->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync", __FILE__, __LINE__);
    return 0;
}
inline int need_unit_alpha(const Matrix * m, const char * kernel)
{
    if (m->alpha == 1.0) return 0;
    return fail(SPMVB200_ERR_UNSUPPORTED, std::string(kernel) + " does not scale its result: use the default kernels with spmvb200_set_alpha");
}
inline RunMode run_mode(Matrix * m)
{
    RunMode r{m->run_pdl && !m->aux_dirty, (m->run_independent && !m->aux_dirty) ? 1 : 0};
    m->aux_dirty = false;
    return r;
}

// ---- launchers (kernels_csr*.cu, kernels_ell.cu, kernels_coo.cu) ------------------------------
int launch_csr(Matrix * m);
int launch_ell(Matrix * m, bool accumulate_into_y);
int launch_coo(Matrix * m);
int csr_build_tiles(Matrix * m);  // (re)build tile_row for the configured tile size
bool csr_uses_sliced_kernel(Matrix * m);  // would launch_csr run the sliced kernel (the one whose threads own whole rows)?
// The row-major column_index / value of a CSR matrix may have been dropped in favour of the slot-major copy
// ("csr.drop_row_major"); everything that reads them calls this first (rebuilds them from the copy).
int csr_ensure_row_major(Matrix * m);
int csr_drop_sliced(Matrix * m);  // rebuild the row-major arrays if needed and free the slot-major copy (an option that shapes it changed)

// ---- builders.cu -----------------------------------------------------------------------------
int matrix_new(Matrix ** out);
int matrix_alloc_vectors(Matrix * m);
void matrix_free(Matrix * m);
// Take ownership of device CSR arrays (unpadded, `stored` entries) and finish the layout.
int csr_adopt(Matrix * m, int64_t rows, int64_t cols, int64_t nnz, int64_t stored, bool off64,
              void * rp, int32_t * col, double * val);
int coo_adopt(Matrix * m, int64_t rows, int64_t cols, int64_t n, int32_t * row, int32_t * col,
              double * val, int mode, bool already_sorted);
int ell_from_csr(const Matrix * src, int skip_padding, bool check_int32, Matrix * dst);
int hyb_from_csr(const Matrix * src, int skip_padding, bool check_int32, Matrix * dst);
int coo_from_csr(const Matrix * src, int mode, Matrix * dst);
int coo_column_blocks(Matrix * m);
int coo_build_hot(Matrix * m);  // builds (or decides against) the hot-column layout; idempotent
int coo_row_major_copy(Matrix * m, int32_t * row2, int32_t * col2, double * val2);
int csr_from_entries_host(int64_t rows, int64_t cols, int64_t n, const int32_t * i, const int32_t * j,
                          const double * a, int32_t row_alignment, Matrix * dst);

}  // namespace spmvb200
