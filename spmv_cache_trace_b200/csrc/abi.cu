// abi.cu -- the extern "C" surface of libspmvb200.so (see include/spmv_b200.h).
//
// Thin by design: argument checks, handle bookkeeping, host<->device copies, and dispatch to the
// builders (builders.cu, generators.cu) and kernel launchers (kernels_*.cu).  No arithmetic of the
// SpMV path is done on the host anywhere in this library.
#include "common.cuh"
#include "mm_host.hpp"

#include <cub/cub.cuh>

#include <atomic>
#include <climits>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <new>
#include <stdexcept>
#include <vector>

// No exception crosses the C ABI: host-side allocations (std::vector, std::string, new) may throw.
#define SPMV_ABI_CATCH                                                                                  \
    catch (const std::bad_alloc &) { return ::spmvb200::fail(SPMVB200_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception & e) { return ::spmvb200::fail(SPMVB200_ERR_INVALID, e.what()); }

namespace spmvb200 {

static thread_local std::string g_error;
static std::atomic<int64_t> g_launches{0};
std::atomic<int> g_force_off64{0};
std::atomic<int64_t> g_coo_col_block_log2{0};

void set_error(const std::string & msg) { g_error = msg; }
int fail(int code, const std::string & msg)
{
    g_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char * what, const char * file, int line)
{
    const char * base = strrchr(file, '/');
    g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + (base ? base + 1 : file) +
              ":" + std::to_string(line) + ")";
    return SPMVB200_ERR_CUDA;
}
void count_launch(int n) { g_launches += n; }

// declared in builders.cu / generators.cu
int store_offsets(Matrix * m, const int64_t * d_rp64, int64_t rows, int64_t stored);
int gen_stencil(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin, int64_t row_end, Matrix * csr);
int gen_rmat(int scale, int edge_factor, uint64_t seed, double a, double b, double c, int64_t row_begin,
             int64_t row_end, Matrix * csr);
int launch_csr_sliced(Matrix * m);  // kernels_csr_sliced.cu
int csr_chunk_colmax(Matrix * m, int64_t rows_per_chunk, int chunks, int * host_out);

__global__ void transpose_to_colmajor_kernel(int64_t rows, int64_t W, int64_t pitch, const int32_t * col_rm,
                                             const double * val_rm, int32_t * ecol, double * eval)
{
    const int64_t total = rows * W;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = k / W, l = k - r * W;
        ecol[l * pitch + r] = col_rm[k];
        eval[l * pitch + r] = val_rm[k];
    }
}

__global__ void transpose_to_rowmajor_kernel(int64_t rows, int64_t W, int64_t pitch, const int32_t * ecol,
                                             const double * eval, int32_t * col_rm, double * val_rm)
{
    const int64_t total = rows * W;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = k / W, l = k - r * W;
        col_rm[k] = ecol[l * pitch + r];
        val_rm[k] = eval[l * pitch + r];
    }
}

__global__ void zero_based_kernel(int64_t n, const int32_t * in, int32_t * out, int32_t limit, int * bad)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = in[k] - 1;
        if (v < 0 || v >= limit) *bad = 1;
        out[k] = v;
    }
}

template <typename OffT>
__global__ void partition_nnz_kernel(int64_t rows, const OffT * rp, int parts, int64_t * starts)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > parts) return;
    if (p == 0) { starts[0] = 0; return; }
    if (p == parts) { starts[p] = rows; return; }
    const unsigned long long nnz = (unsigned long long)rp[rows];
    const unsigned long long target = (unsigned long long)(((unsigned __int128)nnz * (unsigned)p) / (unsigned)parts);
    int64_t lo = 0, hi = rows;  // first r in [0, rows] with rp[r] >= target
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if ((unsigned long long)rp[mid] >= target) hi = mid; else lo = mid + 1;
    }
    starts[p] = lo;
}

// Weighted variant: a row costs its entries plus row_weight_q10 / 1024 "entries" of per-row work (y traffic, padding, a
// reduction per run): start_p = first r with 1024*rp[r] + w*r >= floor(p * (1024*nnz + w*rows) / P).
template <typename OffT>
__global__ void partition_weighted_kernel(int64_t rows, const OffT * rp, int parts, int64_t w, int64_t * starts)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > parts) return;
    if (p == 0) { starts[0] = 0; return; }
    if (p == parts) { starts[p] = rows; return; }
    const unsigned __int128 total = (unsigned __int128)1024 * (unsigned long long)rp[rows] + (unsigned __int128)w * (unsigned long long)rows;
    const unsigned __int128 target = total * (unsigned)p / (unsigned)parts;
    int64_t lo = 0, hi = rows;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        const unsigned __int128 c = (unsigned __int128)1024 * (unsigned long long)rp[mid] + (unsigned __int128)w * (unsigned long long)mid;
        if (c >= target) hi = mid; else lo = mid + 1;
    }
    starts[p] = lo;
}

template <typename OffT>
__global__ void rebase_offsets_kernel(int64_t n, const OffT * rp, int64_t first, int64_t * out)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        out[r] = (int64_t)rp[r] - first;
}

// Column split of a CSR matrix: entries with column in [cb, ce) go to the "inside" matrix, the rest to "outside".
template <typename OffT>
__global__ void column_count_kernel(int64_t rows, const OffT * rp, const int32_t * col, int64_t cb, int64_t ce,
                                    int64_t * n_in, int64_t * n_out)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t in = 0, len = 0;
        if (r < rows) {
            const int64_t lo = (int64_t)rp[r], hi = (int64_t)rp[r + 1];
            len = hi - lo;
            for (int64_t k = lo; k < hi; ++k) in += (col[k] >= cb && col[k] < ce) ? 1 : 0;
        }
        n_in[r] = in;
        n_out[r] = len - in;
    }
}

template <typename OffT>
__global__ void column_fill_kernel(int64_t rows, const OffT * rp, const int32_t * col, const double * val, int64_t cb,
                                   int64_t ce, const int64_t * rp_in, const int64_t * rp_out, int32_t * col_in,
                                   double * val_in, int32_t * col_out, double * val_out)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t pi = rp_in[r], po = rp_out[r];
        for (int64_t k = (int64_t)rp[r]; k < (int64_t)rp[r + 1]; ++k) {
            const int32_t c = col[k];
            if (c >= cb && c < ce) { col_in[pi] = c; val_in[pi++] = val[k]; }
            else { col_out[po] = c; val_out[po++] = val[k]; }
        }
    }
}

// out[0] = min column, out[1] = max column, out[2] = lo_end, out[3] = hi_begin (see spmv_b200.h)
template <typename OffT>
__global__ void column_span_kernel(int64_t rows, const OffT * rp, const int32_t * col, int64_t cb, int64_t ce,
                                   long long * out)
{
    long long cmin = LLONG_MAX, cmax = -1, lo_end = 0, hi_begin = rows;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        long long rmin = LLONG_MAX, rmax = -1;
        for (int64_t k = (int64_t)rp[r]; k < (int64_t)rp[r + 1]; ++k) {
            const long long c = col[k];
            rmin = c < rmin ? c : rmin;
            rmax = c > rmax ? c : rmax;
        }
        if (rmax < 0) continue;
        cmin = rmin < cmin ? rmin : cmin;
        cmax = rmax > cmax ? rmax : cmax;
        if (rmin < cb && r + 1 > lo_end) lo_end = r + 1;
        if (rmax >= ce && r < hi_begin) hi_begin = r;
    }
    for (int o = 16; o > 0; o >>= 1) {
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        lo_end = max(lo_end, __shfl_xor_sync(0xffffffffu, lo_end, o));
        hi_begin = min(hi_begin, __shfl_xor_sync(0xffffffffu, hi_begin, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], cmin);
        atomicMax(&out[1], cmax);
        atomicMax(&out[2], lo_end);
        atomicMin(&out[3], hi_begin);
    }
}

static int upload_ell_rowmajor(Matrix * m, int64_t rows, int64_t W, const int32_t * col, const double * val, int skip)
{
    cudaStream_t s = m->stream;
    m->ell_w = W;
    m->ell_pitch = round_up(std::max<int64_t>(rows, 1), 32);
    m->skip_padding = skip;
    const int64_t slots = m->ell_pitch * W, n = rows * W;
    SPMV_TRY(alloc_streamed(m, &m->ell_col, slots));
    SPMV_TRY(alloc_streamed(m, &m->ell_val, slots));
    if (n > 0) {
        Scratch<int32_t> dc;
        Scratch<double> dv;
        SPMV_TRY(dc.alloc(n)); SPMV_TRY(dv.alloc(n));
        SPMV_CUDA(cudaMemcpyAsync(dc.p, col, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(dv.p, val, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemsetAsync(m->ell_col, 0, sizeof(int32_t) * (size_t)slots, s));
        SPMV_CUDA(cudaMemsetAsync(m->ell_val, 0, sizeof(double) * (size_t)slots, s));
        transpose_to_colmajor_kernel<<<grid_for(n), 256, 0, s>>>(rows, W, m->ell_pitch, dc.p, dv.p, m->ell_col, m->ell_val);
        SPMV_CUDA(cudaGetLastError());
        SPMV_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

static int upload_coo(Matrix * m, int64_t n, const int32_t * row, const int32_t * col, const double * val)
{
    cudaStream_t s = m->stream;
    SPMV_TRY(alloc_streamed(m, &m->coo_row, n));
    SPMV_TRY(alloc_streamed(m, &m->coo_col, n));
    SPMV_TRY(alloc_streamed(m, &m->coo_val, n));
    if (n > 0) {
        SPMV_CUDA(cudaMemcpyAsync(m->coo_row, row, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(m->coo_col, col, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(m->coo_val, val, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    m->coo_n = n;
    return 0;
}

struct Guard {  // frees a half-built matrix on early return
    Matrix * m;
    explicit Guard(Matrix * m) : m(m) {}
    ~Guard() { if (m) matrix_free(m); }
    Matrix * release() { Matrix * q = m; m = nullptr; return q; }
};

static int finish(Guard & g, spmvb200_matrix_t * out)
{
    SPMV_TRY(matrix_alloc_vectors(g.m));
    SPMV_CUDA(cudaStreamSynchronize(g.m->stream));
    *out = g.release();
    return 0;
}

// ---- what is in flight on the streams the library owns ---------------------------------------------------------
// The SpMV kernels read the immutable matrix, gather x and add to y with reductions (RED), which commute.
// Two consecutive launches therefore only need to be ordered when one writes what the other reads.
// For streams created by the library (nothing else can enqueue work on them) the library knows every
// kernel in flight: the record below keeps the bounding address ranges written (y) and read (x) by the
// SpMV kernels launched since the last fully ordered point.  A launch whose x is not written and whose
// y is not read by anything in flight skips griddepcontrol.wait, so its CTAs start while the previous
// kernel drains (the reference protocol -- x constant, y accumulated, profile-kernel.cpp:159-161 --
// is exactly this case).  Any other API call on the matrix invalidates the record; a launch that must
// be ordered behind kernels that did not wait is issued without the PDL attribute, i.e. with ordinary
// stream serialisation behind all of them.  User-provided streams (spmvb200_set_stream) are never
// tracked: foreign work may precede the launch, so it always waits.
struct StreamRec {
    bool valid = false;  // everything enqueued since the last ordered point is an SpMV launch recorded here
    uintptr_t wlo = 0, whi = 0, rlo = 0, rhi = 0;
    uintptr_t slo = 0, shi = 0;  // range written with plain stores (a "beta0" launch): reductions must not overtake it
    int chain = 0;       // launches since the last ordered point that skipped griddepcontrol.wait
};
static std::mutex g_stream_mu;
static std::unordered_map<cudaStream_t, StreamRec> g_streams;

void stream_register(cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_stream_mu);
    g_streams[s] = StreamRec{};
}
void stream_forget(cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_stream_mu);
    g_streams.erase(s);
}
void stream_invalidate(cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_streams.find(s);
    if (it != g_streams.end()) it->second.valid = false;
}
void stream_synced(cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_streams.find(s);
    if (it != g_streams.end()) {
        it->second = StreamRec{};
        it->second.valid = true;
    }
}

static inline bool overlaps(uintptr_t a0, uintptr_t a1, uintptr_t b0, uintptr_t b1) { return a0 < b1 && b0 < a1; }

// Decide how the next kernel of `m` is launched (run_pdl, run_independent) and record it.
void plan_run(Matrix * m, bool conservative)
{
    m->run_pdl = m->opt_pdl != 0;
    m->run_independent = false;
    const uintptr_t x0 = (uintptr_t)m->x, x1 = x0 + 8u * (uintptr_t)m->cols;
    const uintptr_t y0 = (uintptr_t)m->y, y1 = y0 + 8u * (uintptr_t)m->rows;
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_streams.find(m->stream);
    if (it == g_streams.end()) {
        // Not a library stream.  The caller may have tied it to other streams with events (the multi-GPU mode does:
        // exchange <-> compute), and a kernel launched with the PDL attribute next to such an edge was observed to
        // break the ordering the events are there for (tools/dist_check.py, hybrid pieces, steps issued back to back:
        // wrong rows with PDL, none without).  So on foreign streams launches are plain unless "pdl" = 2 insists.
        m->run_pdl = m->opt_pdl == 2;
        m->run_independent = m->run_pdl && m->opt_independent > 0 && !conservative;
        return;
    }
    if (conservative) m->run_pdl = false;  // the host-buffer paths order their pieces with events too
    StreamRec & r = it->second;
    const bool proven = r.valid && !overlaps(x0, x1, r.wlo, r.whi) && !overlaps(y0, y1, r.rlo, r.rhi) &&
                        !overlaps(y0, y1, r.slo, r.shi);
    // The kernels gather x through the read-only path (ld.global.nc), which PTX defines only for data that nothing
    // writes during the kernel's lifetime -- and under PDL that lifetime begins while the predecessor is still
    // running.  A launch whose x may have been written by a kernel in flight (an iteration x_{k+1} = A x_k with
    // the buffers swapped by bind_x / bind_y; or anything the record cannot vouch for) is therefore issued without
    // the PDL attribute: ordinary stream serialisation, the whole predecessor retired before the first CTA starts.
    if ((!r.valid || overlaps(x0, x1, r.wlo, r.whi)) && m->opt_pdl != 2) m->run_pdl = false;  // ("pdl" = 2: experiments)
    if (!conservative && m->run_pdl && !m->opt_beta0 && !m->run_beta0 && !m->run_rmw &&
        (m->opt_independent > 0 || (m->opt_independent == 0 && proven))) {
        m->run_independent = true;  // validity of the record is unchanged; its ranges grow
        r.wlo = r.whi > r.wlo ? std::min(r.wlo, y0) : y0;
        r.whi = std::max(r.whi, y1);
        r.rlo = r.rhi > r.rlo ? std::min(r.rlo, x0) : x0;
        r.rhi = std::max(r.rhi, x1);
        r.chain++;
        return;
    }
    // ordered launch: behind un-waited kernels only ordinary stream serialisation is transitive
    if (r.chain > 0) m->run_pdl = false;
    r = StreamRec{};
    r.valid = !conservative;
    r.wlo = y0; r.whi = y1; r.rlo = x0; r.rhi = x1;
    if (m->run_beta0 || m->opt_beta0 || m->run_rmw) { r.slo = y0; r.shi = y1; }  // plain stores: reductions must not overtake them
}

// Entry check of every API call that is not an SpMV launch: whatever it enqueues is unknown to the record.
static int check(spmvb200_matrix_t m)
{
    if (!m) return fail(SPMVB200_ERR_INVALID, "null matrix handle");
    cudaError_t e = cudaSetDevice(m->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    stream_invalidate(m->stream);
    return 0;
}
// Entry check of the launch paths.
static int check_run(spmvb200_matrix_t m)
{
    if (!m) return fail(SPMVB200_ERR_INVALID, "null matrix handle");
    cudaError_t e = cudaSetDevice(m->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    return 0;
}

}  // namespace spmvb200

using namespace spmvb200;

extern "C" {

const char * spmvb200_last_error(void) { return g_error.c_str(); }
int spmvb200_version(void) { return SPMVB200_VERSION; }
int64_t spmvb200_launch_count(void) { return g_launches.load(); }

int spmvb200_set_global_option(const char * key, int64_t value)
try {
    if (!key) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (!strcmp(key, "force_offsets64")) { g_force_off64 = value ? 1 : 0; return 0; }
    if (!strcmp(key, "coo.col_block_log2")) { g_coo_col_block_log2 = value; return 0; }
    if (!strcmp(key, "mm.gp_partitioner")) { set_gp_partitioner(value ? 1 : 0); return 0; }
    return fail(SPMVB200_ERR_INVALID, std::string("unknown global option ") + key);
}
SPMV_ABI_CATCH

int spmvb200_device_count(int * count)
try {
    if (!count) return fail(SPMVB200_ERR_INVALID, "null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_set_device(int device)
try {
    SPMV_CUDA(cudaSetDevice(device));
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_device_props(int device, char * name, size_t name_cap, int * sm_count, int64_t * l2_bytes,
                          int64_t * mem_bytes, int * cc_major, int * cc_minor)
try {
    cudaDeviceProp p;
    SPMV_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && name_cap) { strncpy(name, p.name, name_cap - 1); name[name_cap - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    if (mem_bytes) *mem_bytes = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}
SPMV_ABI_CATCH

// ---- Matrix Market (host) ---------------------------------------------------------------------

int spmvb200_mm_parse(const char * text, size_t len, spmvb200_mm_t * out)
try {
    if (!text || !out) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_parse_text(text, len, out);
}
SPMV_ABI_CATCH
int spmvb200_mm_load(const char * path, spmvb200_mm_t * out)
try {
    if (!path || !out) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_load_path(path, out);
}
SPMV_ABI_CATCH
int spmvb200_mm_from_entries(int32_t rows, int32_t columns, int32_t n, const int32_t * i, const int32_t * j,
                             const double * a, spmvb200_mm_t * out)
try {
    if (!out || rows < 0 || columns < 0 || n < 0 || (n > 0 && (!i || !j || !a)))
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    return mm_from_entries(rows, columns, n, i, j, a, out);
}
SPMV_ABI_CATCH
int spmvb200_mm_info(spmvb200_mm_t mm, int32_t * rows, int32_t * columns, int32_t * n, int32_t * field,
                     int32_t * symmetry, int32_t * format)
try {
    if (!mm) return fail(SPMVB200_ERR_INVALID, "null mm handle");
    if (rows) *rows = mm->rows;
    if (columns) *columns = mm->columns;
    if (n) *n = mm->num_entries;
    if (field) *field = mm->field;
    if (symmetry) *symmetry = mm->symmetry;
    if (format) *format = mm->format;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_mm_entries(spmvb200_mm_t mm, const int32_t ** i, const int32_t ** j, const double ** a)
try {
    if (!mm) return fail(SPMVB200_ERR_INVALID, "null mm handle");
    if (i) *i = mm->i.data();
    if (j) *j = mm->j.data();
    if (a) *a = mm->a.data();
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_mm_row_lengths(spmvb200_mm_t mm, int32_t * lengths)
try {
    if (!mm || !lengths) return fail(SPMVB200_ERR_INVALID, "null argument");
    return mm_row_lengths(mm, lengths);
}
SPMV_ABI_CATCH
int spmvb200_mm_max_row_length(spmvb200_mm_t mm, int32_t * out)
try {
    if (!mm || !out) return fail(SPMVB200_ERR_INVALID, "null argument");
    std::vector<int32_t> len((size_t)std::max(mm->rows, 1));
    SPMV_TRY(mm_row_lengths(mm, len.data()));
    int32_t best = 0;
    for (int32_t r = 0; r < mm->rows; r++) best = std::max(best, len[r]);
    *out = best;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_mm_sort_row_major(spmvb200_mm_t mm)
try {
    if (!mm) return fail(SPMVB200_ERR_INVALID, "null mm handle");
    return mm_sort(mm, true);
}
SPMV_ABI_CATCH
int spmvb200_mm_sort_column_major(spmvb200_mm_t mm)
try {
    if (!mm) return fail(SPMVB200_ERR_INVALID, "null mm handle");
    return mm_sort(mm, false);
}
SPMV_ABI_CATCH
void spmvb200_mm_free(spmvb200_mm_t mm) { delete mm; }

// ---- builders from Matrix Market ----------------------------------------------------------------

static int need_coordinate(spmvb200_mm_t mm, void * out)
{
    if (!mm || !out) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (mm->format != 0)  // csr-matrix.cpp:197-198 and siblings
        return fail(SPMVB200_ERR_PARSE, "Expected matrix in coordinate format");
    return 0;
}

static int csr_from_mm(spmvb200_mm_t mm, int32_t row_alignment, Matrix ** out)
{
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    SPMV_TRY(csr_from_entries_host(mm->rows, mm->columns, mm->num_entries, mm->i.data(), mm->j.data(), mm->a.data(),
                                   row_alignment, m));
    *out = g.release();
    return 0;
}

int spmvb200_csr_from_mm(spmvb200_mm_t mm, int32_t row_alignment, spmvb200_matrix_t * out)
try {
    SPMV_TRY(need_coordinate(mm, out));
    Matrix * m = nullptr;
    SPMV_TRY(csr_from_mm(mm, row_alignment, &m));
    Guard g(m);
    return finish(g, out);
}
SPMV_ABI_CATCH

int spmvb200_coo_from_mm(spmvb200_mm_t mm, int32_t coo_mode, spmvb200_matrix_t * out)
try {
    SPMV_TRY(need_coordinate(mm, out));
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    const int64_t n = mm->num_entries;
    cudaStream_t s = m->stream;
    // file order kept, 1-based -> 0-based (coo-matrix.cpp:226-239), done on the device
    int32_t *row = nullptr, *col = nullptr;
    double * val = nullptr;
    SPMV_TRY(alloc_streamed(m, &row, n));
    SPMV_TRY(alloc_streamed(m, &col, n));
    SPMV_TRY(alloc_streamed(m, &val, n));
    m->coo_row = row; m->coo_col = col; m->coo_val = val;  // owned from here on
    if (n > 0) {
        Scratch<int32_t> t;
        Scratch<int> bad;
        SPMV_TRY(t.alloc(n)); SPMV_TRY(bad.alloc(1));
        SPMV_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
        SPMV_CUDA(cudaMemcpyAsync(t.p, mm->i.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        zero_based_kernel<<<grid_for(n), 256, 0, s>>>(n, t.p, row, mm->rows, bad.p);
        SPMV_CUDA(cudaMemcpyAsync(t.p, mm->j.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        zero_based_kernel<<<grid_for(n), 256, 0, s>>>(n, t.p, col, mm->columns, bad.p);
        SPMV_CUDA(cudaGetLastError());
        SPMV_CUDA(cudaMemcpyAsync(val, mm->a.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
        int hbad = 0;
        SPMV_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        if (hbad) return fail(SPMVB200_ERR_INVALID, "entry index outside the matrix");
    }
    m->nnz = n;
    SPMV_TRY(coo_adopt(m, mm->rows, mm->columns, n, row, col, val, coo_mode, false));
    return finish(g, out);
}
SPMV_ABI_CATCH

int spmvb200_ell_from_mm(spmvb200_mm_t mm, int32_t skip_padding, spmvb200_matrix_t * out)
try {
    SPMV_TRY(need_coordinate(mm, out));
    Matrix * csr = nullptr;
    SPMV_TRY(csr_from_mm(mm, 1, &csr));
    Guard gc(csr);
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    SPMV_TRY(ell_from_csr(csr, skip_padding, true, m));
    return finish(g, out);
}
SPMV_ABI_CATCH

int spmvb200_hyb_from_mm(spmvb200_mm_t mm, int32_t skip_padding, spmvb200_matrix_t * out)
try {
    SPMV_TRY(need_coordinate(mm, out));
    Matrix * csr = nullptr;
    SPMV_TRY(csr_from_mm(mm, 1, &csr));
    Guard gc(csr);
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    SPMV_TRY(hyb_from_csr(csr, skip_padding, true, m));
    return finish(g, out);
}
SPMV_ABI_CATCH

// ---- from converted host arrays --------------------------------------------------------------------

int spmvb200_csr_create64(int64_t rows, int64_t columns, int64_t num_entries, const int64_t * row_ptr,
                          const int32_t * column_index, const double * value, spmvb200_matrix_t * out)
try {
    if (!out || rows < 0 || columns < 0 || !row_ptr) return fail(SPMVB200_ERR_INVALID, "bad argument");
    if (rows >= INT32_MAX || columns >= INT32_MAX) return fail(SPMVB200_ERR_UNSUPPORTED, "rows/columns must fit int32");
    const int64_t stored = row_ptr[rows];
    if (stored < 0 || (stored > 0 && (!column_index || !value))) return fail(SPMVB200_ERR_INVALID, "bad argument");
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    cudaStream_t s = m->stream;
    m->format = SPMVB200_CSR;
    m->rows = rows; m->cols = columns; m->nnz = num_entries; m->stored = stored;
    SPMV_TRY(alloc_streamed(m, &m->col, stored));
    SPMV_TRY(alloc_streamed(m, &m->val, stored));
    if (stored > 0) {
        SPMV_CUDA(cudaMemcpyAsync(m->col, column_index, sizeof(int32_t) * (size_t)stored, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(m->val, value, sizeof(double) * (size_t)stored, cudaMemcpyHostToDevice, s));
    }
    Scratch<int64_t> rp64;
    SPMV_TRY(rp64.alloc(rows + 1));
    SPMV_CUDA(cudaMemcpyAsync(rp64.p, row_ptr, sizeof(int64_t) * (size_t)(rows + 1), cudaMemcpyHostToDevice, s));
    SPMV_TRY(store_offsets(m, rp64.p, rows, stored));
    SPMV_TRY(csr_build_tiles(m));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return finish(g, out);
}
SPMV_ABI_CATCH

int spmvb200_csr_create(int32_t rows, int32_t columns, int32_t num_entries, const int32_t * row_ptr,
                        const int32_t * column_index, const double * value, spmvb200_matrix_t * out)
try {
    if (!out || rows < 0 || !row_ptr) return fail(SPMVB200_ERR_INVALID, "bad argument");
    std::vector<int64_t> rp((size_t)rows + 1);
    for (int32_t r = 0; r <= rows; r++) rp[r] = row_ptr[r];
    return spmvb200_csr_create64(rows, columns, num_entries, rp.data(), column_index, value, out);
}
SPMV_ABI_CATCH

int spmvb200_coo_create(int32_t rows, int32_t columns, int64_t n, const int32_t * row_index,
                        const int32_t * column_index, const double * value, int32_t coo_mode,
                        spmvb200_matrix_t * out)
try {
    if (!out || rows < 0 || columns < 0 || n < 0 || (n > 0 && (!row_index || !column_index || !value)))
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    SPMV_TRY(upload_coo(m, n, row_index, column_index, value));
    m->nnz = n;
    SPMV_TRY(coo_adopt(m, rows, columns, n, m->coo_row, m->coo_col, m->coo_val, coo_mode, false));
    return finish(g, out);
}
SPMV_ABI_CATCH

int spmvb200_ell_create(int32_t rows, int32_t columns, int32_t num_entries, int32_t row_length,
                        const int32_t * column_index, const double * value, int32_t skip_padding,
                        spmvb200_matrix_t * out)
try {
    if (!out || rows < 0 || columns < 0 || row_length < 0) return fail(SPMVB200_ERR_INVALID, "bad argument");
    if ((int64_t)rows * row_length > 0 && (!column_index || !value)) return fail(SPMVB200_ERR_INVALID, "bad argument");
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    m->format = SPMVB200_ELL;
    m->rows = rows; m->cols = columns; m->nnz = num_entries; m->stored = (int64_t)rows * row_length;
    SPMV_TRY(upload_ell_rowmajor(m, rows, row_length, column_index, value, skip_padding));
    return finish(g, out);
}
SPMV_ABI_CATCH

int spmvb200_hyb_create(int32_t rows, int32_t columns, int32_t num_entries, int32_t ell_row_length,
                        const int32_t * ell_column_index, const double * ell_value, int32_t ell_skip_padding,
                        int32_t num_coo_entries, const int32_t * coo_row_index, const int32_t * coo_column_index,
                        const double * coo_value, spmvb200_matrix_t * out)
try {
    if (!out || rows < 0 || columns < 0 || ell_row_length < 0 || num_coo_entries < 0)
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    m->format = SPMVB200_HYB;
    m->rows = rows; m->cols = columns; m->nnz = num_entries;
    m->n_ell = (int64_t)rows * ell_row_length; m->n_coo = num_coo_entries;
    m->stored = m->n_ell + m->n_coo;
    SPMV_TRY(upload_ell_rowmajor(m, rows, ell_row_length, ell_column_index, ell_value, ell_skip_padding));
    SPMV_TRY(upload_coo(m, num_coo_entries, coo_row_index, coo_column_index, coo_value));
    // the reference builds the tail row-major sorted (hybrid-matrix.cpp:402-408); verify, sort if not
    const int keep_format = m->format;
    const int64_t keep_nnz = m->nnz, keep_stored = m->stored;
    SPMV_TRY(coo_adopt(m, rows, columns, num_coo_entries, m->coo_row, m->coo_col, m->coo_val,
                       SPMVB200_COO_SEGMENTED, false));
    m->format = keep_format; m->nnz = keep_nnz; m->stored = keep_stored;
    return finish(g, out);
}
SPMV_ABI_CATCH

// ---- generators / conversion -------------------------------------------------------------------------

static int convert_from_csr(Matrix * csr, int32_t format, int32_t arg, bool check_int32, spmvb200_matrix_t * out)
{
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    if (format == SPMVB200_ELL) SPMV_TRY(ell_from_csr(csr, arg, check_int32, m));
    else if (format == SPMVB200_HYB) SPMV_TRY(hyb_from_csr(csr, arg, check_int32, m));
    else if (format == SPMVB200_COO) SPMV_TRY(coo_from_csr(csr, arg, m));
    else return fail(SPMVB200_ERR_INVALID, "unknown target format");
    m->row_offset = csr->row_offset;
    return finish(g, out);
}

int spmvb200_convert(spmvb200_matrix_t src, int32_t format, int32_t arg, spmvb200_matrix_t * out)
try {
    SPMV_TRY(check(src));
    if (!out) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (src->format != SPMVB200_CSR || src->row_alignment != 1)
        return fail(SPMVB200_ERR_UNSUPPORTED, "conversion source must be an unpadded CSR matrix");
    SPMV_TRY(csr_ensure_row_major(src));
    SPMV_CUDA(cudaStreamSynchronize(src->stream));
    return convert_from_csr(src, format, arg, false, out);
}
SPMV_ABI_CATCH

int spmvb200_gen_stencil(int32_t kind, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin, int64_t row_end,
                         int32_t format, spmvb200_matrix_t * out)
try {
    if (!out || nx < 1 || ny < 1 || nz < 1) return fail(SPMVB200_ERR_INVALID, "bad argument");
    const int64_t n = nx * ny * nz;
    if (row_begin == 0 && row_end == 0) row_end = n;
    if (row_begin < 0 || row_end > n || row_begin > row_end) return fail(SPMVB200_ERR_INVALID, "bad row range");
    if (n >= INT32_MAX) return fail(SPMVB200_ERR_UNSUPPORTED, "grid too large for int32 column indices");
    Matrix * csr = nullptr;
    SPMV_TRY(matrix_new(&csr));
    Guard g(csr);
    SPMV_TRY(gen_stencil(kind, nx, ny, nz, row_begin, row_end, csr));
    if (format == SPMVB200_CSR) return finish(g, out);
    return convert_from_csr(csr, format, 0, false, out);
}
SPMV_ABI_CATCH

int spmvb200_gen_rmat(int32_t scale, int32_t edge_factor, uint64_t seed, double a, double b, double c,
                      int64_t row_begin, int64_t row_end, int32_t format, int32_t coo_mode, spmvb200_matrix_t * out)
try {
    if (!out || scale < 1 || scale > 30 || edge_factor < 1) return fail(SPMVB200_ERR_INVALID, "bad argument");
    const int64_t n = (int64_t)1 << scale;
    if (row_begin == 0 && row_end == 0) row_end = n;
    if (row_begin < 0 || row_end > n || row_begin > row_end) return fail(SPMVB200_ERR_INVALID, "bad row range");
    Matrix * csr = nullptr;
    SPMV_TRY(matrix_new(&csr));
    Guard g(csr);
    SPMV_TRY(gen_rmat(scale, edge_factor, seed, a, b, c, row_begin, row_end, csr));
    if (format == SPMVB200_CSR) return finish(g, out);
    return convert_from_csr(csr, format, format == SPMVB200_COO ? coo_mode : 0, false, out);
}
SPMV_ABI_CATCH

// ---- inspection / export ----------------------------------------------------------------------------------

int spmvb200_matrix_info(spmvb200_matrix_t m, spmvb200_info * info)
try {
    if (!m || !info) return fail(SPMVB200_ERR_INVALID, "null argument");
    memset(info, 0, sizeof *info);
    info->format = m->format; info->coo_mode = m->coo_mode;
    info->rows = m->rows; info->columns = m->cols; info->num_entries = m->nnz;
    info->stored_entries = m->stored; info->row_alignment = m->row_alignment;
    info->ell_row_length = m->ell_w; info->num_ell_entries = m->n_ell; info->num_coo_entries = m->n_coo;
    info->skip_padding = m->skip_padding; info->offsets_64bit = m->off64 ? 1 : 0;
    switch (m->format) {
    case SPMVB200_CSR: info->matrix_size = 12 * m->stored + 4 * (m->rows + 1); break;  // csr-matrix.cpp:46-60
    case SPMVB200_COO: info->matrix_size = 16 * m->stored; break;                      // coo-matrix.cpp:49-63
    case SPMVB200_ELL: info->matrix_size = 12 * m->rows * m->ell_w; break;             // ell-matrix.cpp:52-65
    case SPMVB200_HYB: info->matrix_size = 12 * m->n_ell + 16 * m->n_coo; break;
    }
    info->x_size = 8 * m->cols; info->y_size = 8 * m->rows;
    info->device_bytes = m->device_bytes; info->row_offset = m->row_offset;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_csr_export(spmvb200_matrix_t m, int64_t * row_ptr, int32_t * column_index, double * value)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_CSR) return fail(SPMVB200_ERR_INVALID, "not a CSR matrix");
    if (column_index || value) SPMV_TRY(csr_ensure_row_major(m));
    cudaStream_t s = m->stream;
    if (row_ptr) {
        if (m->off64) {
            SPMV_CUDA(cudaMemcpyAsync(row_ptr, m->rp, sizeof(int64_t) * (size_t)(m->rows + 1), cudaMemcpyDeviceToHost, s));
            SPMV_CUDA(cudaStreamSynchronize(s));
        } else {
            std::vector<uint32_t> t((size_t)m->rows + 1);
            SPMV_CUDA(cudaMemcpyAsync(t.data(), m->rp, sizeof(uint32_t) * t.size(), cudaMemcpyDeviceToHost, s));
            SPMV_CUDA(cudaStreamSynchronize(s));
            for (size_t r = 0; r < t.size(); r++) row_ptr[r] = t[r];
        }
    }
    if (column_index && m->stored) SPMV_CUDA(cudaMemcpyAsync(column_index, m->col, sizeof(int32_t) * (size_t)m->stored, cudaMemcpyDeviceToHost, s));
    if (value && m->stored) SPMV_CUDA(cudaMemcpyAsync(value, m->val, sizeof(double) * (size_t)m->stored, cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}
SPMV_ABI_CATCH

// device_order = false: the row-major order of the reference (a column-blocked matrix is sorted back first)
static int export_coo_arrays(Matrix * m, int32_t * row, int32_t * col, double * val, bool device_order = false)
{
    cudaStream_t s = m->stream;
    const size_t n = (size_t)m->coo_n;
    if (n) {
        const int32_t * dr = m->coo_row;
        const int32_t * dc = m->coo_col;
        const double * dv = m->coo_val;
        Scratch<int32_t> r2, c2;
        Scratch<double> v2;
        if (m->coo_col_shift > 0 && !device_order) {
            SPMV_TRY(r2.alloc(m->coo_n)); SPMV_TRY(c2.alloc(m->coo_n)); SPMV_TRY(v2.alloc(m->coo_n));
            SPMV_TRY(coo_row_major_copy(m, r2.p, c2.p, v2.p));
            dr = r2.p; dc = c2.p; dv = v2.p;
        }
        if (row) SPMV_CUDA(cudaMemcpyAsync(row, dr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
        if (col) SPMV_CUDA(cudaMemcpyAsync(col, dc, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
        if (val) SPMV_CUDA(cudaMemcpyAsync(val, dv, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

static int export_ell_arrays(Matrix * m, int32_t * col_rm, double * val_rm)
{
    cudaStream_t s = m->stream;
    const int64_t n = m->rows * m->ell_w;
    if (n == 0) return 0;
    Scratch<int32_t> dc;
    Scratch<double> dv;
    SPMV_TRY(dc.alloc(n)); SPMV_TRY(dv.alloc(n));
    transpose_to_rowmajor_kernel<<<grid_for(n), 256, 0, s>>>(m->rows, m->ell_w, m->ell_pitch, m->ell_col, m->ell_val, dc.p, dv.p);
    SPMV_CUDA(cudaGetLastError());
    if (col_rm) SPMV_CUDA(cudaMemcpyAsync(col_rm, dc.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (val_rm) SPMV_CUDA(cudaMemcpyAsync(val_rm, dv.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int spmvb200_coo_export(spmvb200_matrix_t m, int32_t * row_index, int32_t * column_index, double * value)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_COO) return fail(SPMVB200_ERR_INVALID, "not a COO matrix");
    return export_coo_arrays(m, row_index, column_index, value);
}
SPMV_ABI_CATCH

int spmvb200_ell_export(spmvb200_matrix_t m, int32_t * column_index, double * value)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_ELL) return fail(SPMVB200_ERR_INVALID, "not an ELL matrix");
    return export_ell_arrays(m, column_index, value);
}
SPMV_ABI_CATCH

int spmvb200_hyb_export(spmvb200_matrix_t m, int32_t * ell_column_index, double * ell_value, int32_t * coo_row_index,
                        int32_t * coo_column_index, double * coo_value)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_HYB) return fail(SPMVB200_ERR_INVALID, "not a hybrid matrix");
    SPMV_TRY(export_ell_arrays(m, ell_column_index, ell_value));
    return export_coo_arrays(m, coo_row_index, coo_column_index, coo_value);
}
SPMV_ABI_CATCH

// ---- vectors ---------------------------------------------------------------------------------------------------

int spmvb200_set_x(spmvb200_matrix_t m, const double * x)
try {
    SPMV_TRY(check(m));
    if (!x) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_CUDA(cudaMemcpyAsync(m->x, x, sizeof(double) * (size_t)m->cols, cudaMemcpyHostToDevice, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_set_y(spmvb200_matrix_t m, const double * y)
try {
    SPMV_TRY(check(m));
    if (!y) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_CUDA(cudaMemcpyAsync(m->y, y, sizeof(double) * (size_t)m->rows, cudaMemcpyHostToDevice, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_get_x(spmvb200_matrix_t m, double * x)
try {
    SPMV_TRY(check(m));
    if (!x) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_CUDA(cudaMemcpyAsync(x, m->x, sizeof(double) * (size_t)m->cols, cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_get_y(spmvb200_matrix_t m, double * y)
try {
    SPMV_TRY(check(m));
    if (!y) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_CUDA(cudaMemcpyAsync(y, m->y, sizeof(double) * (size_t)m->rows, cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_fill_x(spmvb200_matrix_t m, double v)
try {
    SPMV_TRY(check(m));
    return fill_device(m->x, m->cols, v, m->stream);
}
SPMV_ABI_CATCH
int spmvb200_fill_y(spmvb200_matrix_t m, double v)
try {
    SPMV_TRY(check(m));
    return fill_device(m->y, m->rows, v, m->stream);
}
SPMV_ABI_CATCH
int spmvb200_x_device(spmvb200_matrix_t m, void ** p)
try {
    if (!m || !p) return fail(SPMVB200_ERR_INVALID, "null argument");
    *p = m->x;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_y_device(spmvb200_matrix_t m, void ** p)
try {
    if (!m || !p) return fail(SPMVB200_ERR_INVALID, "null argument");
    *p = m->y;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_bind_x(spmvb200_matrix_t m, void * p)
try {
    SPMV_TRY(check(m));
    if (!p) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (m->own_x) { cudaFree(m->x); m->device_bytes -= 8 * (m->cols + 8); }
    m->x = (double *)p; m->own_x = false;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_bind_y(spmvb200_matrix_t m, void * p)
try {
    SPMV_TRY(check(m));
    if (!p) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (m->own_y) { cudaFree(m->y); m->device_bytes -= 8 * (m->rows + 8); }
    m->y = (double *)p; m->own_y = false;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_set_stream(spmvb200_matrix_t m, void * stream)
try {
    SPMV_TRY(check(m));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    if (m->own_stream) {
        stream_forget(m->stream);
        cudaStreamDestroy(m->stream);
    }
    m->stream = (cudaStream_t)stream; m->own_stream = false;
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_host_alloc(size_t bytes, void ** p)
try {
    if (!p) return fail(SPMVB200_ERR_INVALID, "null argument");
    SPMV_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
    return 0;
}
SPMV_ABI_CATCH
int spmvb200_host_free(void * p)
try {
    SPMV_CUDA(cudaFreeHost(p));
    return 0;
}
SPMV_ABI_CATCH

// ---- run ----------------------------------------------------------------------------------------------------------

// The timing entry points run several matrices on one stream.  The matrices keep owning their own streams: the
// loan is undone when the scope ends, on the error paths too (a matrix left pointing at a stream it does not own
// would destroy it a second time in matrix_free and leak its own).
struct StreamLoan {
    const spmvb200_matrix_t * ms;
    int n;
    std::vector<cudaStream_t> saved;
    StreamLoan(const spmvb200_matrix_t * ms, int n, cudaStream_t s) : ms(ms), n(n), saved((size_t)n)
    {
        for (int k = 0; k < n; k++) { saved[(size_t)k] = ms[k]->stream; ms[k]->stream = s; }
    }
    ~StreamLoan() { for (int k = 0; k < n; k++) ms[k]->stream = saved[(size_t)k]; }
    StreamLoan(const StreamLoan &) = delete;
    StreamLoan & operator=(const StreamLoan &) = delete;
};

static int launch_format(Matrix * m)
{
    switch (m->format) {
    case SPMVB200_CSR: return launch_csr(m);
    case SPMVB200_ELL: return launch_ell(m, true);
    case SPMVB200_COO: return launch_coo(m);
    case SPMVB200_HYB:
        // ELL pass, then the COO tail adds into the same y (hybrid-matrix.cpp:547-566); both only
        // add to y, so the second kernel need not wait for the first
        SPMV_TRY(launch_ell(m, true));
        if (m->coo_n > 0) plan_run(m, false);
        SPMV_TRY(launch_coo(m));
        m->kernel_name = m->opt_coo_algo == 1 ? "ell_kernel+coo_segmented_kernel" : m->opt_coo_algo == 4 ? "ell_kernel+coo_warp_kernel"
                         : (m->coo_n > 0 && !strcmp(m->kernel_name, "coo_hot_kernel")) ? "ell_kernel+coo_hot_kernel" : "ell_kernel+coo_warp4_kernel";
        return 0;
    }
    return fail(SPMVB200_ERR_INVALID, "unknown format");
}

// One y (+)= alpha*A*x.  "beta0" (y = ...): kernels in which one thread owns a whole row (ELL, sliced CSR) store
// their result; the others clear y first (clear_y_for_beta0) and add with reductions as always.
static int launch(Matrix * m)
{
    m->run_beta0 = m->opt_beta0 != 0;
    // "csr.rmw" = 1: y += A*x with the sliced CSR kernel, the lane that owns a row adds to y with a plain load and store
    // instead of a reduction (the reductions make L2 write every y sector back twice: 1.97 GB written for 1.07 GB of y on
    // config 5); such launches must be ordered.  Opt-in: it measured SLOWER (config 5: 7.69 vs 7.41 ms; 27-point 256^3:
    // 0.929 vs 0.908 ms) -- the load of y_old sits at the end of every lane's dependency chain, the reduction does not.
    m->run_rmw = !m->run_beta0 && m->format == SPMVB200_CSR && m->opt_csr_rmw > 0 && csr_uses_sliced_kernel(m);
    plan_run(m, false);
    const int rc = launch_format(m);
    m->run_beta0 = false;
    m->run_rmw = false;
    return rc;
}

int spmvb200_set_alpha(spmvb200_matrix_t m, double alpha)
try {
    if (!m) return fail(SPMVB200_ERR_INVALID, "null matrix handle");
    m->alpha = alpha;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_prepare(spmvb200_matrix_t m)
try {
    SPMV_TRY(check(m));
    if (m->format == SPMVB200_ELL) return 0;  // the ELL kernel keeps no launch metadata
    m->dry_run = true;
    const int rc = m->format == SPMVB200_CSR ? launch_csr(m) : launch_coo(m);  // COO / hybrid tail: hot-column tables
    m->dry_run = false;
    if (rc) return rc;
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    m->aux_dirty = false;
    stream_synced(m->stream);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_spmv(spmvb200_matrix_t m)
try {
    SPMV_TRY(check_run(m));
    return launch(m);
}
SPMV_ABI_CATCH

int spmvb200_sync(spmvb200_matrix_t m)
try {
    SPMV_TRY(check(m));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    stream_synced(m->stream);
    return 0;
}
SPMV_ABI_CATCH

// Host-buffer form of y += A*x.  The bytes that must cross PCIe are fixed (x and y up, y down), so
// the only lever is to use both directions at once: for ELL the rows are cut into chunks; x goes up
// first, then the y chunks go up on a second stream while, chunk by chunk, the kernel runs and the
// finished part of y comes back down on the compute stream (H2D and D2H copy engines overlap).
// Other formats, tiny matrices and pageable buffers take the plain sequence.
static int spmv_host_pipelined(Matrix * m, const double * x, double * y, int chunks, double * y_dev_visible)
{
    cudaStream_t s = m->stream;
    if (!m->upload_stream) {
        SPMV_CUDA(cudaStreamCreateWithFlags(&m->upload_stream, cudaStreamNonBlocking));
        SPMV_CUDA(cudaEventCreateWithFlags(&m->ev_x, cudaEventDisableTiming));
        for (auto & e : m->ev_chunk) SPMV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t up = m->upload_stream;
    // Kernels queued earlier on the matrix's stream (an asynchronous spmvb200_spmv) may still gather from m->x and
    // add to m->y: the uploads below must not overtake them.
    SPMV_CUDA(cudaEventRecord(m->ev_x, s));
    SPMV_CUDA(cudaStreamWaitEvent(up, m->ev_x, 0));
    const int64_t per = round_up((m->rows + chunks - 1) / chunks, 1024);
    SPMV_CUDA(cudaMemcpyAsync(m->x, x, sizeof(double) * (size_t)m->cols, cudaMemcpyHostToDevice, up));
    SPMV_CUDA(cudaEventRecord(m->ev_x, up));
    int used = 0;
    for (int c = 0; c < chunks; c++) {
        const int64_t b = (int64_t)c * per, e = std::min<int64_t>(m->rows, b + per);
        if (b >= e) break;
        if (m->opt_beta0) SPMV_CUDA(cudaMemsetAsync(m->y + b, 0, sizeof(double) * (size_t)(e - b), up));
        else SPMV_CUDA(cudaMemcpyAsync(m->y + b, y + b, sizeof(double) * (size_t)(e - b), cudaMemcpyHostToDevice, up));
        SPMV_CUDA(cudaEventRecord(m->ev_chunk[c], up));
        used = c + 1;
    }
    SPMV_CUDA(cudaStreamWaitEvent(s, m->ev_x, 0));
    const int64_t keep_beta0 = m->opt_beta0;
    m->opt_beta0 = 0;  // y was prepared chunk by chunk above
    int rc = 0;
    for (int c = 0; c < used && rc == 0; c++) {
        const int64_t b = (int64_t)c * per, e = std::min<int64_t>(m->rows, b + per);
        cudaStreamWaitEvent(s, m->ev_chunk[c], 0);
        m->range_begin = b;
        m->range_end = e;
        if (y_dev_visible) {  // the kernel adds onto the uploaded chunk and stores the result in host memory
            m->host_y_in = m->y;
            m->host_y_out = y_dev_visible;
        }
        plan_run(m, true);
        rc = launch_ell(m, true);
        m->host_y_in = nullptr;
        m->host_y_out = nullptr;
        if (!y_dev_visible && rc == 0 &&
            cudaMemcpyAsync(y + b, m->y + b, sizeof(double) * (size_t)(e - b), cudaMemcpyDeviceToHost, s) != cudaSuccess)
            rc = fail(SPMVB200_ERR_CUDA, "cudaMemcpyAsync(y chunk)");
    }
    m->range_begin = m->range_end = 0;
    m->opt_beta0 = keep_beta0;
    if (rc) return rc;
    SPMV_CUDA(cudaStreamSynchronize(s));
    stream_synced(s);
    return 0;
}

// Host-buffer form for the sliced CSR kernel (a lane owns its row, so y is read from and written to pinned host memory by
// the kernel itself).  What is left to overlap is the upload of x: it goes up in `chunks` pieces, and the rows are run in
// as many chunks, each launched as soon as the largest column its rows reference has arrived.  For a banded matrix
// (stencils) chunk c needs x up to piece c + 1, so the kernel starts after 2/chunks of the upload instead of all of it
// and the whole call costs max(x + y_old up, y_new down) over PCIe; a matrix whose rows reference every column degenerates
// to upload-then-run.
static int spmv_host_csr_pipelined(Matrix * m, const double * x, double * y_dev_visible, int chunks, const double * y_dma,
                                   double * y_back)
{
    cudaStream_t s = m->stream;
    if (!m->upload_stream) {
        SPMV_CUDA(cudaStreamCreateWithFlags(&m->upload_stream, cudaStreamNonBlocking));
        SPMV_CUDA(cudaEventCreateWithFlags(&m->ev_x, cudaEventDisableTiming));
    }
    for (int c = 0; c < 2 * chunks; c++)  // [0, chunks): x pieces; [chunks, 2 chunks): y chunks of the DMA form
        if (!m->ev_chunk[c]) SPMV_CUDA(cudaEventCreateWithFlags(&m->ev_chunk[c], cudaEventDisableTiming));
    const int64_t per = round_up((m->rows + chunks - 1) / chunks, 1024);
    if (m->host_chunks != chunks || m->host_rows_per_chunk != per) {
        SPMV_TRY(csr_chunk_colmax(m, per, chunks, m->host_colmax));
        m->host_chunks = chunks;
        m->host_rows_per_chunk = per;
    }
    cudaStream_t up = m->upload_stream;
    SPMV_CUDA(cudaEventRecord(m->ev_x, s));  // kernels queued earlier may still gather from m->x
    SPMV_CUDA(cudaStreamWaitEvent(up, m->ev_x, 0));
    const int64_t keep = m->opt_beta0;
    // "host.zero_copy" = 2 (y_dma): y_old goes up with the copy engine as well -- chunk c right behind the x piece its rows
    // wait for -- and the kernel adds onto the uploaded chunk; only y_new is written by the kernel.  One DMA queue then
    // carries everything that goes up, in the order the kernels need it.
    const bool dma = y_dma && !keep;
    std::vector<int> need_of((size_t)chunks, -1);
    for (int c = 0; c < chunks; c++)
        need_of[(size_t)c] = m->host_colmax[c] < 0 ? -1 : (int)std::min<int64_t>(chunks - 1, m->host_colmax[c] / per);
    int next_y = 0;  // next y chunk to send up
    auto send_y_upto = [&](int piece) -> int {  // every y chunk whose rows need no x piece beyond `piece`
        while (dma && next_y < chunks && need_of[(size_t)next_y] <= piece) {
            const int64_t b = (int64_t)next_y * per, e = std::min<int64_t>(m->rows, b + per);
            if (e > b) SPMV_CUDA(cudaMemcpyAsync(m->y + b, y_dma + b, sizeof(double) * (size_t)(e - b), cudaMemcpyHostToDevice, up));
            SPMV_CUDA(cudaEventRecord(m->ev_chunk[chunks + next_y], up));
            ++next_y;
        }
        return 0;
    };
    for (int c = 0; c < chunks; c++) {  // x pieces on the same grid as the row chunks (the matrix is square)
        const int64_t b = (int64_t)c * per, e = std::min<int64_t>(m->cols, b + per);
        if (e > b) SPMV_CUDA(cudaMemcpyAsync(m->x + b, x + b, sizeof(double) * (size_t)(e - b), cudaMemcpyHostToDevice, up));
        SPMV_CUDA(cudaEventRecord(m->ev_chunk[c], up));
        SPMV_TRY(send_y_upto(c));
    }
    SPMV_TRY(send_y_upto(chunks));
    // "host.zero_copy" = 3 (y_back): the result comes down with the copy engine too; the kernel works on device memory only
    const bool back = dma && y_back;
    m->host_y_in = back || keep ? nullptr : dma ? (const double *)m->y : (const double *)y_dev_visible;
    m->host_y_out = back ? nullptr : y_dev_visible;
    m->opt_beta0 = 0;
    int rc = 0, waited = -1;
    for (int c = 0; c < chunks && rc == 0; c++) {
        const int64_t b = (int64_t)c * per, e = std::min<int64_t>(m->rows, b + per);
        if (b >= e) break;
        const int need = need_of[(size_t)c];
        if (dma) {  // the chunk's y went up behind the x piece it needs: one wait covers both
            if (cudaStreamWaitEvent(s, m->ev_chunk[chunks + c], 0) != cudaSuccess) { rc = fail(SPMVB200_ERR_CUDA, "cudaStreamWaitEvent"); break; }
        } else
        if (need > waited) {
            if (cudaStreamWaitEvent(s, m->ev_chunk[need], 0) != cudaSuccess) { rc = fail(SPMVB200_ERR_CUDA, "cudaStreamWaitEvent"); break; }
            waited = need;
        }
        m->range_begin = b;
        m->range_end = e;
        plan_run(m, true);
        rc = launch_csr_sliced(m);
        if (back && rc == 0 &&
            cudaMemcpyAsync(y_back + b, m->y + b, sizeof(double) * (size_t)(e - b), cudaMemcpyDeviceToHost, s) != cudaSuccess)
            rc = fail(SPMVB200_ERR_CUDA, "cudaMemcpyAsync(y chunk)");
    }
    m->range_begin = m->range_end = 0;
    m->opt_beta0 = keep;
    m->host_y_in = nullptr;
    m->host_y_out = nullptr;
    if (rc) return rc;
    SPMV_CUDA(cudaStreamWaitEvent(s, m->ev_chunk[chunks - 1], 0));  // the call owns x until every piece has landed
    SPMV_CUDA(cudaStreamSynchronize(s));
    stream_synced(s);
    return 0;
}

int spmvb200_spmv_host(spmvb200_matrix_t m, const double * x, double * y)
try {
    SPMV_TRY(check(m));
    if (!x || !y) return fail(SPMVB200_ERR_INVALID, "null argument");
    cudaStream_t s = m->stream;
    if (m->format == SPMVB200_ELL && m->opt_host_zero_copy && m->rows > 0 && m->ell_w > 0) {
        // pinned (page-locked, mapped) buffers: x goes up with the copy engine, then ONE kernel reads
        // y_old from and writes y_new to host memory while it streams the matrix from HBM
        cudaPointerAttributes ay{};
        if (cudaPointerGetAttributes(&ay, y) == cudaSuccess && ay.type == cudaMemoryTypeHost && ay.devicePointer) {
            if (m->opt_host_zero_copy == 2 && m->rows >= 64 * 1024)  // y up by DMA in chunks, results stored by the kernel
                return spmv_host_pipelined(m, x, y, (int)(m->opt_host_chunks ? std::min<int64_t>(m->opt_host_chunks, 16) : 4),
                                           (double *)ay.devicePointer);
            SPMV_CUDA(cudaMemcpyAsync(m->x, x, sizeof(double) * (size_t)m->cols, cudaMemcpyHostToDevice, s));
            m->host_y_in = m->opt_beta0 ? nullptr : (const double *)ay.devicePointer;
            m->host_y_out = (double *)ay.devicePointer;
            const int64_t keep = m->opt_beta0;
            m->opt_beta0 = 0;
            plan_run(m, true);
            const int rc = launch_ell(m, true);
            m->opt_beta0 = keep;
            m->host_y_in = nullptr;
            m->host_y_out = nullptr;
            if (rc) return rc;
            SPMV_CUDA(cudaStreamSynchronize(s));
            stream_synced(s);
            return 0;
        }
        cudaGetLastError();  // not a registered host pointer: fall through to the copying paths
    }
    if (m->format == SPMVB200_CSR && m->opt_host_zero_copy && m->rows > 0 && m->stored > 0 && csr_uses_sliced_kernel(m)) {
        // the sliced CSR kernel owns whole rows like the ELL kernel: same zero-copy form (x up by DMA, then one kernel
        // that reads y_old from and writes y_new to the pinned host buffer while it streams the matrix from HBM)
        cudaPointerAttributes ay{};
        if (!m->slice_col) {  // build the slot-major copy now; if it does not fit the flat kernel runs and y is copied
            m->dry_run = true;
            const int rc = launch_csr(m);
            m->dry_run = false;
            if (rc) return rc;
        }
        if (m->slice_col && cudaPointerGetAttributes(&ay, y) == cudaSuccess && ay.type == cudaMemoryTypeHost && ay.devicePointer) {
            const int chunks = (int)(m->opt_host_chunks ? std::min<int64_t>(m->opt_host_chunks, 32) : m->rows >= 32 * (int64_t)65536 ? 32 : 16);
            // form of the pipelined call: 1 = automatic = 3, everything by copy engines (config 5: 43.6 ms against 44.8 ms for
            // form 4, where the kernel reads y_old from and writes y_new to the pinned buffer); 2 = y_old up by DMA, y_new
            // stored by the kernel
            const int64_t form = m->opt_host_zero_copy == 1 ? 3 : m->opt_host_zero_copy;
            if (chunks > 1 && m->rows >= (int64_t)chunks * 65536 && m->rows == m->cols)
                return spmv_host_csr_pipelined(m, x, (double *)ay.devicePointer, chunks, form == 2 || form == 3 ? y : nullptr,
                                               form == 3 ? y : nullptr);
            SPMV_CUDA(cudaMemcpyAsync(m->x, x, sizeof(double) * (size_t)m->cols, cudaMemcpyHostToDevice, s));
            m->host_y_in = m->opt_beta0 ? nullptr : (const double *)ay.devicePointer;
            m->host_y_out = (double *)ay.devicePointer;
            const int64_t keep = m->opt_beta0;
            m->opt_beta0 = 0;
            plan_run(m, true);
            const int rc = launch_csr_sliced(m);
            m->opt_beta0 = keep;
            m->host_y_in = nullptr;
            m->host_y_out = nullptr;
            if (rc) return rc;
            SPMV_CUDA(cudaStreamSynchronize(s));
            stream_synced(s);
            return 0;
        }
        cudaGetLastError();
    }
    const int chunks = (int)(m->opt_host_chunks ? std::min<int64_t>(m->opt_host_chunks, 16) : 4);
    if (m->format == SPMVB200_ELL && chunks > 1 && m->rows >= 64 * 1024) return spmv_host_pipelined(m, x, y, chunks, nullptr);
    SPMV_CUDA(cudaMemcpyAsync(m->x, x, sizeof(double) * (size_t)m->cols, cudaMemcpyHostToDevice, s));
    if (!m->opt_beta0) SPMV_CUDA(cudaMemcpyAsync(m->y, y, sizeof(double) * (size_t)m->rows, cudaMemcpyHostToDevice, s));
    SPMV_TRY(launch(m));
    SPMV_CUDA(cudaMemcpyAsync(y, m->y, sizeof(double) * (size_t)m->rows, cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    stream_synced(s);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_time(spmvb200_matrix_t m, int warmup, int reps, float * ms)
try {
    SPMV_TRY(check(m));
    if (reps < 0 || warmup < 0 || (reps > 0 && !ms)) return fail(SPMVB200_ERR_INVALID, "bad argument");
    for (int w = 0; w < warmup; w++) SPMV_TRY(launch(m));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    for (int r = 0; r < reps; r++) {
        SPMV_CUDA(cudaEventRecord(m->ev0, m->stream));
        SPMV_TRY(launch(m));
        SPMV_CUDA(cudaEventRecord(m->ev1, m->stream));
        SPMV_CUDA(cudaEventSynchronize(m->ev1));
        SPMV_CUDA(cudaEventElapsedTime(&ms[r], m->ev0, m->ev1));
    }
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_time_rotating(const spmvb200_matrix_t * ms, int n, int warmup, int steps, float * total_ms,
                           float * per_launch_ms)
try {
    if (!ms || n < 1 || steps < 1 || warmup < 0 || !total_ms) return fail(SPMVB200_ERR_INVALID, "bad argument");
    for (int k = 0; k < n; k++) SPMV_TRY(check(ms[k]));
    cudaStream_t s = ms[0]->stream;
    for (int k = 0; k < n; k++) SPMV_CUDA(cudaStreamSynchronize(ms[k]->stream));
    StreamLoan loan(ms, n, s);  // every matrix launches on ms[0]'s stream until this scope ends, however it ends
    int rc = 0;
    for (int k = 0; k < n; k++) SPMV_TRY(spmvb200_prepare(ms[k]));
    auto run = [&]() -> int {
        for (int w = 0; w < warmup; w++) SPMV_TRY(launch(ms[w % n]));
        SPMV_CUDA(cudaStreamSynchronize(s));
        SPMV_CUDA(cudaEventRecord(ms[0]->ev0, s));
        for (int k = 0; k < steps; k++) SPMV_TRY(launch(ms[(warmup + k) % n]));
        SPMV_CUDA(cudaEventRecord(ms[0]->ev1, s));
        SPMV_CUDA(cudaEventSynchronize(ms[0]->ev1));
        SPMV_CUDA(cudaEventElapsedTime(total_ms, ms[0]->ev0, ms[0]->ev1));
        if (per_launch_ms) {
            std::vector<cudaEvent_t> ev((size_t)steps + 1);
            for (auto & e : ev) SPMV_CUDA(cudaEventCreate(&e));
            SPMV_CUDA(cudaEventRecord(ev[0], s));
            for (int k = 0; k < steps; k++) {
                SPMV_TRY(launch(ms[(warmup + k) % n]));
                SPMV_CUDA(cudaEventRecord(ev[(size_t)k + 1], s));
            }
            SPMV_CUDA(cudaEventSynchronize(ev[(size_t)steps]));
            for (int k = 0; k < steps; k++) SPMV_CUDA(cudaEventElapsedTime(&per_launch_ms[k], ev[k], ev[(size_t)k + 1]));
            for (auto & e : ev) cudaEventDestroy(e);
        }
        return 0;
    };
    rc = run();
    if (rc == 0) stream_synced(s);
    return rc;
}
SPMV_ABI_CATCH

int spmvb200_time_host_rotating(const spmvb200_matrix_t * ms, int n, const double * const * xs, double * const * ys,
                                int warmup, int steps, float * total_ms)
try {
    if (!ms || !xs || !ys || n < 1 || steps < 1 || warmup < 0 || !total_ms) return fail(SPMVB200_ERR_INVALID, "bad argument");
    for (int k = 0; k < n; k++) SPMV_TRY(check(ms[k]));
    cudaStream_t s = ms[0]->stream;
    for (int k = 0; k < n; k++) SPMV_CUDA(cudaStreamSynchronize(ms[k]->stream));
    StreamLoan loan(ms, n, s);
    auto run = [&]() -> int {
        for (int w = 0; w < warmup; w++) SPMV_TRY(spmvb200_spmv_host(ms[w % n], xs[w % n], ys[w % n]));
        SPMV_CUDA(cudaEventRecord(ms[0]->ev0, s));
        for (int k = 0; k < steps; k++) {
            const int c = (warmup + k) % n;
            SPMV_TRY(spmvb200_spmv_host(ms[c], xs[c], ys[c]));
        }
        SPMV_CUDA(cudaEventRecord(ms[0]->ev1, s));
        SPMV_CUDA(cudaEventSynchronize(ms[0]->ev1));
        SPMV_CUDA(cudaEventElapsedTime(total_ms, ms[0]->ev0, ms[0]->ev1));
        return 0;
    };
    return run();
}
SPMV_ABI_CATCH

// Yardstick for "isolated launch" figures: the same event-pair protocol around a plain device-to-device cudaMemcpyAsync
// that moves `bytes` (read + write = 2 * bytes of DRAM traffic), rotating over `copies` source/destination pairs so that
// nothing is found in L2.  What this takes is what the protocol + one DRAM round trip cost with no kernel of ours involved.
int spmvb200_time_copy(int64_t bytes, int copies, int warmup, int reps, float * ms)
try {
    if (bytes < 1 || copies < 1 || warmup < 0 || reps < 1 || !ms) return fail(SPMVB200_ERR_INVALID, "bad argument");
    std::vector<char *> src((size_t)copies, nullptr), dst((size_t)copies, nullptr);
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    auto cleanup = [&] {
        for (char * p : src) if (p) cudaFree(p);
        for (char * p : dst) if (p) cudaFree(p);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (s) cudaStreamDestroy(s);
    };
    auto run = [&]() -> int {
        SPMV_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        SPMV_CUDA(cudaEventCreate(&e0));
        SPMV_CUDA(cudaEventCreate(&e1));
        for (int c = 0; c < copies; c++) {
            SPMV_CUDA(cudaMalloc((void **)&src[(size_t)c], (size_t)bytes));
            SPMV_CUDA(cudaMalloc((void **)&dst[(size_t)c], (size_t)bytes));
            SPMV_CUDA(cudaMemsetAsync(src[(size_t)c], 1, (size_t)bytes, s));
        }
        for (int w = 0; w < warmup; w++)
            SPMV_CUDA(cudaMemcpyAsync(dst[(size_t)(w % copies)], src[(size_t)(w % copies)], (size_t)bytes, cudaMemcpyDeviceToDevice, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        for (int r = 0; r < reps; r++) {
            const size_t c = (size_t)((warmup + r) % copies);
            SPMV_CUDA(cudaEventRecord(e0, s));
            SPMV_CUDA(cudaMemcpyAsync(dst[c], src[c], (size_t)bytes, cudaMemcpyDeviceToDevice, s));
            SPMV_CUDA(cudaEventRecord(e1, s));
            SPMV_CUDA(cudaEventSynchronize(e1));
            SPMV_CUDA(cudaEventElapsedTime(&ms[r], e0, e1));
        }
        return 0;
    };
    rc = run();
    cleanup();
    return rc;
}
SPMV_ABI_CATCH

static int64_t * option_slot(Matrix * m, const char * key)
{
    if (!strcmp(key, "csr.tile")) return &m->opt_csr_tile;
    if (!strcmp(key, "csr.stages")) return &m->opt_csr_stages;
    if (!strcmp(key, "csr.threads")) return &m->opt_csr_threads;
    if (!strcmp(key, "pdl")) return &m->opt_pdl;
    if (!strcmp(key, "independent_launches")) return &m->opt_independent;
    if (!strcmp(key, "csr.algo")) return &m->opt_csr_algo;
    if (!strcmp(key, "csr.lanes")) return &m->opt_csr_lanes;
    if (!strcmp(key, "csr.ctas_per_sm")) return &m->opt_csr_ctas;
    if (!strcmp(key, "csr.spare_ctas")) return &m->opt_csr_spare;
    if (!strcmp(key, "csr.batch")) return &m->opt_csr_batch;
    if (!strcmp(key, "csr.probe")) return &m->opt_csr_probe;
    if (!strcmp(key, "csr.rmw")) return &m->opt_csr_rmw;
    if (!strcmp(key, "csr.entries")) return &m->opt_csr_entries;
    if (!strcmp(key, "csr.rowptr_path")) return &m->opt_csr_rowptr_path;
    if (!strcmp(key, "csr.drop_row_major")) return &m->opt_csr_drop;
    if (!strcmp(key, "csr.index_runs")) return &m->opt_csr_index_runs;
    if (!strcmp(key, "csr.regs")) return &m->opt_csr_regs;
    if (!strcmp(key, "ell.rows_per_thread")) return &m->opt_ell_rows;
    if (!strcmp(key, "ell.block")) return &m->opt_ell_block;
    if (!strcmp(key, "coo.stages")) return &m->opt_coo_stages;
    if (!strcmp(key, "coo.threads")) return &m->opt_coo_threads;
    if (!strcmp(key, "coo.ctas_per_sm")) return &m->opt_coo_ctas;
    if (!strcmp(key, "coo.algo")) return &m->opt_coo_algo;
    if (!strcmp(key, "coo.items")) return &m->opt_coo_items;
    if (!strcmp(key, "coo.xload")) return &m->opt_coo_xload;
    if (!strcmp(key, "coo.carveout")) return &m->opt_coo_carveout;
    if (!strcmp(key, "coo.hot")) return &m->opt_coo_hot;
    if (!strcmp(key, "coo.hot_slots")) return &m->opt_coo_hot_slots;
    if (!strcmp(key, "coo.hot_threads")) return &m->opt_coo_hot_threads;
    if (!strcmp(key, "coo.hot_entries")) return &m->opt_coo_hot_entries;
    if (!strcmp(key, "coo.hot_segments")) return &m->opt_coo_hot_segs;
    if (!strcmp(key, "beta0")) return &m->opt_beta0;
    if (!strcmp(key, "host.chunks")) return &m->opt_host_chunks;
    if (!strcmp(key, "host.zero_copy")) return &m->opt_host_zero_copy;
    return nullptr;
}

int spmvb200_set_option(spmvb200_matrix_t m, const char * key, int64_t value)
try {
    if (!m || !key) return fail(SPMVB200_ERR_INVALID, "null argument");
    int64_t * slot = option_slot(m, key);
    if (!slot) return fail(SPMVB200_ERR_INVALID, std::string("unknown option ") + key);
    if (*slot != value && (slot == &m->opt_coo_hot || slot == &m->opt_coo_hot_slots || slot == &m->opt_coo_hot_threads ||
                           slot == &m->opt_coo_hot_segs) && m->coo_hot_tried) {
        // the tables were built for the old shape: drop them, the next launch / prepare rebuilds
        SPMV_CUDA(cudaSetDevice(m->device));
        SPMV_CUDA(cudaStreamSynchronize(m->stream));
        if (m->coo_colh) {
            const int64_t cap = round_up(m->coo_n, 4096) + kPadEntries;
            m->device_bytes -= cap * 4 + (int64_t)m->coo_nseg * m->coo_hot_h * 4 + ((int64_t)m->coo_nseg + 1) * 8;
        }
        cudaFree(m->coo_colh); cudaFree(m->coo_hot_cols); cudaFree(m->coo_seg);
        m->coo_colh = nullptr; m->coo_hot_cols = nullptr; m->coo_seg = nullptr;
        m->coo_hot_tried = false;
    }
    if (*slot != value && slot == &m->opt_csr_index_runs && m->slice_col) {
        // the slot-major copy was built under the old setting: back to row-major, the next launch / prepare rebuilds it
        SPMV_CUDA(cudaSetDevice(m->device));
        SPMV_TRY(csr_drop_sliced(m));
    }
    *slot = value;
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_get_option(spmvb200_matrix_t m, const char * key, int64_t * value)
try {
    if (!m || !key || !value) return fail(SPMVB200_ERR_INVALID, "null argument");
    if (!strcmp(key, "coo.col_block_log2")) {  // read-only: the column-block size the builder applied (0 = none)
        *value = m->coo_col_shift;
        return 0;
    }
    if (!strcmp(key, "coo.hot_coverage_permille")) {  // read-only: share of the gathers served from the hot-column tables
        *value = m->coo_colh ? (int64_t)(1000.0 * m->coo_hot_coverage + 0.5) : 0;
        return 0;
    }
    if (!strcmp(key, "csr.index_runs_active")) {  // read-only: the sliced kernel's column stream is stored with index runs
        *value = m->slice_col && m->slice_runs ? 1 : 0;
        return 0;
    }
    if (!strcmp(key, "csr.index_columns_stored")) {  // read-only: int32 entries of that stream (= stored entries when plain)
        *value = m->slice_col ? m->slice_ccount : 0;
        return 0;
    }
    if (!strcmp(key, "coo.hot_segments_built")) {
        *value = m->coo_colh ? m->coo_nseg : 0;
        return 0;
    }
    if (!strcmp(key, "last_launch.overlapped")) {  // read-only: the last kernel was allowed to skip griddepcontrol.wait
        *value = m->run_independent ? 1 : 0;
        return 0;
    }
    if (!strcmp(key, "last_launch.pdl")) {  // read-only: ... was launched with the PDL attribute
        *value = m->run_pdl ? 1 : 0;
        return 0;
    }
    int64_t * slot = option_slot(m, key);
    if (!slot) return fail(SPMVB200_ERR_INVALID, std::string("unknown option ") + key);
    *value = *slot;
    return 0;
}
SPMV_ABI_CATCH

const char * spmvb200_kernel_name(spmvb200_matrix_t m) { return m ? m->kernel_name : ""; }

int spmvb200_destroy(spmvb200_matrix_t m)
try {
    matrix_free(m);
    return 0;
}
SPMV_ABI_CATCH

// ---- row partition ---------------------------------------------------------------------------------------------------

int spmvb200_partition_rows_ref(int64_t rows, int32_t parts, int64_t * starts)
try {
    if (rows < 0 || parts < 1 || !starts) return fail(SPMVB200_ERR_INVALID, "bad argument");
    const int64_t rpt = (rows + parts - 1) / parts;  // csr-matrix.cpp:79
    for (int32_t p = 0; p <= parts; p++) starts[p] = std::min<int64_t>(rows, (int64_t)p * rpt);
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_partition_rows_nnz(spmvb200_matrix_t m, int32_t parts, int64_t * starts)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_CSR || parts < 1 || !starts) return fail(SPMVB200_ERR_INVALID, "bad argument");
    Scratch<int64_t> d;
    SPMV_TRY(d.alloc(parts + 1));
    const unsigned grid = (unsigned)((parts + 1 + 63) / 64);
    if (m->off64) partition_nnz_kernel<int64_t><<<grid, 64, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, parts, d.p);
    else partition_nnz_kernel<uint32_t><<<grid, 64, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, parts, d.p);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaMemcpyAsync(starts, d.p, sizeof(int64_t) * (size_t)(parts + 1), cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_partition_rows_weighted(spmvb200_matrix_t m, int32_t parts, int64_t row_weight_q10, int64_t * starts)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_CSR || parts < 1 || !starts || row_weight_q10 < 0 || row_weight_q10 > ((int64_t)1 << 30))
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    Scratch<int64_t> d;
    SPMV_TRY(d.alloc(parts + 1));
    const unsigned grid = (unsigned)((parts + 1 + 63) / 64);
    if (m->off64) partition_weighted_kernel<int64_t><<<grid, 64, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, parts, row_weight_q10, d.p);
    else partition_weighted_kernel<uint32_t><<<grid, 64, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, parts, row_weight_q10, d.p);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaMemcpyAsync(starts, d.p, sizeof(int64_t) * (size_t)(parts + 1), cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_csr_column_span(spmvb200_matrix_t m, int64_t col_begin, int64_t col_end, int64_t * col_min,
                             int64_t * col_max, int64_t * lo_end, int64_t * hi_begin)
try {
    SPMV_TRY(check(m));
    if (m->format != SPMVB200_CSR) return fail(SPMVB200_ERR_INVALID, "not a CSR matrix");
    SPMV_TRY(csr_ensure_row_major(m));
    Scratch<long long> d;
    SPMV_TRY(d.alloc(4));
    long long h[4] = {LLONG_MAX, -1, 0, (long long)m->rows};
    SPMV_CUDA(cudaMemcpyAsync(d.p, h, sizeof h, cudaMemcpyHostToDevice, m->stream));
    if (m->rows > 0) {
        if (m->off64) column_span_kernel<int64_t><<<grid_for(m->rows), 256, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, m->col, col_begin, col_end, d.p);
        else column_span_kernel<uint32_t><<<grid_for(m->rows), 256, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, m->col, col_begin, col_end, d.p);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_CUDA(cudaMemcpyAsync(h, d.p, sizeof h, cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    if (col_min) *col_min = h[1] < 0 ? -1 : h[0];
    if (col_max) *col_max = h[1];
    if (lo_end) *lo_end = h[2];
    if (hi_begin) *hi_begin = h[3];
    return 0;
}
SPMV_ABI_CATCH

int spmvb200_csr_row_block(spmvb200_matrix_t src, int64_t row_begin, int64_t row_end, spmvb200_matrix_t * out)
try {
    SPMV_TRY(check(src));
    if (!out || src->format != SPMVB200_CSR || row_begin < 0 || row_end > src->rows || row_begin > row_end)
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    SPMV_TRY(csr_ensure_row_major(src));
    cudaStream_t s0 = src->stream;
    SPMV_CUDA(cudaStreamSynchronize(s0));
    int64_t first = 0, last = 0;
    const size_t osz = src->off64 ? 8 : 4;
    uint64_t tmp[2] = {0, 0};
    SPMV_CUDA(cudaMemcpy(&tmp[0], (const char *)src->rp + osz * (size_t)row_begin, osz, cudaMemcpyDeviceToHost));
    SPMV_CUDA(cudaMemcpy(&tmp[1], (const char *)src->rp + osz * (size_t)row_end, osz, cudaMemcpyDeviceToHost));
    first = src->off64 ? (int64_t)tmp[0] : (int64_t)(uint32_t)tmp[0];
    last = src->off64 ? (int64_t)tmp[1] : (int64_t)(uint32_t)tmp[1];
    const int64_t rows = row_end - row_begin, stored = last - first;
    Matrix * m = nullptr;
    SPMV_TRY(matrix_new(&m));
    Guard g(m);
    cudaStream_t s = m->stream;
    m->format = SPMVB200_CSR;
    m->rows = rows; m->cols = src->cols; m->nnz = stored; m->stored = stored;
    m->row_alignment = src->row_alignment;
    m->row_offset = src->row_offset + row_begin;
    SPMV_TRY(alloc_streamed(m, &m->col, stored));
    SPMV_TRY(alloc_streamed(m, &m->val, stored));
    if (stored > 0) {
        SPMV_CUDA(cudaMemcpyAsync(m->col, src->col + first, sizeof(int32_t) * (size_t)stored, cudaMemcpyDeviceToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(m->val, src->val + first, sizeof(double) * (size_t)stored, cudaMemcpyDeviceToDevice, s));
    }
    Scratch<int64_t> rp64;
    SPMV_TRY(rp64.alloc(rows + 1));
    if (src->off64) rebase_offsets_kernel<int64_t><<<grid_for(rows + 1), 256, 0, s>>>(rows + 1, (const int64_t *)src->rp + row_begin, first, rp64.p);
    else rebase_offsets_kernel<uint32_t><<<grid_for(rows + 1), 256, 0, s>>>(rows + 1, (const uint32_t *)src->rp + row_begin, first, rp64.p);
    SPMV_CUDA(cudaGetLastError());
    SPMV_TRY(store_offsets(m, rp64.p, rows, stored));
    SPMV_TRY(csr_build_tiles(m));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return finish(g, out);
}
SPMV_ABI_CATCH

// ---- cache model on a device matrix -------------------------------------------------------------------------------

int spmvb200_cache_trace(spmvb200_matrix_t m, const spmvb200_cache_config * cfg, spmvb200_cache_misses * out)
try {
    SPMV_TRY(check(m));
    if (!cfg || !out || cfg->parts < 1) return fail(SPMVB200_ERR_INVALID, "bad argument");
    if (m->format == SPMVB200_CSR) {
        std::vector<int64_t> rp((size_t)m->rows + 1);
        std::vector<int32_t> col((size_t)std::max<int64_t>(m->stored, 1));
        SPMV_TRY(spmvb200_csr_export(m, rp.data(), col.data(), nullptr));
        return spmvb200_cache_trace_csr(m->rows, m->cols, rp.data(), col.data(), cfg, out);
    }
    if (m->format == SPMVB200_ELL) {
        std::vector<int32_t> col((size_t)std::max<int64_t>(m->rows * m->ell_w, 1));
        SPMV_TRY(spmvb200_ell_export(m, col.data(), nullptr));
        return spmvb200_cache_trace_ell(m->rows, m->cols, m->ell_w, col.data(), cfg, out);
    }
    if (m->format == SPMVB200_COO) {
        std::vector<int32_t> row((size_t)std::max<int64_t>(m->coo_n, 1)), col((size_t)std::max<int64_t>(m->coo_n, 1));
        SPMV_TRY(export_coo_arrays(m, row.data(), col.data(), nullptr, true));  // the order the kernel walks
        return spmvb200_cache_trace_coo(m->rows, m->cols, m->coo_n, row.data(), col.data(), cfg, out);
    }
    if (m->format == SPMVB200_HYB) {
        std::vector<int32_t> ecol((size_t)std::max<int64_t>(m->rows * m->ell_w, 1));
        std::vector<int32_t> row((size_t)std::max<int64_t>(m->coo_n, 1)), col((size_t)std::max<int64_t>(m->coo_n, 1));
        SPMV_TRY(export_ell_arrays(m, ecol.data(), nullptr));
        SPMV_TRY(export_coo_arrays(m, row.data(), col.data(), nullptr, true));  // the order the kernel walks
        std::vector<spmvb200_cache_misses> tail((size_t)cfg->parts);
        spmvb200_cache_config c2 = *cfg;
        c2.starts = nullptr;  // the tail is cut by entries (hybrid-matrix.cpp:491-528)
        SPMV_TRY(spmvb200_cache_trace_ell(m->rows, m->cols, m->ell_w, ecol.data(), cfg, out));
        SPMV_TRY(spmvb200_cache_trace_coo(m->rows, m->cols, m->coo_n, row.data(), col.data(), &c2, tail.data()));
        for (int p = 0; p < cfg->parts; p++) {
            int64_t * a = reinterpret_cast<int64_t *>(&out[p]);
            const int64_t * b = reinterpret_cast<const int64_t *>(&tail[(size_t)p]);
            for (size_t k = 0; k < sizeof(spmvb200_cache_misses) / sizeof(int64_t); k++) a[k] += b[k];
        }
        return 0;
    }
    return fail(SPMVB200_ERR_INVALID, "unknown format");
}
SPMV_ABI_CATCH

// ---- column split (interior / boundary for matrices that are not banded) ------------------------------------------

int spmvb200_csr_column_split(spmvb200_matrix_t src, int64_t col_begin, int64_t col_end, spmvb200_matrix_t * inside,
                              spmvb200_matrix_t * outside)
try {
    SPMV_TRY(check(src));
    if (!inside || !outside || src->format != SPMVB200_CSR || col_begin < 0 || col_begin > col_end)
        return fail(SPMVB200_ERR_INVALID, "bad argument");
    SPMV_TRY(csr_ensure_row_major(src));
    cudaStream_t s = src->stream;
    const int64_t rows = src->rows;
    Scratch<int64_t> n_in, n_out, rp_in, rp_out;
    Scratch<unsigned char> tmp;
    SPMV_TRY(n_in.alloc(rows + 1)); SPMV_TRY(n_out.alloc(rows + 1)); SPMV_TRY(rp_in.alloc(rows + 1)); SPMV_TRY(rp_out.alloc(rows + 1));
    if (src->off64) column_count_kernel<int64_t><<<grid_for(rows + 1), 256, 0, s>>>(rows, (const int64_t *)src->rp, src->col, col_begin, col_end, n_in.p, n_out.p);
    else column_count_kernel<uint32_t><<<grid_for(rows + 1), 256, 0, s>>>(rows, (const uint32_t *)src->rp, src->col, col_begin, col_end, n_in.p, n_out.p);
    SPMV_CUDA(cudaGetLastError());
    size_t tb = 0;
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, n_in.p, rp_in.p, rows + 1, s));
    SPMV_TRY(tmp.alloc((int64_t)tb));
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, n_in.p, rp_in.p, rows + 1, s));
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, n_out.p, rp_out.p, rows + 1, s));
    int64_t tot_in = 0, tot_out = 0;
    SPMV_CUDA(cudaMemcpyAsync(&tot_in, rp_in.p + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaMemcpyAsync(&tot_out, rp_out.p + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    Matrix * parts[2] = {nullptr, nullptr};
    SPMV_TRY(matrix_new(&parts[0]));
    Guard g0(parts[0]);
    SPMV_TRY(matrix_new(&parts[1]));
    Guard g1(parts[1]);
    const int64_t totals[2] = {tot_in, tot_out};
    for (int k = 0; k < 2; k++) {
        Matrix * m = parts[k];
        m->format = SPMVB200_CSR;
        m->rows = rows; m->cols = src->cols; m->nnz = totals[k]; m->stored = totals[k];
        m->row_alignment = 1;
        m->row_offset = src->row_offset;
        SPMV_TRY(alloc_streamed(m, &m->col, totals[k]));
        SPMV_TRY(alloc_streamed(m, &m->val, totals[k]));
        SPMV_CUDA(cudaStreamSynchronize(m->stream));  // the padding memsets ran on the new matrix's own stream
    }
    if (rows > 0) {
        if (src->off64) column_fill_kernel<int64_t><<<grid_for(rows), 256, 0, s>>>(rows, (const int64_t *)src->rp, src->col, src->val, col_begin, col_end, rp_in.p, rp_out.p, parts[0]->col, parts[0]->val, parts[1]->col, parts[1]->val);
        else column_fill_kernel<uint32_t><<<grid_for(rows), 256, 0, s>>>(rows, (const uint32_t *)src->rp, src->col, src->val, col_begin, col_end, rp_in.p, rp_out.p, parts[0]->col, parts[0]->val, parts[1]->col, parts[1]->val);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_CUDA(cudaStreamSynchronize(s));
    SPMV_TRY(store_offsets(parts[0], rp_in.p, rows, tot_in));
    SPMV_TRY(store_offsets(parts[1], rp_out.p, rows, tot_out));
    SPMV_TRY(csr_build_tiles(parts[0]));
    SPMV_TRY(csr_build_tiles(parts[1]));
    SPMV_TRY(finish(g0, inside));
    const int rc = finish(g1, outside);
    if (rc) {
        matrix_free(*inside);
        *inside = nullptr;
    }
    return rc;
}
SPMV_ABI_CATCH

}  // extern "C"
