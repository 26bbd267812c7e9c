// builders.cu -- device-resident matrix builders.
//
// They replace the reference's host converters
//   csr_matrix::from_matrix_market_row_aligned   matrix/csr-matrix.cpp:193-243
//   coo_matrix::from_matrix_market               matrix/coo-matrix.cpp:220-243
//   ell_matrix::from_matrix_market               matrix/ell-matrix.cpp:190-238
//   hybrid_matrix::from_matrix_market            matrix/hybrid-matrix.cpp:316-417
// and must produce bit-identical arrays (checked by tests/test_gpu_parity.py through the export
// functions).  Everything runs on the GPU: the row-major sort is a stable 64-bit LSD radix sort of
// (row << 32 | column) keys, row pointers are binary searches in the sorted keys, padding and the
// ELL/COO split are closed-form per row.  Sorting and prefix sums use CUB (shipped with the CUDA
// toolkit); they are one-time set-up work, not part of the timed SpMV path.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <climits>
#include <vector>

namespace spmvb200 {

// ---------------------------------------------------------------------------------------------
// handle life cycle
// ---------------------------------------------------------------------------------------------

__global__ void fill_kernel(double * p, int64_t n, double v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

int fill_device(double * p, int64_t n, double v, cudaStream_t s)
{
    if (n <= 0) return 0;
    int64_t grid = std::min<int64_t>((n + 255) / 256, 148 * 16);
    fill_kernel<<<(unsigned)grid, 256, 0, s>>>(p, n, v);
    SPMV_CUDA(cudaGetLastError());
    return 0;
}

int matrix_new(Matrix ** out)
{
    *out = nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice (is a CUDA device present?)", __FILE__, __LINE__);
    Matrix * m = new Matrix();
    m->device = dev;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaDeviceGetAttribute", __FILE__, __LINE__); }
    m->sm_count = sms > 0 ? sms : 148;
    e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete m; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    m->own_stream = true;
    stream_register(m->stream);
    cudaEventCreate(&m->ev0);
    cudaEventCreate(&m->ev1);
    *out = m;
    return 0;
}

int matrix_alloc_vectors(Matrix * m)
{
    // x = 1.0, y = 0.0 as Kernel::init does (reference kernels/csr-spmv.cpp:35-36).
    SPMV_TRY(dev_alloc(m, &m->x, m->cols + 8));
    SPMV_TRY(dev_alloc(m, &m->y, m->rows + 8));
    m->own_x = m->own_y = true;
    SPMV_TRY(fill_device(m->x, m->cols + 8, 1.0, m->stream));
    SPMV_CUDA(cudaMemsetAsync(m->y, 0, sizeof(double) * (size_t)(m->rows + 8), m->stream));
    return 0;
}

void matrix_free(Matrix * m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    cudaFree(m->rp); cudaFree(m->col); cudaFree(m->val); cudaFree(m->tile_row); cudaFree(m->span_row); cudaFree(m->flat_meta); cudaFree(m->flat_rowmap); cudaFree(m->slice_col); cudaFree(m->slice_val); cudaFree(m->slice_flags); cudaFree(m->slice_cofs); cudaFree(m->slice_meta);
    cudaFree(m->ell_col); cudaFree(m->ell_val);
    cudaFree(m->coo_row); cudaFree(m->coo_col); cudaFree(m->coo_val);
    cudaFree(m->coo_colh); cudaFree(m->coo_hot_cols); cudaFree(m->coo_seg);
    if (m->x_tex) cudaDestroyTextureObject(m->x_tex);
    if (m->own_x) cudaFree(m->x);
    if (m->own_y) cudaFree(m->y);
    if (m->upload_stream) cudaStreamDestroy(m->upload_stream);
    if (m->ev_x) cudaEventDestroy(m->ev_x);
    for (auto & e : m->ev_chunk) if (e) cudaEventDestroy(e);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->own_stream && m->stream) {
        stream_forget(m->stream);
        cudaStreamDestroy(m->stream);
    }
    delete m;
}

static int bits_for(int64_t n)  // number of bits needed to represent values in [0, n)
{
    int b = 1;
    while (b < 63 && ((int64_t)1 << b) < n) ++b;
    return b;
}

// ---------------------------------------------------------------------------------------------
// CSR from unsorted 1-based entries
// ---------------------------------------------------------------------------------------------

__global__ void make_keys_kernel(int64_t n, const int32_t * i, const int32_t * j, int32_t rows, int32_t cols,
                                 uint64_t * keys, uint32_t * idx, int * bad)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t r = i[k] - 1, c = j[k] - 1;
        if (r < 0 || r >= rows || c < 0 || c >= cols) *bad = 1;
        keys[k] = ((uint64_t)(uint32_t)r << 32) | (uint32_t)c;
        idx[k] = (uint32_t)k;
    }
}

// rp_raw[r] = number of sorted keys with row < r  (r in [0, rows])
__global__ void row_lower_bound_kernel(int64_t rows, int64_t n, const uint64_t * keys, int64_t * rp_raw)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t target = (uint64_t)r << 32;
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        rp_raw[r] = lo;
    }
}

__global__ void aligned_len_kernel(int64_t rows, const int64_t * rp_raw, int64_t align, int64_t * len_al)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x)
        len_al[r] = r < rows ? (rp_raw[r + 1] - rp_raw[r] + align - 1) / align * align : 0;
}

template <typename OffT>
__global__ void narrow_offsets_kernel(int64_t n, const int64_t * in, OffT * out)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        out[r] = (OffT)in[r];
}

// Sorted entry k (row r) goes to rp[r] + (k - rp_raw[r]).
__global__ void scatter_sorted_kernel(int64_t n, const uint64_t * keys, const uint32_t * idx, const double * a,
                                      const int64_t * rp_raw, const int64_t * rp, int32_t * col, double * val)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[k];
        const int64_t r = (int64_t)(key >> 32);
        const int64_t d = rp[r] + (k - rp_raw[r]);
        col[d] = (int32_t)(uint32_t)key;
        val[d] = a[idx[k]];
    }
}

extern std::atomic<int> g_force_off64;

int store_offsets(Matrix * m, const int64_t * d_rp64, int64_t rows, int64_t stored)
{
    m->off64 = stored >= ((int64_t)1 << 32) || g_force_off64.load() != 0;
    if (m->off64) {
        int64_t * rp = nullptr;
        SPMV_TRY(dev_alloc(m, &rp, rows + 1 + 2048));  // slack: the CSR kernel bulk-copies fixed-size row-pointer windows
        SPMV_CUDA(cudaMemcpyAsync(rp, d_rp64, sizeof(int64_t) * (size_t)(rows + 1), cudaMemcpyDeviceToDevice, m->stream));
        m->rp = rp;
    } else {
        uint32_t * rp = nullptr;
        SPMV_TRY(dev_alloc(m, &rp, rows + 1 + 2048));  // slack: the CSR kernel bulk-copies fixed-size row-pointer windows
        narrow_offsets_kernel<uint32_t><<<grid_for(rows + 1), 256, 0, m->stream>>>(rows + 1, d_rp64, rp);
        SPMV_CUDA(cudaGetLastError());
        m->rp = rp;
    }
    return 0;
}

int csr_from_entries_host(int64_t rows, int64_t cols, int64_t n, const int32_t * hi, const int32_t * hj,
                          const double * ha, int32_t row_alignment, Matrix * m)
{
    if (row_alignment < 1) return fail(SPMVB200_ERR_INVALID, "row_alignment must be >= 1");
    cudaStream_t s = m->stream;
    Scratch<int32_t> di, dj;
    Scratch<double> da;
    Scratch<uint64_t> keys, keys2;
    Scratch<uint32_t> idx, idx2;
    Scratch<int64_t> rp_raw, len_al, rp_al;
    Scratch<int> bad;
    Scratch<unsigned char> tmp;
    SPMV_TRY(di.alloc(n)); SPMV_TRY(dj.alloc(n)); SPMV_TRY(da.alloc(n));
    SPMV_TRY(keys.alloc(n)); SPMV_TRY(keys2.alloc(n)); SPMV_TRY(idx.alloc(n)); SPMV_TRY(idx2.alloc(n));
    SPMV_TRY(rp_raw.alloc(rows + 1)); SPMV_TRY(bad.alloc(1));
    SPMV_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    if (n > 0) {
        SPMV_CUDA(cudaMemcpyAsync(di.p, hi, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(dj.p, hj, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(da.p, ha, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
        make_keys_kernel<<<grid_for(n), 256, 0, s>>>(n, di.p, dj.p, (int32_t)rows, (int32_t)cols, keys.p, idx.p, bad.p);
        SPMV_CUDA(cudaGetLastError());
        int hbad = 0;
        SPMV_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        if (hbad) return fail(SPMVB200_ERR_INVALID, "entry index outside the matrix");
        // stable LSD radix sort by (row, column): sort_matrix_row_major (matrix-market.cpp:897-929)
        size_t tmp_bytes = 0;
        const int end_bit = 32 + bits_for(rows);
        SPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, idx.p, idx2.p, n, 0, end_bit, s));
        SPMV_TRY(tmp.alloc((int64_t)tmp_bytes));
        SPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, idx.p, idx2.p, n, 0, end_bit, s));
    }
    row_lower_bound_kernel<<<grid_for(rows + 1), 256, 0, s>>>(rows, n, keys2.p, rp_raw.p);
    SPMV_CUDA(cudaGetLastError());

    const int64_t * rp64 = rp_raw.p;
    int64_t stored = n;
    if (row_alignment > 1) {
        // padded row lengths and their prefix sum (csr-matrix.cpp:206-217)
        SPMV_TRY(len_al.alloc(rows + 1)); SPMV_TRY(rp_al.alloc(rows + 1));
        aligned_len_kernel<<<grid_for(rows + 1), 256, 0, s>>>(rows, rp_raw.p, row_alignment, len_al.p);
        SPMV_CUDA(cudaGetLastError());
        size_t tb = 0;
        Scratch<unsigned char> t2;
        SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len_al.p, rp_al.p, rows + 1, s));
        SPMV_TRY(t2.alloc((int64_t)tb));
        SPMV_CUDA(cub::DeviceScan::ExclusiveSum(t2.p, tb, len_al.p, rp_al.p, rows + 1, s));
        SPMV_CUDA(cudaMemcpyAsync(&stored, rp_al.p + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        rp64 = rp_al.p;
    }
    if (stored > INT32_MAX) return fail(SPMVB200_ERR_OVERFLOW, "CSR: number of stored entries exceeds int32");

    m->format = SPMVB200_CSR;
    m->rows = rows; m->cols = cols; m->nnz = n; m->stored = stored; m->row_alignment = row_alignment;
    SPMV_TRY(alloc_streamed(m, &m->col, stored));
    SPMV_TRY(alloc_streamed(m, &m->val, stored));
    if (row_alignment > 1) {  // padding = (column 0, 0.0) (csr-matrix.cpp:232-236)
        SPMV_CUDA(cudaMemsetAsync(m->col, 0, sizeof(int32_t) * (size_t)stored, s));
        SPMV_CUDA(cudaMemsetAsync(m->val, 0, sizeof(double) * (size_t)stored, s));
    }
    if (n > 0) {
        scatter_sorted_kernel<<<grid_for(n), 256, 0, s>>>(n, keys2.p, idx2.p, da.p, rp_raw.p, rp64, m->col, m->val);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_TRY(store_offsets(m, rp64, rows, stored));
    SPMV_TRY(csr_build_tiles(m));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// Adopt device arrays that were allocated with alloc-compatible slack by the caller (generators).
int csr_adopt(Matrix * m, int64_t rows, int64_t cols, int64_t nnz, int64_t stored, bool off64, void * rp,
              int32_t * col, double * val)
{
    m->format = SPMVB200_CSR;
    m->rows = rows; m->cols = cols; m->nnz = nnz; m->stored = stored; m->off64 = off64;
    m->rp = rp; m->col = col; m->val = val;
    return csr_build_tiles(m);
}

// ---------------------------------------------------------------------------------------------
// ELL / HYB / COO from a device CSR matrix (row_alignment 1)
// ---------------------------------------------------------------------------------------------

template <typename OffT>
__global__ void max_len_kernel(int64_t rows, const OffT * rp, int * out)
{
    int best = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        int len = (int)(rp[r + 1] - rp[r]);
        best = len > best ? len : best;
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, best);
}

template <typename OffT>
__global__ void len_hist_kernel(int64_t rows, const OffT * rp, unsigned long long * hist)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[rp[r + 1] - rp[r]], 1ull);
}

// ELL part with the reference's padding rule.  Row r, slot l < W:
//   l < len(r): the l-th entry of the row;
//   else: value 0.0, column = INT32_MAX (skip_padding) or the column of the last entry consumed so
//   far by the reference's sequential fill, which is entry rp[r+1]-1 (ell-matrix.cpp:226-232,
//   hybrid-matrix.cpp:387-392); 0 when nothing was consumed yet (the hybrid converter's guard; the
//   ELL converter reads out of bounds there).
// Rows with len >= W keep their first W entries (hybrid-matrix.cpp:394-400).
template <typename OffT>
__global__ void ell_fill_kernel(int64_t rows, int64_t W, int64_t pitch, int skip, const OffT * rp,
                                const int32_t * col, const double * val, int32_t * ecol, double * eval)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = (int64_t)rp[r], e = (int64_t)rp[r + 1];
        const int64_t len = e - b;
        const int32_t padc = skip ? INT_MAX : (e > 0 ? col[e - 1] : 0);
        for (int64_t l = 0; l < W; ++l) {
            const bool real = l < len;
            ecol[l * pitch + r] = real ? col[b + l] : padc;
            eval[l * pitch + r] = real ? val[b + l] : 0.0;
        }
    }
}

template <typename OffT>
__global__ void excess_len_kernel(int64_t rows, int64_t W, const OffT * rp, int64_t * excess)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t len = r < rows ? (int64_t)(rp[r + 1] - rp[r]) : 0;
        excess[r] = len > W ? len - W : 0;
    }
}

// COO tail of the hybrid format: entries W.. of every row longer than W, row-major (hybrid-matrix.cpp:402-408).
template <typename OffT>
__global__ void hyb_tail_kernel(int64_t rows, int64_t W, const OffT * rp, const int32_t * col, const double * val,
                                const int64_t * off, int32_t * crow, int32_t * ccol, double * cval)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = (int64_t)rp[r], e = (int64_t)rp[r + 1];
        int64_t d = off[r];
        for (int64_t k = b + W; k < e; ++k, ++d) {
            crow[d] = (int32_t)r;
            ccol[d] = col[k];
            cval[d] = val[k];
        }
    }
}

template <typename OffT>
__global__ void expand_rows_kernel(int64_t rows, const OffT * rp, int32_t * crow)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        for (int64_t k = (int64_t)rp[r]; k < (int64_t)rp[r + 1]; ++k) crow[k] = (int32_t)r;
}

static void copy_shape(const Matrix * src, Matrix * dst)
{
    dst->rows = src->rows; dst->cols = src->cols; dst->nnz = src->nnz; dst->row_offset = src->row_offset;
}

template <typename OffT>
static int max_row_length(const Matrix * src, cudaStream_t s, int64_t * out)
{
    Scratch<int> d;
    SPMV_TRY(d.alloc(1));
    SPMV_CUDA(cudaMemsetAsync(d.p, 0, sizeof(int), s));
    if (src->rows > 0) {
        max_len_kernel<OffT><<<grid_for(src->rows), 256, 0, s>>>(src->rows, (const OffT *)src->rp, d.p);
        SPMV_CUDA(cudaGetLastError());
    }
    int h = 0;
    SPMV_CUDA(cudaMemcpyAsync(&h, d.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    *out = h;
    return 0;
}

// Longest row of a CSR matrix (cached in the handle; used to pick the SpMV kernel).
int csr_max_row_length(Matrix * m)
{
    if (m->csr_maxlen >= 0) return 0;
    int64_t v = 0;
    SPMV_TRY(m->off64 ? max_row_length<int64_t>(m, m->stream, &v) : max_row_length<uint32_t>(m, m->stream, &v));
    m->csr_maxlen = v;
    return 0;
}

template <typename OffT>
static int ell_part(const Matrix * src, int64_t W, int skip, Matrix * dst)
{
    cudaStream_t s = dst->stream;
    dst->ell_w = W;
    dst->ell_pitch = round_up(std::max<int64_t>(src->rows, 1), 32);
    dst->skip_padding = skip;
    const int64_t slots = dst->ell_pitch * W;
    SPMV_TRY(alloc_streamed(dst, &dst->ell_col, slots));
    SPMV_TRY(alloc_streamed(dst, &dst->ell_val, slots));
    if (slots > 0) {
        SPMV_CUDA(cudaMemsetAsync(dst->ell_col, 0, sizeof(int32_t) * (size_t)slots, s));
        SPMV_CUDA(cudaMemsetAsync(dst->ell_val, 0, sizeof(double) * (size_t)slots, s));
        ell_fill_kernel<OffT><<<grid_for(src->rows), 256, 0, s>>>(src->rows, W, dst->ell_pitch, skip,
                                                                  (const OffT *)src->rp, src->col, src->val,
                                                                  dst->ell_col, dst->ell_val);
        SPMV_CUDA(cudaGetLastError());
    }
    return 0;
}

template <typename OffT>
static int ell_from_csr_t(const Matrix * src, int skip, bool check_int32, Matrix * dst)
{
    int64_t W = 0;
    SPMV_TRY(max_row_length<OffT>(src, dst->stream, &W));  // ell-matrix.cpp:199
    if (check_int32 && src->rows * W > INT32_MAX)          // ell-matrix.cpp:200-205
        return fail(SPMVB200_ERR_OVERFLOW,
                    "Failed to convert to ELLPACK: Integer overflow when computing number of non-zeros");
    copy_shape(src, dst);
    dst->format = SPMVB200_ELL;
    dst->stored = src->rows * W;
    SPMV_TRY(ell_part<OffT>(src, W, skip, dst));
    SPMV_CUDA(cudaStreamSynchronize(dst->stream));
    return 0;
}

int ell_from_csr(const Matrix * src, int skip, bool check_int32, Matrix * dst)
{
    return src->off64 ? ell_from_csr_t<int64_t>(src, skip, check_int32, dst)
                      : ell_from_csr_t<uint32_t>(src, skip, check_int32, dst);
}

template <typename OffT>
static int hyb_from_csr_t(const Matrix * src, int skip, bool check_int32, Matrix * dst)
{
    cudaStream_t s = dst->stream;
    int64_t maxlen = 0;
    SPMV_TRY(max_row_length<OffT>(src, s, &maxlen));
    // histogram of row lengths on the device, 2/3-quantile rule on the host (hybrid-matrix.cpp:329-344)
    Scratch<unsigned long long> dh;
    SPMV_TRY(dh.alloc(maxlen + 1));
    SPMV_CUDA(cudaMemsetAsync(dh.p, 0, sizeof(unsigned long long) * (size_t)(maxlen + 1), s));
    if (src->rows > 0) {
        len_hist_kernel<OffT><<<grid_for(src->rows), 256, 0, s>>>(src->rows, (const OffT *)src->rp, dh.p);
        SPMV_CUDA(cudaGetLastError());
    }
    std::vector<unsigned long long> hist((size_t)maxlen + 1);
    SPMV_CUDA(cudaMemcpyAsync(hist.data(), dh.p, sizeof(unsigned long long) * hist.size(), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    int64_t W = 0, below = 0;
    const int64_t target = (2 * src->rows) / 3;
    while (below < target) {
        below += (int64_t)hist[(size_t)W];
        W++;
    }
    W = W == 0 ? 0 : W - 1;
    if (check_int32 && src->rows * W > INT32_MAX)  // hybrid-matrix.cpp:349-354
        return fail(SPMVB200_ERR_OVERFLOW,
                    "Failed to convert to HYBRID: Integer overflow when computing number of non-zeros");
    copy_shape(src, dst);
    dst->format = SPMVB200_HYB;
    dst->n_ell = src->rows * W;
    SPMV_TRY(ell_part<OffT>(src, W, skip, dst));

    Scratch<int64_t> excess, off;
    Scratch<unsigned char> tmp;
    SPMV_TRY(excess.alloc(src->rows + 1)); SPMV_TRY(off.alloc(src->rows + 1));
    excess_len_kernel<OffT><<<grid_for(src->rows + 1), 256, 0, s>>>(src->rows, W, (const OffT *)src->rp, excess.p);
    SPMV_CUDA(cudaGetLastError());
    size_t tb = 0;
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, excess.p, off.p, src->rows + 1, s));
    SPMV_TRY(tmp.alloc((int64_t)tb));
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, excess.p, off.p, src->rows + 1, s));
    int64_t ncoo = 0;
    SPMV_CUDA(cudaMemcpyAsync(&ncoo, off.p + src->rows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    dst->n_coo = ncoo;
    dst->coo_n = ncoo;
    dst->coo_sorted = true;
    dst->coo_mode = SPMVB200_COO_SEGMENTED;
    dst->stored = dst->n_ell + ncoo;
    SPMV_TRY(alloc_streamed(dst, &dst->coo_row, ncoo));
    SPMV_TRY(alloc_streamed(dst, &dst->coo_col, ncoo));
    SPMV_TRY(alloc_streamed(dst, &dst->coo_val, ncoo));
    if (ncoo > 0) {
        hyb_tail_kernel<OffT><<<grid_for(src->rows), 256, 0, s>>>(src->rows, W, (const OffT *)src->rp, src->col,
                                                                  src->val, off.p, dst->coo_row, dst->coo_col,
                                                                  dst->coo_val);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_CUDA(cudaStreamSynchronize(s));
    return coo_column_blocks(dst);
}

int hyb_from_csr(const Matrix * src, int skip, bool check_int32, Matrix * dst)
{
    return src->off64 ? hyb_from_csr_t<int64_t>(src, skip, check_int32, dst)
                      : hyb_from_csr_t<uint32_t>(src, skip, check_int32, dst);
}

int coo_from_csr(const Matrix * src, int mode, Matrix * dst)
{
    cudaStream_t s = dst->stream;
    copy_shape(src, dst);
    const int64_t n = src->stored;
    int32_t *row = nullptr, *col = nullptr;
    double * val = nullptr;
    SPMV_TRY(alloc_streamed(dst, &row, n));
    SPMV_TRY(alloc_streamed(dst, &col, n));
    SPMV_TRY(alloc_streamed(dst, &val, n));
    if (n > 0) {
        if (src->off64) expand_rows_kernel<int64_t><<<grid_for(src->rows), 256, 0, s>>>(src->rows, (const int64_t *)src->rp, row);
        else expand_rows_kernel<uint32_t><<<grid_for(src->rows), 256, 0, s>>>(src->rows, (const uint32_t *)src->rp, row);
        SPMV_CUDA(cudaGetLastError());
        SPMV_CUDA(cudaMemcpyAsync(col, src->col, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        SPMV_CUDA(cudaMemcpyAsync(val, src->val, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    }
    return coo_adopt(dst, src->rows, src->cols, n, row, col, val, mode, true);
}

// ---------------------------------------------------------------------------------------------
// COO
// ---------------------------------------------------------------------------------------------

__global__ void is_sorted_kernel(int64_t n, const int32_t * row, int * unsorted)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; k < n; k += (int64_t)gridDim.x * blockDim.x)
        if (row[k] < row[k - 1]) *unsorted = 1;
}

__global__ void iota_kernel(int64_t n, uint32_t * idx)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        idx[k] = (uint32_t)k;
}

__global__ void gather_kernel(int64_t n, const uint32_t * idx, const int32_t * col, const double * val,
                              int32_t * col2, double * val2)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        col2[k] = col[idx[k]];
        val2[k] = val[idx[k]];
    }
}

// ---- column-blocked order for matrices whose x does not fit in L2 -----------------------------------
// With x several times larger than L2 the gathers of a row-sorted sweep miss (R-MAT 2^26: the COO kernel
// read 64 GB for 33 GB of entries, profiles/r01_ncu_c4_hyb_warp4.txt).  The order of COO entries is
// free, so the row-sorted entries are stably partitioned by column block (blocks of 2^shift columns,
// at most L2/2 bytes of x each): inside a block the rows still ascend, so runs of equal rows stay
// adjacent for the segmented reduce, the block's slice of x stays in L2 for the whole pass, and the
// price is one sweep over y per block; the reorder is done only when that is a win (R-MAT 2^26 hybrid:
// 12.9 -> 9.0 ms with 16 blocks of 2^22 columns, profiles/r01_sweep_k_coo_column_blocks.log).  Only applied when the row-sorted order has
// ascending columns inside each row, so that a stable sort by row restores it exactly for export.
extern std::atomic<int64_t> g_coo_col_block_log2;  // global option "coo.col_block_log2": -1 never, 0 auto, k block of 2^k columns

__global__ void rowcol_sorted_kernel(int64_t n, const int32_t * row, const int32_t * col, int * unsorted)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; k < n; k += (int64_t)gridDim.x * blockDim.x)
        if (row[k] < row[k - 1] || (row[k] == row[k - 1] && col[k] < col[k - 1])) *unsorted = 1;
}

__global__ void block_hist_kernel(int64_t n, const int32_t * col, int shift, unsigned long long * hist /* 256 */)
{
    __shared__ unsigned int local[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) local[t] = 0;
    __syncthreads();
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&local[(col[k] >> shift) & 255], 1u);
    __syncthreads();
    for (int t = threadIdx.x; t < 256; t += blockDim.x)
        if (local[t]) atomicAdd(&hist[t], (unsigned long long)local[t]);
}

// For every chunk of 2^20 consecutive entries: which column blocks does it touch (256-bit mask)?
__global__ void block_spread_kernel(int64_t n, const int32_t * col, int shift, unsigned int * masks /* [chunks][8] */)
{
    __shared__ unsigned int local[8];
    const int64_t chunk = blockIdx.y;
    if (threadIdx.x < 8) local[threadIdx.x] = 0;
    __syncthreads();
    const int64_t b = chunk << 20, e = min(n, b + ((int64_t)1 << 20));
    unsigned int mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = b + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < e; k += (int64_t)gridDim.x * blockDim.x) {
        const int blk = (col[k] >> shift) & 255;
        mine[blk >> 5] |= 1u << (blk & 31);
    }
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (mine[w]) atomicOr(&local[w], mine[w]);
    __syncthreads();
    if (threadIdx.x < 8 && local[threadIdx.x]) atomicOr(&masks[chunk * 8 + threadIdx.x], local[threadIdx.x]);
}

__global__ void block_key_kernel(int64_t n, const int32_t * col, int shift, unsigned char * key, uint32_t * idx)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        key[k] = (unsigned char)(col[k] >> shift);
        idx[k] = (uint32_t)k;
    }
}

__global__ void gather3_kernel(int64_t n, const uint32_t * idx, const int32_t * row, const int32_t * col, const double * val,
                               int32_t * row2, int32_t * col2, double * val2)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = idx[k];
        row2[k] = row[s];
        col2[k] = col[s];
        val2[k] = val[s];
    }
}

int coo_column_blocks(Matrix * m)
{
    const int64_t opt = g_coo_col_block_log2.load();
    const int64_t n = m->coo_n;
    m->coo_col_shift = 0;
    if (opt < 0 || n < 2 || n >= ((int64_t)1 << 32) || !m->coo_sorted) return 0;
    int l2 = 0;
    SPMV_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, m->device));
    int shift = (int)opt;
    if (opt == 0) {
        if (l2 <= 0 || 8 * m->cols <= (int64_t)l2 / 4 * 3) return 0;  // x fits in L2 with room to spare: nothing to gain
        shift = 0;
        while (((int64_t)16 << shift) <= l2 / 2) shift++;  // largest block with 8 * 2^shift <= L2/2
    }
    if (opt == 0)
        while (((m->cols + ((int64_t)1 << shift) - 1) >> shift) > 256) shift++;  // one-byte block keys: at most 256 blocks
    const int64_t nblocks = (m->cols + ((int64_t)1 << shift) - 1) >> shift;
    if (nblocks < 2 || nblocks > 256) return 0;
    cudaStream_t s = m->stream;
    if (opt == 0) {
        // automatic: look at the column blocks the entries really fall into (a piece of a column-split matrix
        // references a part of x only).  Worth it when the x referenced does not fit in L2 and the gather
        // misses saved (about 12 B per entry measured on R-MAT 2^26) outweigh the extra sweeps over y (a 32 B
        // sector read + written per 4 rows, about half the rows touched per block, never more than its entries).
        Scratch<unsigned long long> dh;
        SPMV_TRY(dh.alloc(256));
        SPMV_CUDA(cudaMemsetAsync(dh.p, 0, 256 * sizeof(unsigned long long), s));
        block_hist_kernel<<<grid_for(n, m->sm_count), 256, 0, s>>>(n, m->coo_col, shift, dh.p);
        SPMV_CUDA(cudaGetLastError());
        unsigned long long hist[256];
        SPMV_CUDA(cudaMemcpyAsync(hist, dh.p, sizeof hist, cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        int64_t nonempty = 0, ycost = 0;
        for (int64_t b = 0; b < nblocks; b++) {
            if (hist[b]) nonempty++;
            ycost += std::min<int64_t>(m->rows, (int64_t)hist[b]) * 8;
        }
        if ((nonempty * 8) << shift <= (int64_t)l2 / 4 * 3) return 0;
        if (n * 12 <= ycost) return 0;
        // ... and only when the gathers are NOT local already: a banded matrix (stencil) touches one or two column
        // blocks per million consecutive entries, and cutting its rows into blocks would only add sweeps over y
        const int64_t chunks = (n + ((int64_t)1 << 20) - 1) >> 20;
        Scratch<unsigned int> dm;
        SPMV_TRY(dm.alloc(chunks * 8));
        SPMV_CUDA(cudaMemsetAsync(dm.p, 0, sizeof(unsigned int) * 8 * (size_t)chunks, s));
        block_spread_kernel<<<dim3(16, (unsigned)chunks), 256, 0, s>>>(n, m->coo_col, shift, dm.p);
        SPMV_CUDA(cudaGetLastError());
        std::vector<unsigned int> masks((size_t)chunks * 8);
        SPMV_CUDA(cudaMemcpyAsync(masks.data(), dm.p, sizeof(unsigned int) * masks.size(), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        int64_t touched = 0;
        for (unsigned int v : masks) touched += __builtin_popcount(v);
        if (touched <= 3 * chunks) return 0;  // at most three blocks per chunk on average: local enough
    }
    Scratch<int> flag;
    SPMV_TRY(flag.alloc(1));
    SPMV_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
    rowcol_sorted_kernel<<<grid_for(n), 256, 0, s>>>(n, m->coo_row, m->coo_col, flag.p);
    SPMV_CUDA(cudaGetLastError());
    int unsorted = 0;
    SPMV_CUDA(cudaMemcpyAsync(&unsorted, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    if (unsorted) return 0;
    {
        Scratch<unsigned char> key, key2, tmp;
        Scratch<uint32_t> idx, idx2;
        SPMV_TRY(key.alloc(n)); SPMV_TRY(key2.alloc(n)); SPMV_TRY(idx.alloc(n)); SPMV_TRY(idx2.alloc(n));
        block_key_kernel<<<grid_for(n), 256, 0, s>>>(n, m->coo_col, shift, key.p, idx.p);
        SPMV_CUDA(cudaGetLastError());
        size_t tb = 0;
        const int end_bit = bits_for(nblocks);
        SPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, idx2.p, n, 0, end_bit, s));
        SPMV_TRY(tmp.alloc((int64_t)tb));
        SPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, idx2.p, n, 0, end_bit, s));
        int32_t *row2 = nullptr, *col2 = nullptr;
        double * val2 = nullptr;
        SPMV_TRY(alloc_streamed(m, &row2, n));
        SPMV_TRY(alloc_streamed(m, &col2, n));
        SPMV_TRY(alloc_streamed(m, &val2, n));
        gather3_kernel<<<grid_for(n), 256, 0, s>>>(n, idx2.p, m->coo_row, m->coo_col, m->coo_val, row2, col2, val2);
        SPMV_CUDA(cudaGetLastError());
        SPMV_CUDA(cudaStreamSynchronize(s));
        const int64_t cap = round_up(n, 4096) + kPadEntries;
        cudaFree(m->coo_row); cudaFree(m->coo_col); cudaFree(m->coo_val);
        m->device_bytes -= cap * 16;
        m->coo_row = row2; m->coo_col = col2; m->coo_val = val2;
    }
    m->coo_col_shift = shift;
    return 0;
}

// ---- hot-column tables for scattered gathers -----------------------------------------------------------------------------
// On a power-law matrix every lane of a gather hits its own 32 B sector: the COO kernel is bound by the sectors
// the L1 and the L2 can deliver (ncu: l1tex 86 %, lts 78 %, DRAM 57 % busy on R-MAT 2^24), not by HBM.  But the
// column popularity is as skewed as the row lengths: inside any stretch of a few hundred thousand entries, the 24 576 most
// referenced columns receive more than half of the gathers (8 192: a third).  The builder therefore cuts the entries
// into segments and gives every segment a table of its most referenced columns; the CTA that runs the segment
// loads their x values into shared memory once (192 KB) and serves those gathers from there -- no sector traffic, and a
// quarter of the L1 wavefronts of a divergent global load.
__global__ void hot_sample_spread_kernel(int64_t n, const int32_t * col, int samples, int64_t stride, unsigned int * distinct)
{
    // one warp per sampled span of 128 entries: how many distinct 128 B lines of x do they gather from?
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= samples) return;
    const int64_t base = ((int64_t)w * stride) & ~(int64_t)127;
    unsigned int count = 0;
    int lines[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t k = base + 4 * lane + j;
        lines[j] = k < n ? (col[k] >> 4) : -1 - lane;
    }
    // distinct among the 128 values: a value counts if no EARLIER entry of the span has it (O(128) per entry, tiny sample)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bool first = true;
        for (int src = 0; src < 32; ++src) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int other = __shfl_sync(0xffffffffu, lines[jj], src);
                if ((src < lane || (src == lane && jj < j)) && other == lines[j]) first = false;
            }
        }
        count += first ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane == 0) atomicAdd(distinct, count);
}

__global__ void hot_scatter_slots_kernel(int h, const int32_t * hot, int32_t * slot_of /* [columns] */, int value_is_slot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < h) slot_of[hot[i]] = value_is_slot ? i : -1;
}

__global__ void hot_remap_kernel(int64_t lo, int64_t hi, const int32_t * col, const int32_t * slot_of, int32_t * colh,
                                 unsigned long long * hits)
{
    unsigned long long mine = 0;
    for (int64_t k = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = col[k];
        const int32_t sl = slot_of[c];
        colh[k] = sl >= 0 ? (int32_t)(0x80000000u | (uint32_t)sl) : c;
        mine += sl >= 0 ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(hits, mine);
}

__global__ void hot_keep_repeated_kernel(int runs, int h, const int32_t * counts_desc, int * kept)
{
    // the runs are sorted by count, descending: keep the first h of them, but none that is referenced only once
    // (its x value would be fetched once either way)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < runs && i < h && counts_desc[i] >= 2 && (i + 1 == runs || i + 1 == h || counts_desc[i + 1] < 2)) *kept = i + 1;
}

int coo_build_hot(Matrix * m)
{
    if (m->coo_hot_tried) return 0;
    m->coo_hot_tried = true;
    const int64_t n = m->coo_n;
    // Opt-in only ("coo.hot" = 1; 2 = with the scattered-gather test below).  Measured on R-MAT 2^24 x 16 it LOSES to
    // the plain kernel (1.11 ms at best vs 1.04 ms, profiles/r02_sweep_q_coo_hot_columns.log): what bounds the divergent
    // gather is the number of misses the L1 can hold in flight, which shrinks with every KB of shared memory the tables take
    // (profiles/r02_sweep_r_coo_gather_paths.log: the plain kernel at 0/25/50/75/100 % carve-out: 1.05/1.09/1.22/1.81/3.46 ms).
    if (m->opt_coo_hot <= 0 || !m->coo_sorted || n >= ((int64_t)1 << 32)) return 0;
    cudaStream_t s = m->stream;
    if (m->opt_coo_hot == 2) {
        // scattered gathers?  sample 1024 spans of 128 entries: a banded matrix gathers from a handful of lines per
        // span, a power-law matrix from ~128
        const int samples = 1024;
        Scratch<unsigned int> d;
        SPMV_TRY(d.alloc(1));
        SPMV_CUDA(cudaMemsetAsync(d.p, 0, sizeof(unsigned int), s));
        hot_sample_spread_kernel<<<samples / 4, 128, 0, s>>>(n, m->coo_col, samples, std::max<int64_t>(128, n / samples), d.p);
        SPMV_CUDA(cudaGetLastError());
        unsigned int distinct = 0;
        SPMV_CUDA(cudaMemcpyAsync(&distinct, d.p, sizeof distinct, cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        if (distinct < 64u * samples) return 0;  // fewer than 64 distinct lines per 128 gathers: local enough
    }
    int h = (int)(m->opt_coo_hot_slots ? m->opt_coo_hot_slots : 24576);
    h = std::max(1024, std::min(27648, h / 1024 * 1024));
    const int threads = (int)(m->opt_coo_hot_threads ? m->opt_coo_hot_threads : 1024);
    const int ctas_per_sm = std::max(1, std::min(2048 / threads, (int)((int64_t)(227 * 1024 - 1024) / ((int64_t)h * 8))));
    const int per_cta = (int)(m->opt_coo_hot_segs ? m->opt_coo_hot_segs : 4);
    const int64_t want = (int64_t)m->sm_count * ctas_per_sm * per_cta;

    // ---- segments: equal pieces of every column block, cut at multiples of 128 entries ---------------------------------
    std::vector<int64_t> block_begin{0, n};
    if (m->coo_col_shift > 0) {
        const int64_t nblocks = (m->cols + ((int64_t)1 << m->coo_col_shift) - 1) >> m->coo_col_shift;
        Scratch<unsigned long long> dh;
        SPMV_TRY(dh.alloc(256));
        SPMV_CUDA(cudaMemsetAsync(dh.p, 0, 256 * sizeof(unsigned long long), s));
        block_hist_kernel<<<grid_for(n, m->sm_count), 256, 0, s>>>(n, m->coo_col, m->coo_col_shift, dh.p);
        SPMV_CUDA(cudaGetLastError());
        unsigned long long hist[256];
        SPMV_CUDA(cudaMemcpyAsync(hist, dh.p, sizeof hist, cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        block_begin.assign(1, 0);
        for (int64_t b = 0; b < nblocks; b++) block_begin.push_back(block_begin.back() + (int64_t)hist[b]);
    }
    std::vector<int64_t> seg{0};
    const int64_t target = std::max<int64_t>(4096, (n + want - 1) / want);
    for (size_t b = 0; b + 1 < block_begin.size(); b++) {
        const int64_t lo = block_begin[b], hi = block_begin[b + 1];
        if (hi <= lo) continue;
        const int64_t pieces = std::max<int64_t>(1, (hi - lo + target / 2) / target);
        for (int64_t q = 1; q < pieces; q++) {
            const int64_t cut = (lo + (hi - lo) * q / pieces) & ~(int64_t)127;
            if (cut > seg.back() && cut < hi) seg.push_back(cut);
        }
        seg.push_back(hi);
    }
    const int nseg = (int)seg.size() - 1;
    if (nseg < 1) return 0;

    // ---- per segment: the h most referenced columns ---------------------------------------------------------------------
    int64_t longest = 0;
    for (int q = 0; q < nseg; q++) longest = std::max(longest, seg[(size_t)q + 1] - seg[(size_t)q]);
    Scratch<int32_t> keys2, uniq, counts, uniq2, counts2, slot_of;
    Scratch<int> druns, dkept;
    Scratch<unsigned long long> dhits;
    Scratch<unsigned char> tmp;
    SPMV_TRY(keys2.alloc(longest)); SPMV_TRY(uniq.alloc(longest)); SPMV_TRY(counts.alloc(longest));
    SPMV_TRY(uniq2.alloc(longest)); SPMV_TRY(counts2.alloc(longest)); SPMV_TRY(slot_of.alloc(m->cols));
    SPMV_TRY(druns.alloc(1)); SPMV_TRY(dkept.alloc(1)); SPMV_TRY(dhits.alloc(1));
    size_t tb = 0, t1 = 0, t2 = 0, t3 = 0;
    const int col_bits = bits_for(m->cols);
    SPMV_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t1, (const int32_t *)m->coo_col, keys2.p, longest, 0, col_bits, s));
    SPMV_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, t2, keys2.p, uniq.p, counts.p, druns.p, longest, s));
    SPMV_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, t3, counts.p, counts2.p, uniq.p, uniq2.p, longest, 0, 32, s));
    tb = std::max(t1, std::max(t2, t3));
    SPMV_TRY(tmp.alloc((int64_t)tb));
    int32_t *colh = nullptr, *hot = nullptr;
    int64_t * dseg = nullptr;
    SPMV_TRY(alloc_streamed(m, &colh, n));
    m->coo_colh = colh;  // owned from here on (freed with the matrix, or below if the layout is rejected)
    SPMV_TRY(dev_alloc(m, &hot, (int64_t)nseg * h));
    m->coo_hot_cols = hot;
    SPMV_TRY(dev_alloc(m, &dseg, nseg + 1));
    m->coo_seg = dseg;
    SPMV_CUDA(cudaMemsetAsync(hot, 0, sizeof(int32_t) * (size_t)nseg * (size_t)h, s));
    SPMV_CUDA(cudaMemsetAsync(slot_of.p, 0xff, sizeof(int32_t) * (size_t)m->cols, s));
    SPMV_CUDA(cudaMemsetAsync(dhits.p, 0, sizeof(unsigned long long), s));
    SPMV_CUDA(cudaMemcpyAsync(dseg, seg.data(), sizeof(int64_t) * seg.size(), cudaMemcpyHostToDevice, s));
    for (int q = 0; q < nseg; q++) {
        const int64_t lo = seg[(size_t)q], len = seg[(size_t)q + 1] - lo;
        int32_t * table = hot + (int64_t)q * h;
        size_t t = tb;
        SPMV_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, t, (const int32_t *)(m->coo_col + lo), keys2.p, len, 0, col_bits, s));
        t = tb;
        SPMV_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, t, keys2.p, uniq.p, counts.p, druns.p, len, s));
        int runs = 0;
        SPMV_CUDA(cudaMemcpyAsync(&runs, druns.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        if (runs <= 0) continue;
        t = tb;
        SPMV_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, t, counts.p, counts2.p, uniq.p, uniq2.p, runs, 0, 32, s));
        SPMV_CUDA(cudaMemsetAsync(dkept.p, 0, sizeof(int), s));
        hot_keep_repeated_kernel<<<(std::min(runs, h) + 255) / 256, 256, 0, s>>>(runs, h, counts2.p, dkept.p);
        SPMV_CUDA(cudaGetLastError());
        int kept = 0;
        SPMV_CUDA(cudaMemcpyAsync(&kept, dkept.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        if (kept > 0) {
            SPMV_CUDA(cudaMemcpyAsync(table, uniq2.p, sizeof(int32_t) * (size_t)kept, cudaMemcpyDeviceToDevice, s));
            hot_scatter_slots_kernel<<<(kept + 255) / 256, 256, 0, s>>>(kept, table, slot_of.p, 1);
        }
        hot_remap_kernel<<<grid_for(len, m->sm_count), 256, 0, s>>>(lo, lo + len, m->coo_col, slot_of.p, colh, dhits.p);
        if (kept > 0) hot_scatter_slots_kernel<<<(kept + 255) / 256, 256, 0, s>>>(kept, table, slot_of.p, 0);
        SPMV_CUDA(cudaGetLastError());
    }
    unsigned long long hits = 0;
    SPMV_CUDA(cudaMemcpyAsync(&hits, dhits.p, sizeof hits, cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    m->coo_hot_coverage = n > 0 ? (double)hits / (double)n : 0.0;
    if (m->opt_coo_hot == 2 && m->coo_hot_coverage < 0.15) {  // not worth a persistent kernel with half the warps
        const int64_t cap = round_up(n, 4096) + kPadEntries;
        cudaFree(m->coo_colh); cudaFree(m->coo_hot_cols); cudaFree(m->coo_seg);
        m->coo_colh = nullptr; m->coo_hot_cols = nullptr; m->coo_seg = nullptr;
        m->device_bytes -= cap * 4 + (int64_t)nseg * h * 4 + (nseg + 1) * 8;
        return 0;
    }
    m->coo_hot_h = h;
    m->coo_nseg = nseg;
    m->aux_dirty = true;
    return 0;
}

// Copies of the COO arrays back in row-major order (undoes coo_column_blocks): stable sort by row.
int coo_row_major_copy(Matrix * m, int32_t * row2, int32_t * col2, double * val2)
{
    const int64_t n = m->coo_n;
    cudaStream_t s = m->stream;
    Scratch<uint32_t> idx, idx2;
    Scratch<int32_t> key2;
    Scratch<unsigned char> tmp;
    SPMV_TRY(idx.alloc(n)); SPMV_TRY(idx2.alloc(n)); SPMV_TRY(key2.alloc(n));
    iota_kernel<<<grid_for(n), 256, 0, s>>>(n, idx.p);
    SPMV_CUDA(cudaGetLastError());
    size_t tb = 0;
    const int end_bit = bits_for(m->rows);
    SPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, m->coo_row, key2.p, idx.p, idx2.p, n, 0, end_bit, s));
    SPMV_TRY(tmp.alloc((int64_t)tb));
    SPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, m->coo_row, key2.p, idx.p, idx2.p, n, 0, end_bit, s));
    gather3_kernel<<<grid_for(n), 256, 0, s>>>(n, idx2.p, m->coo_row, m->coo_col, m->coo_val, row2, col2, val2);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int coo_adopt(Matrix * m, int64_t rows, int64_t cols, int64_t n, int32_t * row, int32_t * col, double * val,
              int mode, bool already_sorted)
{
    cudaStream_t s = m->stream;
    m->format = SPMVB200_COO;
    m->coo_mode = mode;
    m->rows = rows; m->cols = cols; m->stored = n; m->coo_n = n;
    if (m->nnz == 0) m->nnz = n;
    m->coo_row = row; m->coo_col = col; m->coo_val = val;
    m->coo_sorted = already_sorted;
    if (mode == SPMVB200_COO_SEGMENTED && !already_sorted && n > 1) {
        if (n >= ((int64_t)1 << 32)) return fail(SPMVB200_ERR_UNSUPPORTED, "COO sort supports < 2^32 entries");
        Scratch<int> flag;
        SPMV_TRY(flag.alloc(1));
        SPMV_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
        is_sorted_kernel<<<grid_for(n), 256, 0, s>>>(n, row, flag.p);
        SPMV_CUDA(cudaGetLastError());
        int unsorted = 0;
        SPMV_CUDA(cudaMemcpyAsync(&unsorted, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        if (unsorted) {
            // stable sort by row only: the column order inside a row stays the file order
            Scratch<uint32_t> idx, idx2;
            Scratch<unsigned char> tmp;
            int32_t *row2 = nullptr, *col2 = nullptr;
            double * val2 = nullptr;
            SPMV_TRY(idx.alloc(n)); SPMV_TRY(idx2.alloc(n));
            SPMV_TRY(alloc_streamed(m, &row2, n));
            SPMV_TRY(alloc_streamed(m, &col2, n));
            SPMV_TRY(alloc_streamed(m, &val2, n));
            iota_kernel<<<grid_for(n), 256, 0, s>>>(n, idx.p);
            size_t tb = 0;
            const int end_bit = bits_for(rows);
            SPMV_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, row, row2, idx.p, idx2.p, n, 0, end_bit, s));
            SPMV_TRY(tmp.alloc((int64_t)tb));
            SPMV_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, row, row2, idx.p, idx2.p, n, 0, end_bit, s));
            gather_kernel<<<grid_for(n), 256, 0, s>>>(n, idx2.p, col, val, col2, val2);
            SPMV_CUDA(cudaGetLastError());
            SPMV_CUDA(cudaStreamSynchronize(s));
            const int64_t cap = round_up(n, 4096) + kPadEntries;
            cudaFree(row); cudaFree(col); cudaFree(val);
            m->device_bytes -= cap * 16;
            m->coo_row = row2; m->coo_col = col2; m->coo_val = val2;
        }
        m->coo_sorted = true;
    }
    SPMV_CUDA(cudaStreamSynchronize(s));
    if (mode == SPMVB200_COO_SEGMENTED) SPMV_TRY(coo_column_blocks(m));
    return 0;
}

}  // namespace spmvb200
