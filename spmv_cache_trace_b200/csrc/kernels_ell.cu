// kernels_ell.cu -- ELLPACK SpMV (see kernels_csr.cu for the design notes shared by all kernels).
//
// Replaces ell_spmv + ell_spmv_inner_loop[_skip_padding] (reference matrix/ell-matrix.cpp:243-307) and
// the ELL pass of the hybrid format (matrix/hybrid-matrix.cpp:422-452).
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// =============================================================================================
// ELLPACK (column-major)
// =============================================================================================

template <int R>
struct EllLoad;
template <>
struct EllLoad<1> {
    static __device__ __forceinline__ void cols(const int32_t * p, int (&c)[1], uint64_t pol) { c[0] = ldg_stream_i1(p, pol); }
    static __device__ __forceinline__ void vals(const double * p, double (&a)[1], uint64_t pol) { a[0] = ldg_stream_d1(p, pol); }
};
template <>
struct EllLoad<2> {
    static __device__ __forceinline__ void cols(const int32_t * p, int (&c)[2], uint64_t pol)
    {
        int2 v = ldg_stream_i2(p, pol);
        c[0] = v.x; c[1] = v.y;
    }
    static __device__ __forceinline__ void vals(const double * p, double (&a)[2], uint64_t pol)
    {
        double2 v = ldg_stream_d2(p, pol);
        a[0] = v.x; a[1] = v.y;
    }
};
template <>
struct EllLoad<4> {
    static __device__ __forceinline__ void cols(const int32_t * p, int (&c)[4], uint64_t pol)
    {
        int4 v = ldg_stream_i4(p, pol);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    }
    static __device__ __forceinline__ void vals(const double * p, double (&a)[4], uint64_t)
    {
        ldg_stream_d4(p, a);  // one 256-bit load
    }
};

// W_STATIC > 0: the row length is a compile-time constant and the slot loop is fully unrolled,
// so all of a thread's matrix loads are issued before the first gather returns.
template <int R, int W_STATIC, bool SKIP>
__global__ void __launch_bounds__(256)
ell_kernel(int64_t row0, int64_t rows, int64_t pitch, int w_runtime, int independent, int store, double alpha,
           const int32_t * __restrict__ col,
           const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y,
           const double * __restrict__ y_in_host, double * __restrict__ y_out_host)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // rows [row0, rows) of the matrix (row0 a multiple of 4: the vector loads stay aligned)
    const int64_t i0 = row0 + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * R;
    if (i0 >= rows) return;
    // Zero-copy form (spmvb200_spmv_host with pinned buffers): y_old is read from and y_new written
    // to mapped HOST memory by this kernel, so the two PCIe directions run at the same time.
    // (y_in_host may also point at device memory: the chunked path uploads y with the copy engine.)
    double yo[R];
    const bool whole = i0 + R <= rows;  // vector access keeps PCIe transactions full-width
    if (y_in_host && whole && R == 2) {
        const double2 t = __ldcs(reinterpret_cast<const double2 *>(y_in_host + i0));
        yo[0] = t.x; yo[R - 1] = t.y;
    } else if (y_in_host && whole && R == 4) {
        const double2 t = __ldcs(reinterpret_cast<const double2 *>(y_in_host + i0));
        const double2 u = __ldcs(reinterpret_cast<const double2 *>(y_in_host + i0 + 2));
        yo[0] = t.x; yo[1 % R] = t.y; yo[2 % R] = u.x; yo[3 % R] = u.y;
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) yo[r] = (y_in_host && i0 + r < rows) ? __ldcs(y_in_host + i0 + r) : 0.0;
    }
    const int W = W_STATIC > 0 ? W_STATIC : w_runtime;
    const uint64_t pol = policy_evict_first();
    double z[R];
    bool live[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { z[r] = 0.0; live[r] = true; }

    if (W_STATIC > 0) {
        int c[W_STATIC > 0 ? W_STATIC : 1][R];
        double a[W_STATIC > 0 ? W_STATIC : 1][R];
#pragma unroll
        for (int l = 0; l < W_STATIC; ++l) {
            EllLoad<R>::cols(col + (int64_t)l * pitch + i0, c[l], pol);
            EllLoad<R>::vals(val + (int64_t)l * pitch + i0, a[l], pol);
        }
        // The matrix is immutable; x and y may have been written by the previous launch.
        if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
        double xv[W_STATIC > 0 ? W_STATIC : 1][R];
#pragma unroll
        for (int l = 0; l < W_STATIC; ++l)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (SKIP) live[r] = live[r] && (c[l][r] != INT_MAX);
                xv[l][r] = (!SKIP || live[r]) ? ldx(x + c[l][r]) : 0.0;
            }
#pragma unroll
        for (int l = 0; l < W_STATIC; ++l)
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (!SKIP || c[l][r] != INT_MAX) z[r] = __dadd_rn(z[r], __dmul_rn(a[l][r], xv[l][r]));
    } else {
        if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll 4
        for (int l = 0; l < W; ++l) {
            int c[R];
            double a[R];
            EllLoad<R>::cols(col + (int64_t)l * pitch + i0, c, pol);
            EllLoad<R>::vals(val + (int64_t)l * pitch + i0, a, pol);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (SKIP) live[r] = live[r] && (c[r] != INT_MAX);
                if (!SKIP || live[r]) z[r] = __dadd_rn(z[r], __dmul_rn(a[r], ldx(x + c[r])));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) z[r] = __dmul_rn(alpha, z[r]);  // alpha = 1.0 is exact
    if (y_out_host) {
#pragma unroll
        for (int r = 0; r < R; ++r) z[r] = __dadd_rn(yo[r], z[r]);
        if (whole && R == 2) {
            __stcs(reinterpret_cast<double2 *>(y_out_host + i0), make_double2(z[0], z[R - 1]));
        } else if (whole && R == 4) {
            __stcs(reinterpret_cast<double2 *>(y_out_host + i0), make_double2(z[0], z[1 % R]));
            __stcs(reinterpret_cast<double2 *>(y_out_host + i0 + 2), make_double2(z[2 % R], z[3 % R]));
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (i0 + r < rows) __stcs(y_out_host + i0 + r, z[r]);
        }
        return;
    }
    if (store) {  // y = alpha*A*x: a thread owns its rows
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (i0 + r < rows) y[i0 + r] = z[r];
        return;
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (i0 + r < rows) red_add_f64(y + i0 + r, z[r]);
}

template <int R, bool SKIP>
static int launch_ell_rw(Matrix * m, int block)
{
    // optional row range (the pipelined host-buffer path runs the matrix in row chunks)
    const int64_t row0 = m->range_end > 0 ? m->range_begin : 0;
    const int64_t row1 = m->range_end > 0 ? m->range_end : m->rows;
    const int64_t threads = (row1 - row0 + R - 1) / R;
    const unsigned grid = (unsigned)((threads + block - 1) / block);
    const int w = (int)m->ell_w;
    const RunMode rm = run_mode(m);
    const bool pdl = rm.pdl;
    const int indep = rm.independent;
    const int store = (m->run_beta0 && !m->host_y_out) ? 1 : 0;
    m->run_beta0 = false;
    cudaError_t e;
#define SPMV_ELL_CASE(WS)                                                                                      \
    case WS:                                                                                                   \
        e = launch_kernel(ell_kernel<R, WS, SKIP>, grid, (unsigned)block, 0, m->stream, pdl, row0, row1,    \
                          m->ell_pitch, w, indep, store, m->alpha, (const int32_t *)m->ell_col, (const double *)m->ell_val,            \
                          (const double *)m->x, m->y, (const double *)m->host_y_in, m->host_y_out);              \
        break;
    switch (w) {
        SPMV_ELL_CASE(1) SPMV_ELL_CASE(2) SPMV_ELL_CASE(3) SPMV_ELL_CASE(4) SPMV_ELL_CASE(5)
        SPMV_ELL_CASE(6) SPMV_ELL_CASE(7) SPMV_ELL_CASE(8) SPMV_ELL_CASE(9)
    default:
        e = launch_kernel(ell_kernel<R, 0, SKIP>, grid, (unsigned)block, 0, m->stream, pdl, row0, row1, m->ell_pitch, w, indep, store, m->alpha,
                          (const int32_t *)m->ell_col, (const double *)m->ell_val, (const double *)m->x, m->y,
                          (const double *)m->host_y_in, m->host_y_out);
    }
#undef SPMV_ELL_CASE
    SPMV_CUDA(e);
    count_launch();
    return 0;
}

int launch_ell(Matrix * m, bool)
{
    if (m->rows == 0) return 0;
    if (m->ell_w == 0) return clear_y_for_beta0(m);  // nothing to add; y = alpha*A*x still has to clear y
    int R = (int)(m->opt_ell_rows ? m->opt_ell_rows : 2);
    int block = (int)(m->opt_ell_block ? m->opt_ell_block : 128);
    if (block < 32 || block > 256 || block % 32) return fail(SPMVB200_ERR_INVALID, "ell.block must be 32..256");
    m->kernel_name = "ell_kernel";
    const bool skip = m->skip_padding != 0;
    if (R == 1) return skip ? launch_ell_rw<1, true>(m, block) : launch_ell_rw<1, false>(m, block);
    if (R == 2) return skip ? launch_ell_rw<2, true>(m, block) : launch_ell_rw<2, false>(m, block);
    if (R == 4) return skip ? launch_ell_rw<4, true>(m, block) : launch_ell_rw<4, false>(m, block);
    return fail(SPMVB200_ERR_INVALID, "ell.rows_per_thread must be 1, 2 or 4");
}

}  // namespace spmvb200
