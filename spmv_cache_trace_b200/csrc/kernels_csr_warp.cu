// kernels_csr_warp.cu -- CSR SpMV, warp-granular variant:  y += A*x.
//
// Same job as csr_stream_kernel (kernels_csr.cu; reference matrix/csr-matrix-spmv.cpp:21-33, 63-76)
// with the ELL kernel's execution shape: short-lived warps, all matrix loads of a warp issued up
// front into REGISTERS, no CTA-wide barrier, no mbarrier ring, hardware block scheduling.
//
//   * The non-zeros are cut into spans of SPAN (256) consecutive entries; one warp owns one span.
//     Its 32 lanes load the span's values and column indices with 128/64-bit coalesced streaming
//     loads (SPAN*12 B in flight per warp, held in registers; with ~40 resident warps per SM that
//     is >100 KB in flight per SM without occupying shared memory while the data is in flight).
//   * When the data lands it is parked in a warp-private slice of shared memory (__syncwarp only),
//     and the row pass of the stream kernel's direct mode runs on it: G lanes per row, the lanes of
//     a warp hold consecutive rows so the x gathers of a banded matrix are contiguous, eight
//     gathers in flight per lane; slices longer than kLongRow*G are summed by the whole warp.
//   * span_row[w] (the row holding the first entry of span w) comes from a table built with the
//     matrix; the row pointers of the span's rows are read straight from global memory (coalesced,
//     prefetched one round ahead).
//   * y += sum through RED.ADD.F64; programmatic dependent launch with the wait placed after the
//     matrix loads; evict-first L2 policy on the streamed arrays.
//
// Arithmetic order as in kernels_csr.cu: G = 1 rows that lie inside one span are summed left to
// right with separate multiply/add roundings (bit-identical to the reference's scalar loop).
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>

namespace spmvb200 {

using namespace ptx;

constexpr int kWarpSpan = 256;
constexpr int kWarpLongRow = 96;

template <typename OffT>
__global__ void csr_span_rows_kernel(int64_t rows, int64_t nspans, int64_t span, const OffT * __restrict__ rp,
                                     int32_t * __restrict__ span_row)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w > nspans) return;
    if (w == nspans) {
        span_row[w] = (int32_t)(rows - 1);
        return;
    }
    const int64_t target = w * span;  // largest r in [0, rows-1] with rp[r] <= target
    int64_t lo = 0, hi = rows - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if ((int64_t)rp[mid] <= target) lo = mid; else hi = mid - 1;
    }
    span_row[w] = (int32_t)lo;
}

template <typename OffT, int G, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
csr_warp_kernel(int64_t stored, int64_t nspans, int independent, const OffT * __restrict__ rp, const int32_t * __restrict__ col,
                const double * __restrict__ val, const int32_t * __restrict__ span_row,
                const double * __restrict__ x, double * __restrict__ y)
{
    constexpr int SPAN = kWarpSpan;
    constexpr int VEC = SPAN / 64;  // 128-bit value loads (two entries) per lane
    __shared__ __align__(16) double sval[WARPS][SPAN];
    __shared__ __align__(16) int32_t scol[WARPS][SPAN];

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * WARPS + warp;
    if (w >= nspans) return;  // no CTA-wide barrier below
    const int64_t base = w * SPAN;
    const int64_t end = min(base + (int64_t)SPAN, stored);
    const uint64_t pol = policy_evict_first();

    // 1. the whole span in flight: 2*VEC vector loads per lane (arrays are padded past `stored`)
    double2 v[VEC];
    int2 c[VEC];
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
        v[u] = ldg_stream_d2(val + base + 2 * lane + 64 * u, pol);
        c[u] = ldg_stream_i2(col + base + 2 * lane + 64 * u, pol);
    }
    const int r0 = __ldg(span_row + w);
    const int r1 = __ldg(span_row + w + 1);

    // 2. row extents of the first round of rows (RPW rows per round)
    constexpr int RPW = 32 / G;
    const int g = lane % G;
    int r = r0 + lane / G;
    int64_t lo = 0, hi = 0;
    if (r <= r1) {
        lo = (int64_t)rp[r];
        hi = (int64_t)rp[r + 1];
    }

    // 3. park the span in the warp's slice of shared memory
    double * pv = sval[warp];
    int32_t * pc = scol[warp];
#pragma unroll
    for (int u = 0; u < VEC; ++u) {
        *reinterpret_cast<double2 *>(pv + 2 * lane + 64 * u) = v[u];
        *reinterpret_cast<int2 *>(pc + 2 * lane + 64 * u) = c[u];
    }
    __syncwarp();

    // The matrix is immutable; x and y may have been written by the previous launch.
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");

    // 4. row pass
    for (int rw = r0; rw <= r1; rw += RPW) {  // warp-uniform
        const int rn = r + RPW;
        int64_t lon = 0, hin = 0;
        if (rw + RPW <= r1 && rn <= r1) {  // prefetch the next round's extents
            lon = (int64_t)rp[rn];
            hin = (int64_t)rp[rn + 1];
        }
        int a = 0, b = 0;
        if (r <= r1) {
            a = (int)((lo > base ? lo : base) - base);
            b = (int)((hi < end ? hi : end) - base);
            if (b < a) b = a;
        }
        const bool is_long = (b - a) > kWarpLongRow * G;
        const int bn = is_long ? a : b;
        double sum = 0.0;
        for (int kb = a; kb < bn; kb += 8 * G) {
            double pvv[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = kb + g + u * G;
                const bool ok = k < bn;
                pvv[u] = ok ? pv[k] : 0.0;
                xv[u] = ok ? ldx(x + pc[k]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kb + g + u * G < bn) sum = __dadd_rn(sum, __dmul_rn(pvv[u], xv[u]));
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (g == 0 && bn > a) red_add_f64(y + r, sum);

        unsigned longs = __ballot_sync(0xffffffffu, is_long && g == 0);
        while (longs) {
            const int src = __ffs(longs) - 1;
            longs &= longs - 1;
            const int la = __shfl_sync(0xffffffffu, a, src);
            const int lb = __shfl_sync(0xffffffffu, b, src);
            const int lr = __shfl_sync(0xffffffffu, r, src);
            double ls = 0.0;
            for (int k = la + lane; k < lb; k += 32) ls = __dadd_rn(ls, __dmul_rn(pv[k], ldx(x + pc[k])));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
            if (lane == 0) red_add_f64(y + lr, ls);
        }
        r = rn; lo = lon; hi = hin;
    }
}

// span_row[w] = row holding entry w*span, for the warp-granular kernels (span 256 here, 128 in
// kernels_csr_flat.cu); rebuilt when a kernel with a different span size is selected.
int csr_build_span_table(Matrix * m, int span)
{
    if (m->span_row && m->span_size == span) return 0;
    if (m->span_row) {
        cudaFree(m->span_row);
        m->span_row = nullptr;
    }
    const int64_t nspans = (m->stored + span - 1) / span;
    SPMV_TRY(dev_alloc((Matrix *)nullptr, &m->span_row, nspans + 1));
    m->span_size = span;
    const unsigned grid = (unsigned)((nspans + 1 + 255) / 256);
    if (m->off64) csr_span_rows_kernel<int64_t><<<grid, 256, 0, m->stream>>>(m->rows, nspans, span, (const int64_t *)m->rp, m->span_row);
    else csr_span_rows_kernel<uint32_t><<<grid, 256, 0, m->stream>>>(m->rows, nspans, span, (const uint32_t *)m->rp, m->span_row);
    SPMV_CUDA(cudaGetLastError());
    m->aux_dirty = true;
    return 0;
}

template <typename OffT, int G, int WARPS>
static int launch_warp_variant(Matrix * m)
{
    if (m->dry_run) return 0;
    const int64_t nspans = (m->stored + kWarpSpan - 1) / kWarpSpan;
    const unsigned grid = (unsigned)((nspans + WARPS - 1) / WARPS);
    const RunMode rm = run_mode(m);
    SPMV_CUDA(launch_kernel(csr_warp_kernel<OffT, G, WARPS>, grid, WARPS * 32u, 0, m->stream, rm.pdl, m->stored,
                            nspans, rm.independent, (const OffT *)m->rp, (const int32_t *)m->col, (const double *)m->val,
                            (const int32_t *)m->span_row, (const double *)m->x, m->y));
    count_launch();
    return 0;
}

template <typename OffT, int G>
static int launch_warp_warps(Matrix * m, int warps)
{
    switch (warps) {
    case 2: return launch_warp_variant<OffT, G, 2>(m);
    case 4: return launch_warp_variant<OffT, G, 4>(m);
    case 8: return launch_warp_variant<OffT, G, 8>(m);
    }
    return fail(SPMVB200_ERR_INVALID, "csr.threads must be 64, 128 or 256 for the warp kernel");
}

template <typename OffT>
static int launch_warp_t(Matrix * m, int lanes, int warps)
{
    SPMV_TRY(csr_build_span_table(m, kWarpSpan));
    switch (lanes) {
    case 1: return launch_warp_warps<OffT, 1>(m, warps);
    case 2: return launch_warp_warps<OffT, 2>(m, warps);
    case 4: return launch_warp_warps<OffT, 4>(m, warps);
    case 8: return launch_warp_warps<OffT, 8>(m, warps);
    }
    return fail(SPMVB200_ERR_INVALID, "csr.lanes must be 1, 2, 4 or 8");
}

int launch_csr_warp(Matrix * m, int lanes)
{
    const int threads = (int)(m->opt_csr_threads ? m->opt_csr_threads : 128);
    m->kernel_name = "csr_warp_kernel";
    return m->off64 ? launch_warp_t<int64_t>(m, lanes, threads / 32) : launch_warp_t<uint32_t>(m, lanes, threads / 32);
}

}  // namespace spmvb200
