// kernels.cu -- the hand-written sm_100a SpMV kernels and their launchers.
//
//   csr_stream_kernel   y += A*x, CSR.  Replaces csr_spmv + csr_spmv_inner_loop
//                       (reference matrix/csr-matrix-spmv.cpp:21-33, 63-76).
//   ell_kernel          y += A*x, ELLPACK (column-major on the device).  Replaces ell_spmv +
//                       ell_spmv_inner_loop[_skip_padding] (matrix/ell-matrix.cpp:243-307).
//   coo_segmented_kernel / coo_atomic_kernel
//                       y += A*x, COO.  Replace coo_spmv (matrix/coo-matrix.cpp:248-285) and
//                       coo_spmv_atomic (:287-309); also the COO tail of the hybrid format
//                       (matrix/hybrid-matrix.cpp:491-528).
//
// All three are HBM-bandwidth bound (2 flop per 12-16 streamed bytes), so the design is about
// keeping many bytes in flight and touching every matrix byte exactly once:
//   * CSR and COO stream fixed-size tiles of NON-ZEROS (not rows) through shared memory with 1-D
//     bulk-async copies (TMA engine, cp.async.bulk + mbarrier) in a multi-stage ring, so the
//     global loads are perfectly coalesced and balanced whatever the row lengths are.  Products
//     a[k]*x[j[k]] overwrite the staged values; a second phase sums each row's slice of the tile.
//   * ELL is stored column-major so a warp reads 32*R consecutive rows of one slot with 128-bit
//     loads; each thread owns R consecutive rows.
// x is gathered through the read-only L1/L2 path; streamed matrix data is marked evict-first so
// it does not push x out of L2.
//
// Arithmetic order: rows that fit inside one tile (CSR) / every row (ELL) are summed left to
// right with separate multiply and add roundings (__dmul_rn/__dadd_rn), which is exactly what the
// reference's scalar loops do when compiled for baseline x86-64 (no FMA contraction).  Rows cut
// by a tile boundary, rows longer than kLongRow in a tile, and all COO sums are combined in a
// different order (fp64 reductions in L2), within BASELINE.json's per-row tolerance.
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// =============================================================================================
// CSR
// =============================================================================================

constexpr int kLongRow = 96;  // pieces longer than this are summed by a whole warp

template <typename OffT>
__global__ void csr_tile_rows_kernel(int64_t rows, int64_t ntiles, int64_t tile, const OffT * __restrict__ rp,
                                     int32_t * __restrict__ tile_row)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) {
        tile_row[t] = (int32_t)(rows - 1);
        return;
    }
    // largest r in [0, rows-1] with rp[r] <= t*tile
    const int64_t target = t * tile;
    int64_t lo = 0, hi = rows - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if ((int64_t)rp[mid] <= target) lo = mid; else hi = mid - 1;
    }
    tile_row[t] = (int32_t)lo;
}

template <typename OffT, int TILE, int STAGES>
__global__ void __launch_bounds__(kCsrThreads)
csr_stream_kernel(int64_t ntiles, const OffT * __restrict__ rp, const int32_t * __restrict__ col,
                  const double * __restrict__ val, const int32_t * __restrict__ tile_row,
                  const double * __restrict__ x, double * __restrict__ y)
{
    constexpr int T = kCsrThreads;
    constexpr int PER = TILE / T;
    static_assert(TILE % T == 0 && TILE % 4 == 0, "tile must be a multiple of the block and of 16 bytes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double * sval = reinterpret_cast<double *>(smem_raw);
    int32_t * scol = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * TILE * 8);
    uint64_t * full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TILE * 12);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    uint64_t policy = 0;

    auto issue = [&](int s, int64_t t) {
        mbar_arrive_expect_tx(&full[s], TILE * 12);
        bulk_g2s(sval + (size_t)s * TILE, val + t * TILE, TILE * 8, &full[s], policy);
        bulk_g2s(scol + (size_t)s * TILE, col + t * TILE, TILE * 4, &full[s], policy);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        policy = policy_evict_first();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (t < ntiles) issue(s, t);
        }
    }

    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = (int)(it % STAGES);
        const uint32_t parity = (uint32_t)((it / STAGES) & 1);
        const int r0 = __ldg(tile_row + t);
        const int r1 = __ldg(tile_row + t + 1);
        const int64_t base = t * TILE;
        const int64_t end = base + TILE;

        // Row extents (and the old y) of this thread's first row: in flight while the tile lands.
        int r = r0 + tid;
        int64_t lo = 0, hi = 0;
        double yo = 0.0;
        if (r <= r1) {
            lo = (int64_t)rp[r];
            hi = (int64_t)rp[r + 1];
            yo = y[r];
        }

        mbar_wait(&full[s], parity);
        double * pv = sval + (size_t)s * TILE;
        const int32_t * pc = scol + (size_t)s * TILE;

        // Phase 1: products, one coalesced pass over the tile, PER gathers in flight per thread.
        {
            double a[PER], xv[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                a[i] = pv[tid + i * T];
                xv[i] = __ldg(x + pc[tid + i * T]);
            }
#pragma unroll
            for (int i = 0; i < PER; ++i) pv[tid + i * T] = __dmul_rn(a[i], xv[i]);
        }
        __syncthreads();

        // Phase 2: one thread per row sums its slice [a, b) of the tile, left to right.
        while (r - lane <= r1) {  // warp-uniform trip count
            const int rn = r + T;
            int64_t lon = 0, hin = 0;
            double yon = 0.0;
            if (rn <= r1) {
                lon = (int64_t)rp[rn];
                hin = (int64_t)rp[rn + 1];
                yon = y[rn];
            }
            int a = 0, b = 0;
            bool partial = false;
            if (r <= r1) {
                a = (int)((lo > base ? lo : base) - base);
                b = (int)((hi < end ? hi : end) - base);
                partial = (lo < base) || (hi > end);
            }
            const bool is_long = (b - a) > kLongRow;
            if (b > a && !is_long) {
                double sum = 0.0;
                for (int k = a; k < b; ++k) sum = __dadd_rn(sum, pv[k]);
                if (partial) red_add_f64(y + r, sum);
                else y[r] = __dadd_rn(yo, sum);
            }
            // Long pieces: the whole warp sums each one (strided, then a shuffle tree).
            unsigned longs = __ballot_sync(0xffffffffu, is_long);
            while (longs) {
                const int src = __ffs(longs) - 1;
                longs &= longs - 1;
                const int la = __shfl_sync(0xffffffffu, a, src);
                const int lb = __shfl_sync(0xffffffffu, b, src);
                const int lr = __shfl_sync(0xffffffffu, r, src);
                double sum = 0.0;
                for (int k = la + lane; k < lb; k += 32) sum = __dadd_rn(sum, pv[k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) red_add_f64(y + lr, sum);
            }
            r = rn; lo = lon; hi = hin; yo = yon;
        }
        __syncthreads();

        if (tid == 0) {
            const int64_t tn = t + (int64_t)STAGES * gridDim.x;
            if (tn < ntiles) {
                fence_proxy_async();
                issue(s, tn);
            }
        }
    }
}

template <typename OffT>
static int csr_build_tiles_t(Matrix * m, int tile)
{
    if (m->tile_row) {
        cudaFree(m->tile_row);
        m->device_bytes -= (int64_t)sizeof(int32_t) * (m->ntiles + 1);
        m->tile_row = nullptr;
    }
    m->csr_tile = tile;
    m->ntiles = (m->stored + tile - 1) / tile;
    SPMV_TRY(dev_alloc(m, &m->tile_row, m->ntiles + 1));
    if (m->rows == 0) return 0;
    const int64_t n = m->ntiles + 1;
    csr_tile_rows_kernel<OffT><<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(
        m->rows, m->ntiles, tile, (const OffT *)m->rp, m->tile_row);
    SPMV_CUDA(cudaGetLastError());
    return 0;
}

int csr_build_tiles(Matrix * m)
{
    int tile = (int)(m->opt_csr_tile ? m->opt_csr_tile : 2048);
    if (tile != 1024 && tile != 2048 && tile != 4096) return fail(SPMVB200_ERR_INVALID, "csr.tile must be 1024, 2048 or 4096");
    return m->off64 ? csr_build_tiles_t<int64_t>(m, tile) : csr_build_tiles_t<uint32_t>(m, tile);
}

template <typename OffT, int TILE, int STAGES>
static int launch_csr_variant(Matrix * m, int ctas_per_sm)
{
    auto kernel = csr_stream_kernel<OffT, TILE, STAGES>;
    const size_t smem = (size_t)STAGES * TILE * 12 + 8 * STAGES + 16;
    static bool configured = false;  // per instantiation
    if (!configured) {
        SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int64_t grid = (int64_t)m->sm_count * ctas_per_sm;
    if (grid > m->ntiles) grid = m->ntiles;
    if (grid < 1) return 0;
    kernel<<<(unsigned)grid, kCsrThreads, smem, m->stream>>>(m->ntiles, (const OffT *)m->rp, m->col, m->val,
                                                            m->tile_row, m->x, m->y);
    SPMV_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

template <typename OffT>
static int launch_csr_t(Matrix * m)
{
    const int tile = m->csr_tile;
    int stages = (int)m->opt_csr_stages;
    int ctas = (int)m->opt_csr_ctas;
    if (tile == 1024) {
        if (!stages) stages = 4;
        if (!ctas) ctas = 4;
        if (stages == 1) return launch_csr_variant<OffT, 1024, 1>(m, ctas);
        if (stages == 2) return launch_csr_variant<OffT, 1024, 2>(m, ctas);
        if (stages == 4) return launch_csr_variant<OffT, 1024, 4>(m, ctas);
    } else if (tile == 2048) {
        if (!stages) stages = 3;
        if (!ctas) ctas = 3;
        if (stages == 1) return launch_csr_variant<OffT, 2048, 1>(m, ctas);
        if (stages == 2) return launch_csr_variant<OffT, 2048, 2>(m, ctas);
        if (stages == 3) return launch_csr_variant<OffT, 2048, 3>(m, ctas);
    } else if (tile == 4096) {
        if (!stages) stages = 2;
        if (!ctas) ctas = 2;
        if (stages == 1) return launch_csr_variant<OffT, 4096, 1>(m, ctas);
        if (stages == 2) return launch_csr_variant<OffT, 4096, 2>(m, ctas);
    }
    return fail(SPMVB200_ERR_INVALID, "unsupported csr.tile / csr.stages combination");
}

int launch_csr(Matrix * m)
{
    if (m->rows == 0 || m->stored == 0) return 0;
    const int want = (int)(m->opt_csr_tile ? m->opt_csr_tile : 2048);
    if (!m->tile_row || m->csr_tile != want) SPMV_TRY(csr_build_tiles(m));
    m->kernel_name = "csr_stream_kernel";
    return m->off64 ? launch_csr_t<int64_t>(m) : launch_csr_t<uint32_t>(m);
}

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
ell_kernel(int64_t rows, int64_t pitch, int w_runtime, const int32_t * __restrict__ col,
           const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y)
{
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * R;
    if (i0 >= rows) return;
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
        double xv[W_STATIC > 0 ? W_STATIC : 1][R];
#pragma unroll
        for (int l = 0; l < W_STATIC; ++l)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (SKIP) live[r] = live[r] && (c[l][r] != INT_MAX);
                xv[l][r] = (!SKIP || live[r]) ? __ldg(x + c[l][r]) : 0.0;
            }
#pragma unroll
        for (int l = 0; l < W_STATIC; ++l)
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (!SKIP || c[l][r] != INT_MAX) z[r] = __dadd_rn(z[r], __dmul_rn(a[l][r], xv[l][r]));
    } else {
#pragma unroll 4
        for (int l = 0; l < W; ++l) {
            int c[R];
            double a[R];
            EllLoad<R>::cols(col + (int64_t)l * pitch + i0, c, pol);
            EllLoad<R>::vals(val + (int64_t)l * pitch + i0, a, pol);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (SKIP) live[r] = live[r] && (c[r] != INT_MAX);
                if (!SKIP || live[r]) z[r] = __dadd_rn(z[r], __dmul_rn(a[r], __ldg(x + c[r])));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (i0 + r < rows) y[i0 + r] = __dadd_rn(y[i0 + r], z[r]);
}

template <int R, bool SKIP>
static int launch_ell_rw(Matrix * m, int block)
{
    const int64_t threads = (m->rows + R - 1) / R;
    const unsigned grid = (unsigned)((threads + block - 1) / block);
    const int w = (int)m->ell_w;
#define SPMV_ELL_CASE(WS)                                                                              \
    case WS:                                                                                           \
        ell_kernel<R, WS, SKIP><<<grid, block, 0, m->stream>>>(m->rows, m->ell_pitch, w, m->ell_col,   \
                                                               m->ell_val, m->x, m->y);                \
        break;
    switch (w) {
        SPMV_ELL_CASE(1) SPMV_ELL_CASE(2) SPMV_ELL_CASE(3) SPMV_ELL_CASE(4) SPMV_ELL_CASE(5)
        SPMV_ELL_CASE(6) SPMV_ELL_CASE(7) SPMV_ELL_CASE(8) SPMV_ELL_CASE(9)
    default:
        ell_kernel<R, 0, SKIP><<<grid, block, 0, m->stream>>>(m->rows, m->ell_pitch, w, m->ell_col, m->ell_val,
                                                             m->x, m->y);
    }
#undef SPMV_ELL_CASE
    SPMV_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_ell(Matrix * m, bool)
{
    if (m->rows == 0 || m->ell_w == 0) return 0;
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

// =============================================================================================
// COO
// =============================================================================================

// Row-sorted entries: segmented reduction.  Tiles of kCooTile entries are staged like the CSR
// tiles; phase 1 forms the products with a strided (coalesced, conflict-free) pass; in phase 2
// every thread walks kCooItems CONSECUTIVE entries and emits one fp64 reduction per run of equal
// row indices.  Work per thread is constant whatever the row-length distribution is.
template <int STAGES>
__global__ void __launch_bounds__(kCooThreads)
coo_segmented_kernel(int64_t n, int64_t ntiles, const int32_t * __restrict__ row, const int32_t * __restrict__ col,
                     const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y)
{
    constexpr int T = kCooThreads, TILE = kCooTile, ITEMS = kCooItems;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double * sval = reinterpret_cast<double *>(smem_raw);
    int32_t * srow = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * TILE * 8);
    int32_t * scol = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * TILE * 12);
    uint64_t * full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TILE * 16);
    const int tid = threadIdx.x;
    uint64_t policy = 0;

    auto issue = [&](int s, int64_t t) {
        mbar_arrive_expect_tx(&full[s], TILE * 16);
        bulk_g2s(sval + (size_t)s * TILE, val + t * TILE, TILE * 8, &full[s], policy);
        bulk_g2s(srow + (size_t)s * TILE, row + t * TILE, TILE * 4, &full[s], policy);
        bulk_g2s(scol + (size_t)s * TILE, col + t * TILE, TILE * 4, &full[s], policy);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        policy = policy_evict_first();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (t < ntiles) issue(s, t);
        }
    }

    int64_t it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = (int)(it % STAGES);
        const uint32_t parity = (uint32_t)((it / STAGES) & 1);
        mbar_wait(&full[s], parity);
        double * pv = sval + (size_t)s * TILE;
        const int32_t * pr = srow + (size_t)s * TILE;
        const int32_t * pc = scol + (size_t)s * TILE;
        {
            double a[ITEMS], xv[ITEMS];
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                a[i] = pv[tid + i * T];
                xv[i] = __ldg(x + pc[tid + i * T]);
            }
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) pv[tid + i * T] = __dmul_rn(a[i], xv[i]);
        }
        __syncthreads();

        const int c0 = tid * ITEMS;
        const int64_t k0 = t * TILE + c0;
        int rprev = -1;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (k0 + j < n) {
                const int r = pr[c0 + j];
                if (r != rprev) {
                    if (rprev >= 0) red_add_f64(y + rprev, acc);
                    acc = 0.0;
                    rprev = r;
                }
                acc = __dadd_rn(acc, pv[c0 + j]);
            }
        }
        if (rprev >= 0) red_add_f64(y + rprev, acc);
        __syncthreads();

        if (tid == 0) {
            const int64_t tn = t + (int64_t)STAGES * gridDim.x;
            if (tn < ntiles) {
                fence_proxy_async();
                issue(s, tn);
            }
        }
    }
}

// Entries in file order: one fp64 reduction per entry, two entries per thread and iteration
// (64/128-bit loads; the arrays are padded so the vector loads stay in bounds).
__global__ void __launch_bounds__(256)
coo_atomic_kernel(int64_t n, const int32_t * __restrict__ row, const int32_t * __restrict__ col,
                  const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    const uint64_t pol = policy_evict_first();
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; k < n; k += stride) {
        const int2 r = ldg_stream_i2(row + k, pol);
        const int2 c = ldg_stream_i2(col + k, pol);
        const double2 a = ldg_stream_d2(val + k, pol);
        red_add_f64(y + r.x, __dmul_rn(a.x, __ldg(x + c.x)));
        if (k + 1 < n) red_add_f64(y + r.y, __dmul_rn(a.y, __ldg(x + c.y)));
    }
}

template <int STAGES>
static int launch_coo_seg(Matrix * m, int ctas_per_sm)
{
    auto kernel = coo_segmented_kernel<STAGES>;
    const size_t smem = (size_t)STAGES * kCooTile * 16 + 8 * STAGES + 16;
    static bool configured = false;
    if (!configured) {
        SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int64_t ntiles = (m->coo_n + kCooTile - 1) / kCooTile;
    int64_t grid = std::min<int64_t>(ntiles, (int64_t)m->sm_count * ctas_per_sm);
    if (grid < 1) return 0;
    kernel<<<(unsigned)grid, kCooThreads, smem, m->stream>>>(m->coo_n, ntiles, m->coo_row, m->coo_col, m->coo_val,
                                                            m->x, m->y);
    SPMV_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_coo(Matrix * m)
{
    if (m->coo_n == 0 || m->rows == 0) return 0;
    if (m->coo_mode == SPMVB200_COO_ATOMIC || !m->coo_sorted) {
        m->kernel_name = "coo_atomic_kernel";
        const int64_t pairs = (m->coo_n + 1) / 2;
        int64_t grid = std::min<int64_t>((pairs + 255) / 256, (int64_t)m->sm_count * 8 * 4);
        coo_atomic_kernel<<<(unsigned)grid, 256, 0, m->stream>>>(m->coo_n, m->coo_row, m->coo_col, m->coo_val, m->x,
                                                                m->y);
        SPMV_CUDA(cudaGetLastError());
        count_launch();
        return 0;
    }
    m->kernel_name = "coo_segmented_kernel";
    int stages = (int)(m->opt_coo_stages ? m->opt_coo_stages : 3);
    int ctas = (int)(m->opt_coo_ctas ? m->opt_coo_ctas : 2);
    if (stages == 1) return launch_coo_seg<1>(m, ctas);
    if (stages == 2) return launch_coo_seg<2>(m, ctas);
    if (stages == 3) return launch_coo_seg<3>(m, ctas);
    if (stages == 4) return launch_coo_seg<4>(m, ctas);
    return fail(SPMVB200_ERR_INVALID, "coo.stages must be 1..4");
}

}  // namespace spmvb200
