// kernels_coo.cu -- COO SpMV (see kernels_csr.cu for the design notes shared by all kernels).
//
// Replaces coo_spmv (reference matrix/coo-matrix.cpp:248-285), coo_spmv_atomic (:287-309) and the COO
// tail of the hybrid format (matrix/hybrid-matrix.cpp:491-528).
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// =============================================================================================
// COO
// =============================================================================================

// Row-sorted entries: segmented reduction.  Tiles of kCooTile entries are staged like the CSR
// tiles; phase 1 forms the products with a strided (coalesced, conflict-free) pass; in phase 2
// every thread walks kCooItems CONSECUTIVE entries and emits one fp64 reduction per run of equal
// row indices.  Work per thread is constant whatever the row-length distribution is.
template <int STAGES>
__global__ void __launch_bounds__(kCooThreads)
coo_segmented_kernel(int64_t n, int64_t ntiles, int independent, const int32_t * __restrict__ row, const int32_t * __restrict__ col,
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

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

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
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");

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
coo_atomic_kernel(int64_t n, int independent, const int32_t * __restrict__ row, const int32_t * __restrict__ col,
                  const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    const uint64_t pol = policy_evict_first();
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; k < n; k += stride) {
        const int2 r = ldg_stream_i2(row + k, pol);
        const int2 c = ldg_stream_i2(col + k, pol);
        const double2 a = ldg_stream_d2(val + k, pol);
        red_add_f64(y + r.x, __dmul_rn(a.x, __ldg(x + c.x)));
        if (k + 1 < n) red_add_f64(y + r.y, __dmul_rn(a.y, __ldg(x + c.y)));
    }
}

template <int STAGES>
static int launch_coo_seg(Matrix * m)
{
    auto kernel = coo_segmented_kernel<STAGES>;
    const size_t smem = (size_t)STAGES * kCooTile * 16 + 8 * STAGES + 16;
    static int occupancy = 0;
    if (!occupancy) {
        SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, kCooThreads, smem));
        if (occupancy < 1) return fail(SPMVB200_ERR_CUDA, "coo_segmented_kernel does not fit on an SM");
    }
    const int ctas = m->opt_coo_ctas ? (int)std::min<int64_t>(m->opt_coo_ctas, occupancy) : occupancy;
    const int64_t ntiles = (m->coo_n + kCooTile - 1) / kCooTile;
    int64_t grid = std::min<int64_t>(ntiles, (int64_t)m->sm_count * ctas);
    if (grid < 1) return 0;
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, kCooThreads, smem, m->stream, m->opt_pdl != 0, m->coo_n, ntiles, (int)(m->opt_independent != 0),
                            (const int32_t *)m->coo_row, (const int32_t *)m->coo_col, (const double *)m->coo_val,
                            (const double *)m->x, m->y));
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
        SPMV_CUDA(launch_kernel(coo_atomic_kernel, (unsigned)grid, 256u, 0, m->stream, m->opt_pdl != 0, m->coo_n, (int)(m->opt_independent != 0),
                                (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
                                (const double *)m->coo_val, (const double *)m->x, m->y));
        count_launch();
        return 0;
    }
    m->kernel_name = "coo_segmented_kernel";
    const int stages = (int)(m->opt_coo_stages ? m->opt_coo_stages : 2);
    if (stages == 1) return launch_coo_seg<1>(m);
    if (stages == 2) return launch_coo_seg<2>(m);
    if (stages == 3) return launch_coo_seg<3>(m);
    if (stages == 4) return launch_coo_seg<4>(m);
    return fail(SPMVB200_ERR_INVALID, "coo.stages must be 1..4");
}

}  // namespace spmvb200
