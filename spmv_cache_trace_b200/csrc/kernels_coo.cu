// kernels_coo.cu -- COO SpMV (see kernels_csr.cu for the design notes shared by all kernels).
//
// Replaces coo_spmv (reference matrix/coo-matrix.cpp:248-285), coo_spmv_atomic (:287-309) and the COO
// tail of the hybrid format (matrix/hybrid-matrix.cpp:491-528).  The reference's T x rows workspace
// (one private y per thread, then a reduce) has no counterpart: partial sums go to y with fp64
// reductions in L2.
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// Row-sorted entries: segmented reduction.  Like the CSR stream kernel, the entries are cut into
// equal contiguous chunks, one per CTA of a persistent grid, walked in tiles of THREADS*7 entries
// that are staged in shared memory by bulk-async copies (row indices, column indices, values:
// 16 B per entry).  Phase 1 forms the products with a strided (coalesced, conflict-free) pass; in
// phase 2 every thread walks 7 CONSECUTIVE entries (7 is odd: the blocked walk is bank-conflict
// free) and emits one fp64 reduction per run of equal row indices.  Work per thread is constant
// whatever the row-length distribution is -- the point of COO for power-law matrices.
template <int THREADS, int STAGES>
__global__ void __launch_bounds__(THREADS)
coo_segmented_kernel(int64_t n, int64_t chunk, int independent, const int32_t * __restrict__ row,
                     const int32_t * __restrict__ col, const double * __restrict__ val,
                     const double * __restrict__ x, double * __restrict__ y)
{
    constexpr int T = THREADS, ITEMS = kCooItems, TILE = THREADS * kCooItems;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double * sval = reinterpret_cast<double *>(smem_raw);
    int32_t * srow = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * TILE * 8);
    int32_t * scol = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * TILE * 12);
    uint64_t * full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TILE * 16);
    const int tid = threadIdx.x;
    const int64_t cbegin = (int64_t)blockIdx.x * chunk;
    const int64_t cend = min(cbegin + chunk, n);
    const int ntiles = cend > cbegin ? (int)((cend - cbegin + TILE - 1) / TILE) : 0;
    uint64_t policy = 0;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    auto issue = [&](int s, int i) {
        const int64_t tb = cbegin + (int64_t)i * TILE;
        const uint32_t cnt = (uint32_t)((min(tb + (int64_t)TILE, cend) - tb + 15) & ~(int64_t)15);
        mbar_arrive_expect_tx(&full[s], cnt * 16u);
        bulk_g2s(sval + (size_t)s * TILE, val + tb, cnt * 8u, &full[s], policy);
        bulk_g2s(srow + (size_t)s * TILE, row + tb, cnt * 4u, &full[s], policy);
        bulk_g2s(scol + (size_t)s * TILE, col + tb, cnt * 4u, &full[s], policy);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        policy = policy_evict_first();
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (s < ntiles) issue(s, s);
    }
    __syncthreads();
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");

    for (int i = 0; i < ntiles; ++i) {
        const int s = i % STAGES;
        const uint32_t parity = (uint32_t)((i / STAGES) & 1);
        const int64_t base = cbegin + (int64_t)i * TILE;
        const int cnt = (int)(min(base + (int64_t)TILE, cend) - base);
        mbar_wait(&full[s], parity);
        double * pv = sval + (size_t)s * TILE;
        const int32_t * pr = srow + (size_t)s * TILE;
        const int32_t * pc = scol + (size_t)s * TILE;
        {
            double a[ITEMS], xv[ITEMS];
#pragma unroll
            for (int u = 0; u < ITEMS; ++u) {
                const int k = tid + u * T;
                a[u] = k < cnt ? pv[k] : 0.0;
                xv[u] = k < cnt ? __ldg(x + pc[k]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < ITEMS; ++u)
                if (tid + u * T < cnt) pv[tid + u * T] = __dmul_rn(a[u], xv[u]);
        }
        __syncthreads();

        const int c0 = tid * ITEMS;
        int rprev = -1;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (c0 + j < cnt) {
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

        if (tid == 0 && i + STAGES < ntiles) {
            fence_proxy_async();
            issue(s, i + STAGES);
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

template <int THREADS, int STAGES>
static int launch_coo_seg(Matrix * m)
{
    auto kernel = coo_segmented_kernel<THREADS, STAGES>;
    constexpr size_t smem = (size_t)STAGES * THREADS * kCooItems * 16 + 8 * STAGES + 16;
    static int occupancy = 0;
    if (!occupancy) {
        SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, THREADS, smem));
        if (occupancy < 1) return fail(SPMVB200_ERR_CUDA, "coo_segmented_kernel does not fit on an SM");
    }
    const int ctas = m->opt_coo_ctas ? (int)std::min<int64_t>(m->opt_coo_ctas, occupancy) : occupancy;
    int64_t grid = std::min<int64_t>((int64_t)m->sm_count * ctas, std::max<int64_t>(1, (m->coo_n + 15) / 16));
    const int64_t chunk = round_up((m->coo_n + grid - 1) / grid, 16);
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, (unsigned)THREADS, smem, m->stream, m->opt_pdl != 0, m->coo_n, chunk,
                            (int)(m->opt_independent != 0), (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
                            (const double *)m->coo_val, (const double *)m->x, m->y));
    count_launch();
    return 0;
}

template <int THREADS>
static int launch_coo_stages(Matrix * m, int stages)
{
    switch (stages) {
    case 2: return launch_coo_seg<THREADS, 2>(m);
    case 3: return launch_coo_seg<THREADS, 3>(m);
    case 4: return launch_coo_seg<THREADS, 4>(m);
    }
    return fail(SPMVB200_ERR_INVALID, "coo.stages must be 2, 3 or 4");
}

int launch_coo(Matrix * m)
{
    if (m->coo_n == 0 || m->rows == 0) return 0;
    if (m->coo_mode == SPMVB200_COO_ATOMIC || !m->coo_sorted) {
        m->kernel_name = "coo_atomic_kernel";
        const int64_t pairs = (m->coo_n + 1) / 2;
        int64_t grid = std::min<int64_t>((pairs + 255) / 256, (int64_t)m->sm_count * 8 * 4);
        SPMV_CUDA(launch_kernel(coo_atomic_kernel, (unsigned)grid, 256u, 0, m->stream, m->opt_pdl != 0, m->coo_n,
                                (int)(m->opt_independent != 0), (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
                                (const double *)m->coo_val, (const double *)m->x, m->y));
        count_launch();
        return 0;
    }
    m->kernel_name = "coo_segmented_kernel";
    const int stages = (int)(m->opt_coo_stages ? m->opt_coo_stages : 2);
    const int threads = (int)(m->opt_coo_threads ? m->opt_coo_threads : 256);  // sweep: profiles/r01_sweep_f_coo.log
    if (threads == 64) return launch_coo_stages<64>(m, stages);
    if (threads == 128) return launch_coo_stages<128>(m, stages);
    if (threads == 256) return launch_coo_stages<256>(m, stages);
    return fail(SPMVB200_ERR_INVALID, "coo.threads must be 64, 128 or 256");
}

}  // namespace spmvb200
