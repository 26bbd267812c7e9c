// kernels_coo.cu -- COO SpMV (see kernels_csr.cu for the design notes shared by all kernels).
//
// Replaces coo_spmv (reference matrix/coo-matrix.cpp:248-285), coo_spmv_atomic (:287-309) and the COO
// tail of the hybrid format (matrix/hybrid-matrix.cpp:491-528).  The reference's T x rows workspace
// (one private y per thread, then a reduce) has no counterpart: partial sums go to y with fp64
// reductions in L2.
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "segreduce.cuh"

#include <algorithm>
#include <atomic>
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
                xv[u] = k < cnt ? ldx(x + pc[k]) : 0.0;
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

// Row-sorted entries, register-staged: no shared memory, no CTA barrier.  A warp owns 32*U consecutive
// entries; lane l loads entries l, l+32, ... (three coalesced streaming loads per stripe), so all 3*U
// matrix loads and then all U gathers of x are in flight at once, and with nothing but registers in
// use 48-64 warps are resident per SM.  This is what a gather-latency-bound matrix (power law: the
// x-gathers miss L1 and half of them miss L2) needs: the staged kernel above keeps at most
// 3 CTAs x 256 threads resident and alternates load, gather and reduce phases behind CTA barriers.
// Each 32-entry stripe is reduced by a segmented warp scan keyed on the runs of equal row index
// (Kogge-Stone with shuffles, cut short at the longest run of the stripe); the last lane of every
// run adds its total to y with one fp64 reduction.  Correct for any entry order; efficient when equal
// rows are adjacent.
template <int U, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
coo_warp_kernel(int64_t n, int independent, const int32_t * __restrict__ row, const int32_t * __restrict__ col,
                const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int64_t k0 = ((int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * (32 * U) + lane;
    if (k0 - lane >= n) return;
    const uint64_t pol = policy_evict_first();
    int r[U], c[U];
    double a[U];
    // the arrays are padded with zeroed entries far beyond n: whole stripes can be loaded unguarded
#pragma unroll
    for (int u = 0; u < U; ++u) {
        r[u] = ldg_stream_i1(row + k0 + 32 * u, pol);
        c[u] = ldg_stream_i1(col + k0 + 32 * u, pol);
        a[u] = ldg_stream_d1(val + k0 + 32 * u, pol);
    }
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldx(x + c[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int rr = (k0 + 32 * u < n) ? r[u] : -1;
        double s = __dmul_rn(a[u], v[u]);
        const int rp = __shfl_up_sync(0xffffffffu, rr, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || rr != rp);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // first lane of my run
        const int dist = lane - start;
        const int longest = __reduce_max_sync(0xffffffffu, dist);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            if (d <= longest) {  // warp-uniform
                const double t = __shfl_up_sync(0xffffffffu, s, d);
                if (dist >= d) s = __dadd_rn(s, t);
            }
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (tail && rr >= 0) red_add_f64(y + rr, s);
    }
}

// The same idea with a BLOCKED assignment: lane l owns the 4 consecutive entries 4l..4l+3 of the warp's
// 128, fetched with one 128-bit load each for the row and column indices and one 256-bit load for
// the values (3 load instructions instead of 12).  The runs of equal rows are summed by
// warp_segmented_add4 (segreduce.cuh): serially inside a lane, ONE segmented scan across the warp
// per 128 entries -- a quarter of the shuffles of the striped kernel, whose l1tex/MIO pipe was 70 %
// busy with gathers + shuffles (profiles/r01_ncu_c3_coo_warp.txt).
template <int WARPS, int XPATH = 0>
__global__ void __launch_bounds__(WARPS * 32)
coo_warp4_kernel(int64_t n, int independent, const int32_t * __restrict__ row, const int32_t * __restrict__ col,
                 const double * __restrict__ val, const double * __restrict__ x, double * __restrict__ y, double alpha,
                 cudaTextureObject_t xtex)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int64_t kw = ((int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * 128;
    if (kw >= n) return;
    const int64_t k0 = kw + 4 * lane;
    const uint64_t pol = policy_evict_first();
    const int4 r4 = ldg_stream_i4(row + k0, pol);
    const int4 c4 = ldg_stream_i4(col + k0, pol);
    double a[4];
    ldg_stream_d4(val + k0, a);
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
    double x0, x1, x2, x3;
    if (XPATH == 4) {
        // experiment: gather through cp.async (LDGSTS.BYPASS): the 16 B pair holding x[c] goes L2 -> shared memory without
        // passing through an L1 line.  (x must be 16 B aligned.)
        __shared__ __align__(16) double2 stage[WARPS * 128];
        double2 * mine = stage + (threadIdx.x >> 5) * 128 + 4 * lane;
        const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(mine + j)), "l"(x + (cc[j] & ~1)) : "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
        x0 = (cc[0] & 1) ? mine[0].y : mine[0].x;
        x1 = (cc[1] & 1) ? mine[1].y : mine[1].x;
        x2 = (cc[2] & 1) ? mine[2].y : mine[2].x;
        x3 = (cc[3] & 1) ? mine[3].y : mine[3].x;
    } else if (XPATH == 5) {
        // experiment: gather through the texture pipe (TEX.LL / TLD in SASS): x as a linear texture of int2 texels.  The
        // texture path has its own input stage in L1TEX; does it also have its own divergent-address rate?
        const int2 t0 = tex1Dfetch<int2>(xtex, c4.x), t1 = tex1Dfetch<int2>(xtex, c4.y);
        const int2 t2 = tex1Dfetch<int2>(xtex, c4.z), t3 = tex1Dfetch<int2>(xtex, c4.w);
        x0 = __hiloint2double(t0.y, t0.x); x1 = __hiloint2double(t1.y, t1.x);
        x2 = __hiloint2double(t2.y, t2.x); x3 = __hiloint2double(t3.y, t3.x);
    } else if (XPATH == 6) {
        // experiment: half of the gathers through the texture pipe, half through the LSU
        const int2 t0 = tex1Dfetch<int2>(xtex, c4.x), t2 = tex1Dfetch<int2>(xtex, c4.z);
        x1 = __ldg(x + c4.y); x3 = __ldg(x + c4.w);
        x0 = __hiloint2double(t0.y, t0.x); x2 = __hiloint2double(t2.y, t2.x);
    } else {
        x0 = ld_x<XPATH>(x + c4.x); x1 = ld_x<XPATH>(x + c4.y); x2 = ld_x<XPATH>(x + c4.z); x3 = ld_x<XPATH>(x + c4.w);
    }
    const int r[4] = {k0 < n ? r4.x : -1, k0 + 1 < n ? r4.y : -1, k0 + 2 < n ? r4.z : -1, k0 + 3 < n ? r4.w : -1};
    const double p[4] = {__dmul_rn(a[0], x0), __dmul_rn(a[1], x1), __dmul_rn(a[2], x2), __dmul_rn(a[3], x3)};
    warp_segmented_add4(lane, r, p, y, alpha);
}

// Scattered gathers (power-law matrices): the hot-column kernel.  The builder (coo_build_hot, builders.cu) has cut the
// entries into segments and given every segment a table of its H most referenced columns; colh = column_index with
// those columns replaced by 0x80000000 | slot.  A persistent grid (CTAs = what fits the SMs with H*8 bytes of shared
// memory each) takes the segments round-robin.  Per segment the CTA loads the x values of the table into shared memory
// (one coalesced read of the column list, H gathers) and then runs the blocked segmented reduction of coo_warp4_kernel
// over the segment's spans of 32*E entries, warps taking spans round-robin; a gather is an LDS for a hot column and a
// global load otherwise.  More than half of the gathers of an R-MAT matrix are hot (24 576 slots), which removes their
// L2 -> L1 sector traffic (32 B moved for 8 B used) and most of their L1 wavefronts -- the two pipes the plain kernel
// saturates (profiles/r01_ncu_c3_coo_final.txt: l1tex 86 %, lts 78 %, DRAM 57 %).
// MEASURED: slower than the plain kernel for every table size (profiles/r02_sweep_q_coo_hot_columns.log: 1.11 ms with
// 8 192 slots ... 1.80 ms with 24 576, against 1.04 ms).  The divergent gather is bound by the misses the L1 can hold in
// flight, and the tables take that capacity away (profiles/r02_sweep_r_coo_gather_paths.log).  Kept as an opt-in
// ("coo.hot" = 1) because it is the documented negative result, with its tests.
template <int THREADS, int E>
__global__ void __launch_bounds__(THREADS, 1)
coo_hot_kernel(int nseg, int h, const int64_t * __restrict__ seg, const int32_t * __restrict__ hot_cols,
               const int32_t * __restrict__ row, const int32_t * __restrict__ colh, const double * __restrict__ val,
               const double * __restrict__ x, double * __restrict__ y, double alpha)
{
    extern __shared__ __align__(16) double xs[];
    constexpr int SPAN = 32 * E, WARPS = THREADS / 32;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t pol = policy_evict_first();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the table is made of x values
    for (int q = blockIdx.x; q < nseg; q += gridDim.x) {
        const int64_t lo = seg[q], hi = seg[q + 1];
        __syncthreads();  // everybody is done with the previous segment's table
        const int32_t * table = hot_cols + (int64_t)q * h;
        for (int i = threadIdx.x; i < h; i += THREADS) xs[i] = ldx(x + __ldg(table + i));
        __syncthreads();
        const int64_t first = lo & ~(int64_t)(SPAN - 1);
        for (int64_t kw = first + (int64_t)warp * SPAN; kw < hi; kw += (int64_t)WARPS * SPAN) {
            const int64_t k0 = kw + E * lane;
            int r[E], c[E];
            double a[E];
#pragma unroll
            for (int g = 0; g < E / 4; ++g) {
                const int4 r4 = ldg_stream_i4(row + k0 + 4 * g, pol);
                const int4 c4 = ldg_stream_i4(colh + k0 + 4 * g, pol);
                double a4[4];
                ldg_stream_d4(val + k0 + 4 * g, a4);
                r[4 * g] = r4.x; r[4 * g + 1] = r4.y; r[4 * g + 2] = r4.z; r[4 * g + 3] = r4.w;
                c[4 * g] = c4.x; c[4 * g + 1] = c4.y; c[4 * g + 2] = c4.z; c[4 * g + 3] = c4.w;
                a[4 * g] = a4[0]; a[4 * g + 1] = a4[1]; a[4 * g + 2] = a4[2]; a[4 * g + 3] = a4[3];
            }
            double p[E];
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const bool inside = k0 + j >= lo && k0 + j < hi;  // spans are aligned; a segment may start or end inside one
                if (!inside) { r[j] = -1; c[j] = 0; }
            }
#pragma unroll
            for (int j = 0; j < E; ++j) p[j] = c[j] < 0 ? xs[c[j] & 0x7fffffff] : ldx(x + c[j]);
#pragma unroll
            for (int j = 0; j < E; ++j) p[j] = __dmul_rn(a[j], p[j]);
            warp_segmented_add<E>(lane, r, p, y, alpha);
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
        red_add_f64(y + r.x, __dmul_rn(a.x, ldx(x + c.x)));
        if (k + 1 < n) red_add_f64(y + r.y, __dmul_rn(a.y, ldx(x + c.y)));
    }
}

template <int THREADS, int STAGES>
static int launch_coo_seg(Matrix * m)
{
    auto kernel = coo_segmented_kernel<THREADS, STAGES>;
    constexpr size_t smem = (size_t)STAGES * THREADS * kCooItems * 16 + 8 * STAGES + 16;
    // per device (the attribute is a property of the function ON a device) and safe to race: the worst case is that
    // two threads set the same attribute and compute the same number
    static std::atomic<int> occupancy_of[64];
    const int dev = m->device >= 0 && m->device < 64 ? m->device : 0;
    int occupancy = occupancy_of[dev].load(std::memory_order_acquire);
    if (!occupancy) {
        SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, THREADS, smem));
        if (occupancy < 1) return fail(SPMVB200_ERR_CUDA, "coo_segmented_kernel does not fit on an SM");
        occupancy_of[dev].store(occupancy, std::memory_order_release);
    }
    const int ctas = m->opt_coo_ctas ? (int)std::min<int64_t>(m->opt_coo_ctas, occupancy) : occupancy;
    int64_t grid = std::min<int64_t>((int64_t)m->sm_count * ctas, std::max<int64_t>(1, (m->coo_n + 15) / 16));
    const int64_t chunk = round_up((m->coo_n + grid - 1) / grid, 16);
    const RunMode rm = run_mode(m);
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, (unsigned)THREADS, smem, m->stream, rm.pdl, m->coo_n, chunk,
                            rm.independent, (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
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

template <int U, int WARPS>
static int launch_coo_warp_variant(Matrix * m)
{
    const int64_t per_cta = (int64_t)WARPS * 32 * U;
    const int64_t grid = (m->coo_n + per_cta - 1) / per_cta;
    if (grid > INT_MAX) return fail(SPMVB200_ERR_OVERFLOW, "COO matrix too large for one launch");
    const RunMode rm = run_mode(m);
    SPMV_CUDA(launch_kernel(coo_warp_kernel<U, WARPS>, (unsigned)grid, (unsigned)(WARPS * 32), 0, m->stream, rm.pdl, m->coo_n,
                            rm.independent, (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
                            (const double *)m->coo_val, (const double *)m->x, m->y));
    count_launch();
    return 0;
}

// (experiment, "coo.xload" = 5 / 6) x as a linear texture of 8-byte texels, rebuilt when x was rebound
static int x_texture(Matrix * m)
{
    if (m->x_tex && m->x_tex_ptr == m->x) return 0;
    if (m->x_tex) cudaDestroyTextureObject(m->x_tex);
    m->x_tex = 0;
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeLinear;
    res.res.linear.devPtr = m->x;
    res.res.linear.desc = cudaCreateChannelDesc<int2>();
    res.res.linear.sizeInBytes = sizeof(double) * (size_t)m->cols;
    cudaTextureDesc tex = {};
    tex.readMode = cudaReadModeElementType;
    SPMV_CUDA(cudaCreateTextureObject(&m->x_tex, &res, &tex, nullptr));
    m->x_tex_ptr = m->x;
    return 0;
}

template <int WARPS>
static int launch_coo_warp4_variant(Matrix * m)
{
    const int64_t per_cta = (int64_t)WARPS * 128;
    const int64_t grid = (m->coo_n + per_cta - 1) / per_cta;
    if (grid > INT_MAX) return fail(SPMVB200_ERR_OVERFLOW, "COO matrix too large for one launch");
    const RunMode rm = run_mode(m);
    auto kernel = coo_warp4_kernel<WARPS, 0>;
    if (WARPS == 4) {  // experiment switches: cache path of the x gather, L1 / shared-memory carve-out
        if (m->opt_coo_xload == 1) kernel = coo_warp4_kernel<WARPS == 4 ? 4 : WARPS, WARPS == 4 ? 1 : 0>;
        else if (m->opt_coo_xload == 2) kernel = coo_warp4_kernel<WARPS == 4 ? 4 : WARPS, WARPS == 4 ? 2 : 0>;
        else if (m->opt_coo_xload == 3) kernel = coo_warp4_kernel<WARPS == 4 ? 4 : WARPS, WARPS == 4 ? 3 : 0>;
        else if (m->opt_coo_xload == 4) kernel = coo_warp4_kernel<WARPS == 4 ? 4 : WARPS, WARPS == 4 ? 4 : 0>;
        else if (m->opt_coo_xload == 5) kernel = coo_warp4_kernel<WARPS == 4 ? 4 : WARPS, WARPS == 4 ? 5 : 0>;
        else if (m->opt_coo_xload == 6) kernel = coo_warp4_kernel<WARPS == 4 ? 4 : WARPS, WARPS == 4 ? 6 : 0>;
        if (m->opt_coo_xload == 5 || m->opt_coo_xload == 6) SPMV_TRY(x_texture(m));
        if (m->opt_coo_carveout >= 0)
            SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)m->opt_coo_carveout));
    }
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, (unsigned)(WARPS * 32), 0, m->stream, rm.pdl, m->coo_n,
                            rm.independent, (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
                            (const double *)m->coo_val, (const double *)m->x, m->y, m->alpha, m->x_tex));
    count_launch();
    return 0;
}

static int launch_coo_warp(Matrix * m)
{
    const int threads = (int)(m->opt_coo_threads ? m->opt_coo_threads : 128);  // sweep: profiles/r01_sweep_i_coo_warp4.log
    if (m->opt_coo_algo != 4) {  // blocked lanes
        if (threads == 64) return launch_coo_warp4_variant<2>(m);
        if (threads == 128) return launch_coo_warp4_variant<4>(m);
        if (threads == 256) return launch_coo_warp4_variant<8>(m);
        return fail(SPMVB200_ERR_INVALID, "coo.threads must be 64, 128 or 256");
    }
    const int items = (int)(m->opt_coo_items ? m->opt_coo_items : 4);
#define SPMV_COO_WARP(U)                                                  \
    case U:                                                               \
        if (threads == 64) return launch_coo_warp_variant<U, 2>(m);       \
        if (threads == 128) return launch_coo_warp_variant<U, 4>(m);      \
        if (threads == 256) return launch_coo_warp_variant<U, 8>(m);      \
        break;
    switch (items) {
        SPMV_COO_WARP(2) SPMV_COO_WARP(4) SPMV_COO_WARP(8)
    default: return fail(SPMVB200_ERR_INVALID, "coo.items must be 2, 4 or 8");
    }
#undef SPMV_COO_WARP
    return fail(SPMVB200_ERR_INVALID, "coo.threads must be 64, 128 or 256");
}

template <int THREADS, int E>
static int launch_coo_hot_variant(Matrix * m)
{
    auto kernel = coo_hot_kernel<THREADS, E>;
    const size_t smem = (size_t)m->coo_hot_h * 8;
    SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // (per device, cheap)
    int occupancy = 0;
    SPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, THREADS, smem));
    if (occupancy < 1) return fail(SPMVB200_ERR_CUDA, "coo_hot_kernel does not fit on an SM");
    const int64_t grid = std::min<int64_t>((int64_t)m->sm_count * occupancy, m->coo_nseg);
    const RunMode rm = run_mode(m);
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, (unsigned)THREADS, smem, m->stream, rm.pdl, m->coo_nseg, m->coo_hot_h,
                            (const int64_t *)m->coo_seg, (const int32_t *)m->coo_hot_cols, (const int32_t *)m->coo_row,
                            (const int32_t *)m->coo_colh, (const double *)m->coo_val, (const double *)m->x, m->y, m->alpha));
    count_launch();
    return 0;
}

static int launch_coo_hot(Matrix * m)
{
    const int threads = (int)(m->opt_coo_hot_threads ? m->opt_coo_hot_threads : 1024);
    const int entries = (int)(m->opt_coo_hot_entries ? m->opt_coo_hot_entries : 4);
    m->kernel_name = "coo_hot_kernel";
    if (threads == 1024 && entries == 4) return launch_coo_hot_variant<1024, 4>(m);
    if (threads == 1024 && entries == 8) return launch_coo_hot_variant<1024, 8>(m);
    if (threads == 512 && entries == 4) return launch_coo_hot_variant<512, 4>(m);
    if (threads == 512 && entries == 8) return launch_coo_hot_variant<512, 8>(m);
    if (threads == 256 && entries == 4) return launch_coo_hot_variant<256, 4>(m);
    if (threads == 256 && entries == 8) return launch_coo_hot_variant<256, 8>(m);
    return fail(SPMVB200_ERR_INVALID, "coo.hot_threads must be 256, 512 or 1024 and coo.hot_entries 4 or 8");
}

int launch_coo(Matrix * m)
{
    if (m->rows == 0) return 0;
    if (m->coo_n > 0 && m->opt_coo_algo == 0 && m->opt_coo_hot > 0 && !m->coo_hot_tried) SPMV_TRY(coo_build_hot(m));
    if (m->dry_run) return 0;
    SPMV_TRY(clear_y_for_beta0(m));  // every COO kernel adds partial sums
    if (m->coo_n == 0) return 0;
    if (m->opt_coo_algo == 0 && m->coo_colh && m->opt_coo_hot > 0) return launch_coo_hot(m);
    if (m->opt_coo_algo == 1 || m->opt_coo_algo == 3 || m->opt_coo_algo == 4) SPMV_TRY(need_unit_alpha(m, "this COO kernel"));
    // coo.algo: 0 automatic, 1 shared-memory staged tiles (sorted entries only), 2 register-staged warp
    // stripes (any order), 3 one reduction per entry
    // Automatic = 2 for both modes: on file-order entries the warp kernel still merges adjacent equal
    // rows and keeps 4 gathers per lane in flight (R-MAT 2^24: 1.55 ms vs 5.07 ms for algo 3).
    const bool unsorted = m->coo_mode == SPMVB200_COO_ATOMIC || !m->coo_sorted;
    if (m->opt_coo_algo == 3 || (unsorted && m->opt_coo_algo == 1)) {
        m->kernel_name = "coo_atomic_kernel";
        const int64_t pairs = (m->coo_n + 1) / 2;
        int64_t grid = std::min<int64_t>((pairs + 255) / 256, (int64_t)m->sm_count * 8 * 4);
        const RunMode rm = run_mode(m);
        SPMV_CUDA(launch_kernel(coo_atomic_kernel, (unsigned)grid, 256u, 0, m->stream, rm.pdl, m->coo_n,
                                rm.independent, (const int32_t *)m->coo_row, (const int32_t *)m->coo_col,
                                (const double *)m->coo_val, (const double *)m->x, m->y));
        count_launch();
        return 0;
    }
    if (m->opt_coo_algo != 1) {
        m->kernel_name = m->opt_coo_algo == 4 ? "coo_warp_kernel" : "coo_warp4_kernel";
        return launch_coo_warp(m);
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
