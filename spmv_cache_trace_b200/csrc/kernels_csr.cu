// kernels_csr.cu -- CSR SpMV for sm_100a:  y += A*x.  Kernel selection (launch_csr) and the shared-memory
// staged "stream" kernel.
//
// Replaces csr_spmv + csr_spmv_inner_loop (reference matrix/csr-matrix-spmv.cpp:21-33, 63-76),
// the body of csr_spmv_kernel::run (kernels/csr-spmv.cpp:64-67).
//
// The default kernels are register-staged and live in kernels_csr_flat.cu (split by non-zeros) and
// kernels_csr_sliced.cu (lane per row, long regular rows); the stream kernel below is the first
// generation, kept as "csr.algo" 1 / 2 and as the path that sums rows inside a tile in the reference's
// order.  Its design notes; the points on reductions, PDL and cache policy hold for every kernel:
//   * SpMV is HBM-bandwidth bound (2 flop per 12 streamed bytes): what matters is keeping enough
//     bytes in flight and touching every matrix byte once.  The non-zeros (not the rows) are cut
//     into equal contiguous chunks, one per CTA of a persistent grid, so every CTA streams the same
//     number of bytes and all CTAs finish together whatever the row lengths are.  A CTA walks its
//     chunk in tiles; each tile's values, column indices AND the window of row pointers it needs
//     are brought into shared memory by 1-D bulk-async copies (the TMA engine: cp.async.bulk +
//     mbarrier complete_tx, UBLKCP in SASS) through a multi-stage ring, so the global loads are
//     perfectly coalesced and need no registers.
//   * "direct" mode: G lanes (1, 2, 4 or 8, chosen from the mean row length) own a row and form the
//     products of its slice of the tile themselves, eight gathers of x in flight per lane.  The 32
//     lanes of a warp hold 32/G consecutive rows, so for banded matrices the gathers of a warp fall
//     into a few cache lines -- a quarter of the L1 wavefronts of a pass that walks the non-zeros in
//     storage order, which is what limits that pass (measured: l1tex 90 % busy, profiles/).
//     "product" mode (long rows on average): a coalesced pass overwrites the staged values with the
//     products a[k]*x[j[k]], then one thread per row adds its slice.
//   * Results are added to y with fp64 reductions in L2 (RED.ADD.F64): one rounding, y_old + sum,
//     exactly like the reference's `y[i] += z`, without a read round trip through the SM.
//   * Launches use programmatic dependent launch: the prologue (barrier set-up, first STAGES tiles of
//     immutable matrix data) runs while the previous kernel drains; x and y are touched only after
//     griddepcontrol.wait.
//   * x is gathered through the read-only L1/L2 path; streamed matrix data carries an evict-first L2
//     policy so it does not push x out of L2.
//
// Arithmetic order: with G = 1 a row that lies inside one tile and is not longer than kLongRow is
// summed left to right with separate multiply and add roundings (__dmul_rn/__dadd_rn), which is
// exactly what the reference's scalar loop does when compiled for baseline x86-64 (no FMA
// contraction): bit-identical results.  Rows cut by a tile boundary, rows shared by G > 1 lanes
// and long rows are combined in a different order, within BASELINE.json's per-row tolerance
// |y - y_ref| <= 1e-12 * sum_j |a_ij x_j|.
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <atomic>
#include <climits>

namespace spmvb200 {

using namespace ptx;

constexpr int kLongRow = 96;  // per lane-group: longer slices are summed by a whole warp

// Row range of every tile of every CTA.  CTA b owns non-zeros [b*chunk, (b+1)*chunk); its i-th tile
// starts at b*chunk + i*tile.  r0 = row holding the tile's first non-zero, r1 = row holding the
// first non-zero after the tile (rows-1 at the very end): the rows a tile touches are r0..r1.
template <typename OffT>
__global__ void csr_tile_table_kernel(int64_t rows, int64_t stored, int64_t chunk, int tpc, int tile, int grid,
                                      const OffT * __restrict__ rp, int2 * __restrict__ table)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)grid * tpc) return;
    const int64_t b = idx / tpc, i = idx - b * tpc;
    const int64_t cend = min((b + 1) * chunk, stored);
    const int64_t tb = b * chunk + i * tile;
    if (tb >= cend) {
        table[idx] = make_int2(0, -1);
        return;
    }
    const int64_t te = min(tb + (int64_t)tile, cend);
    auto row_of = [&](int64_t p) {  // largest r in [0, rows-1] with rp[r] <= p
        int64_t lo = 0, hi = rows - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if ((int64_t)rp[mid] <= p) lo = mid; else hi = mid - 1;
        }
        return (int)lo;
    };
    table[idx] = make_int2(row_of(tb), te < stored ? row_of(te) : (int)(rows - 1));
}

template <int TILE>
__host__ __device__ constexpr int csr_rwin() { return TILE / 4 + 8; }

template <typename OffT, int TILE, int STAGES>
constexpr size_t csr_smem_bytes()
{
    return (size_t)STAGES * (TILE * 12 + csr_rwin<TILE>() * sizeof(OffT) + 16) + 8 * STAGES;
}

// G > 0: direct mode with G lanes per row.  G == 0: product mode.
template <typename OffT, int THREADS, int TILE, int STAGES, int G>
__global__ void __launch_bounds__(THREADS)
csr_stream_kernel(int64_t stored, int64_t chunk, int tpc, int independent, const OffT * __restrict__ rp,
                  const int32_t * __restrict__ col, const double * __restrict__ val,
                  const int2 * __restrict__ table, const double * __restrict__ x, double * __restrict__ y)
{
    constexpr int T = THREADS;
    constexpr int PER = TILE / T;
    constexpr int RWIN = csr_rwin<TILE>();
    constexpr int LANES = G > 0 ? G : 1;  // lanes per row in the row pass
    static_assert(TILE % T == 0 && TILE % 16 == 0, "tile must be a multiple of the block and of 16");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double * sval = reinterpret_cast<double *>(smem_raw);
    int32_t * scol = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * TILE * 8);
    OffT * srp = reinterpret_cast<OffT *>(smem_raw + (size_t)STAGES * TILE * 12);
    int32_t * smeta = reinterpret_cast<int32_t *>(smem_raw + (size_t)STAGES * (TILE * 12 + RWIN * sizeof(OffT)));
    uint64_t * full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * (TILE * 12 + RWIN * sizeof(OffT) + 16));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int64_t cbegin = (int64_t)blockIdx.x * chunk;
    const int64_t cend = min(cbegin + chunk, stored);
    const int2 * my_table = table + (int64_t)blockIdx.x * tpc;
    const int ntiles = cend > cbegin ? (int)((cend - cbegin + TILE - 1) / TILE) : 0;
    uint64_t policy = 0;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Thread 0 is the producer: one barrier arrival + three bulk copies per tile (values, column
    // indices, and the window of row pointers starting at the 16 B aligned row before r0; r0 and r1
    // come from the tile table and are handed to the consumers through shared memory).
    // Tile lengths are multiples of 16 entries (chunk is); all arrays are padded past their end.
    // (A speculative placement of the row-pointer window from the mean row length, which would take
    // the table lookup off the critical path of a CTA's first tile, was tried and dropped: boundary
    // planes of 3-D stencils shift the estimate by thousands of rows.)
    auto issue = [&](int s, int i, int2 rr) {
        const int64_t tb = cbegin + (int64_t)i * TILE;
        const uint32_t n = (uint32_t)((min(tb + (int64_t)TILE, cend) - tb + 15) & ~(int64_t)15);
        smeta[s * 4 + 0] = rr.x;
        smeta[s * 4 + 1] = rr.y;
        mbar_arrive_expect_tx(&full[s], n * 12u + (uint32_t)(RWIN * sizeof(OffT)));
        bulk_g2s(sval + (size_t)s * TILE, val + tb, n * 8u, &full[s], policy);
        bulk_g2s(scol + (size_t)s * TILE, col + tb, n * 4u, &full[s], policy);
        bulk_g2s(srp + (size_t)s * RWIN, rp + (rr.x & ~3), (uint32_t)(RWIN * sizeof(OffT)), &full[s], policy);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        policy = policy_evict_first();
    }
    __syncthreads();
    if (tid == 0) {
        int2 rr[STAGES];
#pragma unroll
        for (int s = 0; s < STAGES; ++s) rr[s] = s < ntiles ? __ldg(my_table + s) : make_int2(0, -1);
#pragma unroll
        for (int s = 0; s < STAGES; ++s)
            if (s < ntiles) issue(s, s, rr[s]);
    }

    // Everything above reads only the immutable matrix; x and y may come from the previous launch.
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");

    for (int i = 0; i < ntiles; ++i) {
        const int s = i % STAGES;
        const uint32_t parity = (uint32_t)((i / STAGES) & 1);
        const int64_t base = cbegin + (int64_t)i * TILE;
        const int64_t end = min(base + (int64_t)TILE, cend);

        // The producer fetches the row range of the tile it will issue at the end of this iteration.
        const int in = i + STAGES;
        int2 nrr = make_int2(0, -1);
        if (tid == 0 && in < ntiles) nrr = __ldg(my_table + in);

        mbar_wait(&full[s], parity);
        double * pv = sval + (size_t)s * TILE;
        const int32_t * pc = scol + (size_t)s * TILE;
        const OffT * wrp = srp + (size_t)s * RWIN;
        const int r0 = smeta[s * 4 + 0];
        const int r1 = smeta[s * 4 + 1];
        const int r0a = r0 & ~3;

        if (G == 0) {
            // Product pass: coalesced over the tile, PER gathers in flight per thread.
            const int n = (int)(end - base);
            double a[PER], xv[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int k = tid + u * T;
                a[u] = k < n ? pv[k] : 0.0;
                xv[u] = k < n ? ldx(x + pc[k]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < PER; ++u)
                if (tid + u * T < n) pv[tid + u * T] = __dmul_rn(a[u], xv[u]);
            __syncthreads();
        }

        // Row pass: LANES lanes per row, rows r0..r1, each row's slice [a, b) of the tile.
        const int g = lane % LANES;
        for (int rw = r0 + (tid - lane) / LANES; rw <= r1; rw += T / LANES) {  // warp-uniform trip count
            const int r = rw + lane / LANES;
            int a = 0, b = 0;
            if (r <= r1) {
                const int idx = r - r0a;
                int64_t lo, hi;
                if (idx + 1 < RWIN) {
                    lo = (int64_t)wrp[idx];
                    hi = (int64_t)wrp[idx + 1];
                } else {  // more rows in this tile than the staged window holds (many empty/short rows)
                    lo = (int64_t)rp[r];
                    hi = (int64_t)rp[r + 1];
                }
                a = (int)((lo > base ? lo : base) - base);
                b = (int)((hi < end ? hi : end) - base);
                if (b < a) b = a;
            }
            const bool is_long = (b - a) > kLongRow * LANES;
            const int bn = is_long ? a : b;  // long slices are left to the warp loop below
            double sum = 0.0;
            if (G > 0) {
                for (int kb = a; kb < bn; kb += 8 * LANES) {
                    double v[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = kb + g + u * LANES;
                        const bool ok = k < bn;
                        v[u] = ok ? pv[k] : 0.0;
                        xv[u] = ok ? ldx(x + pc[k]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (kb + g + u * LANES < bn) sum = __dadd_rn(sum, __dmul_rn(v[u], xv[u]));
                }
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            } else {
                int k = a;
                for (; k + 4 <= bn; k += 4) {
                    const double v0 = pv[k], v1 = pv[k + 1], v2 = pv[k + 2], v3 = pv[k + 3];
                    sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, v0), v1), v2), v3);
                }
                for (; k < bn; ++k) sum = __dadd_rn(sum, pv[k]);
            }
            if (g == 0 && bn > a) red_add_f64(y + r, sum);

            // Long slices: the whole warp sums each one (strided, then a shuffle tree).
            unsigned longs = __ballot_sync(0xffffffffu, is_long && g == 0);
            while (longs) {
                const int src = __ffs(longs) - 1;
                longs &= longs - 1;
                const int la = __shfl_sync(0xffffffffu, a, src);
                const int lb = __shfl_sync(0xffffffffu, b, src);
                const int lr = __shfl_sync(0xffffffffu, r, src);
                double ls = 0.0;
                if (G > 0) {
                    for (int k = la + lane; k < lb; k += 32) ls = __dadd_rn(ls, __dmul_rn(pv[k], ldx(x + pc[k])));
                } else {
                    for (int k = la + lane; k < lb; k += 32) ls = __dadd_rn(ls, pv[k]);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
                if (lane == 0) red_add_f64(y + lr, ls);
            }
        }
        __syncthreads();

        if (tid == 0 && in < ntiles) {
            fence_proxy_async();
            issue(s, in, nrr);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

struct CsrConfig {
    int threads, tile, stages, lanes;  // lanes 0 = product mode
};

static CsrConfig csr_config(const Matrix * m)
{
    CsrConfig c;
    const int64_t avg = m->rows > 0 ? (m->stored + m->rows - 1) / m->rows : 0;
    // csr.algo: 0 automatic, 1 direct, 2 product.  Automatic: direct, lanes per row from the mean
    // row length; product when rows are long on average.
    int lanes = avg <= 10 ? 1 : avg <= 20 ? 2 : avg <= 48 ? 4 : avg <= 160 ? 8 : 0;
    if (m->opt_csr_algo == 2) lanes = 0;
    if (m->opt_csr_algo == 1 && lanes == 0) lanes = 8;
    if (m->opt_csr_lanes) lanes = (int)m->opt_csr_lanes;
    c.lanes = lanes;
    // defaults from the sweeps in profiles/: short rows like 256-thread CTAs, shared rows 128
    c.threads = (int)(m->opt_csr_threads ? m->opt_csr_threads : (lanes >= 2 ? 128 : 256));
    c.tile = (int)(m->opt_csr_tile ? m->opt_csr_tile : (lanes == 0 ? 2048 : 1024));
    c.stages = (int)(m->opt_csr_stages ? m->opt_csr_stages : 2);
    return c;
}

template <typename OffT>
static int csr_build_table(Matrix * m, int tile, int grid)
{
    if (m->tile_row) {
        cudaFree(m->tile_row);
        m->tile_row = nullptr;
    }
    m->csr_tile = tile;
    m->csr_grid = grid;
    // equal contiguous chunks of non-zeros, multiples of 16 entries so every bulk copy is 16 B aligned
    m->csr_chunk = round_up((m->stored + grid - 1) / grid, 16);
    m->csr_tpc = (int)((m->csr_chunk + tile - 1) / tile);
    const int64_t n = (int64_t)grid * m->csr_tpc;
    int2 * table = nullptr;
    SPMV_TRY(dev_alloc((Matrix *)nullptr, &table, n));
    m->tile_row = reinterpret_cast<int32_t *>(table);
    csr_tile_table_kernel<OffT><<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(
        m->rows, m->stored, m->csr_chunk, m->csr_tpc, tile, grid, (const OffT *)m->rp, table);
    SPMV_CUDA(cudaGetLastError());
    m->aux_dirty = true;
    return 0;
}

int csr_build_tiles(Matrix * m)
{
    // The table depends on the launch configuration; it is (re)built lazily by launch_csr.
    if (m->tile_row) {
        cudaFree(m->tile_row);
        m->tile_row = nullptr;
    }
    m->csr_grid = 0;
    return 0;
}

template <typename OffT, int THREADS, int TILE, int STAGES, int G>
static int launch_csr_variant(Matrix * m)
{
    auto kernel = csr_stream_kernel<OffT, THREADS, TILE, STAGES, G>;
    constexpr size_t smem = csr_smem_bytes<OffT, TILE, STAGES>();
    static std::atomic<int> occupancy_of[64];  // per instantiation and device (the attribute is per device)
    const int dev = m->device >= 0 && m->device < 64 ? m->device : 0;
    int occupancy = occupancy_of[dev].load(std::memory_order_acquire);
    if (!occupancy) {
        SPMV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPMV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, kernel, THREADS, smem));
        if (occupancy < 1) return fail(SPMVB200_ERR_CUDA, "csr_stream_kernel does not fit on an SM");
        occupancy_of[dev].store(occupancy, std::memory_order_release);
    }
    // csr.spare_ctas: leave that many CTA slots per SM free, e.g. for the NCCL kernel of an exchange
    // that should run concurrently (a persistent grid at full occupancy would lock it out).
    int ctas = m->opt_csr_ctas ? (int)std::min<int64_t>(m->opt_csr_ctas, occupancy) : occupancy;
    ctas = std::max<int>(1, ctas - (int)m->opt_csr_spare);
    // persistent grid: one wave, but never more CTAs than there are 16-entry groups
    int64_t grid = std::min<int64_t>((int64_t)m->sm_count * ctas, std::max<int64_t>(1, (m->stored + 15) / 16));
    if (!m->tile_row || m->csr_tile != TILE || m->csr_grid != (int)grid) SPMV_TRY(csr_build_table<OffT>(m, TILE, (int)grid));
    if (m->dry_run) return 0;
    const RunMode rm = run_mode(m);
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, THREADS, smem, m->stream, rm.pdl, m->stored, m->csr_chunk,
                            m->csr_tpc, rm.independent, (const OffT *)m->rp, (const int32_t *)m->col, (const double *)m->val,
                            (const int2 *)m->tile_row, (const double *)m->x, m->y));
    count_launch();
    return 0;
}

template <typename OffT, int THREADS, int TILE, int STAGES>
static int launch_csr_lanes(Matrix * m, int lanes)
{
    switch (lanes) {
    case 0: return launch_csr_variant<OffT, THREADS, TILE, STAGES, 0>(m);
    case 1: return launch_csr_variant<OffT, THREADS, TILE, STAGES, 1>(m);
    case 2: return launch_csr_variant<OffT, THREADS, TILE, STAGES, 2>(m);
    case 4: return launch_csr_variant<OffT, THREADS, TILE, STAGES, 4>(m);
    case 8: return launch_csr_variant<OffT, THREADS, TILE, STAGES, 8>(m);
    }
    return fail(SPMVB200_ERR_INVALID, "csr.lanes must be 1, 2, 4 or 8");
}

template <typename OffT, int THREADS, int TILE>
static int launch_csr_stages(Matrix * m, const CsrConfig & c)
{
    switch (c.stages) {
    case 2: return launch_csr_lanes<OffT, THREADS, TILE, 2>(m, c.lanes);
    case 3: return launch_csr_lanes<OffT, THREADS, TILE, 3>(m, c.lanes);
    }
    return fail(SPMVB200_ERR_INVALID, "csr.stages must be 2 or 3");
}

template <typename OffT>
static int launch_csr_t(Matrix * m, const CsrConfig & c)
{
    if (c.threads == 32) {
        if (c.tile == 256) return launch_csr_stages<OffT, 32, 256>(m, c);
        if (c.tile == 512) return launch_csr_stages<OffT, 32, 512>(m, c);
    } else if (c.threads == 64) {
        if (c.tile == 256) return launch_csr_stages<OffT, 64, 256>(m, c);
        if (c.tile == 512) return launch_csr_stages<OffT, 64, 512>(m, c);
    } else if (c.threads == 128) {
        if (c.tile == 512) return launch_csr_stages<OffT, 128, 512>(m, c);
        if (c.tile == 1024) return launch_csr_stages<OffT, 128, 1024>(m, c);
    } else if (c.threads == 256) {
        if (c.tile == 1024) return launch_csr_stages<OffT, 256, 1024>(m, c);
        if (c.tile == 2048) return launch_csr_stages<OffT, 256, 2048>(m, c);
    }
    return fail(SPMVB200_ERR_INVALID, "unsupported csr.threads / csr.tile combination");
}

int launch_csr_warp(Matrix * m, int lanes);  // kernels_csr_warp.cu
int launch_csr_flat(Matrix * m);             // kernels_csr_flat.cu
int launch_csr_sliced(Matrix * m);           // kernels_csr_sliced.cu
int csr_max_row_length(Matrix * m);          // builders.cu

bool csr_uses_sliced_kernel(Matrix * m)
{
    if (m->format != SPMVB200_CSR || m->rows == 0 || m->stored == 0 || m->opt_csr_probe != 0) return false;
    if (m->opt_csr_algo == 5) return true;
    if (m->opt_csr_algo != 0 || m->slice_unavailable) return false;
    const int64_t avg = m->stored / std::max<int64_t>(m->rows, 1);
    if (avg < 10) return false;
    if (csr_max_row_length(m) != 0) return false;
    return m->csr_maxlen <= 2 * avg;
}

int launch_csr(Matrix * m)
{
    if (m->rows == 0) return 0;
    if (m->stored == 0) return m->dry_run ? 0 : clear_y_for_beta0(m);  // nothing to add; y = alpha*A*x still clears y
    const CsrConfig c = csr_config(m);
    // csr.algo: 0 automatic, 1 stream/direct, 2 stream/product, 3 warp-granular register-staged,
    // 4 flat (register-staged, split by non-zeros, rows from span metadata; kernels_csr_flat.cu).
    // Automatic = flat: fastest on every matrix measured (profiles/r01_sweep_j_csr_flat.log: 2D 5-point
    // 14.3 vs 16.2 us, 3D 7-point 35.0 vs 38.1 us, 27-point 256^3 1.01 vs 1.07 ms, R-MAT 2^24 1.18 vs
    // 1.55 ms for the best of the others).
    // 5 sliced (lane per row on a slot-major copy; kernels_csr_sliced.cu): automatic for long regular rows
    // (mean >= 10 entries, longest row <= 2x the mean), where the flat kernel's gathers saturate the L1.
    if (m->opt_csr_probe != 0) {  // the traffic probes live in the flat kernel
        SPMV_TRY(csr_ensure_row_major(m));
        return launch_csr_flat(m);
    }
    if (m->opt_csr_algo == 5) return launch_csr_sliced(m);
    if (m->opt_csr_algo == 0) {
        const int64_t avg = m->stored / std::max<int64_t>(m->rows, 1);
        if (avg >= 10) {
            SPMV_TRY(csr_max_row_length(m));
            if (m->csr_maxlen <= 2 * avg && !m->slice_unavailable) {
                const int rc = launch_csr_sliced(m);
                if (rc == 0 || m->slice_col) return rc;
                // no room for the second copy of the entries: the flat kernel needs only 0.25 B per entry
                m->slice_unavailable = true;
                cudaGetLastError();
            }
        }
    }
    SPMV_TRY(csr_ensure_row_major(m));  // every other kernel walks the row-major arrays
    if (m->opt_csr_algo == 0 || m->opt_csr_algo == 4) return launch_csr_flat(m);
    SPMV_TRY(need_unit_alpha(m, "this CSR kernel"));
    if (!m->dry_run) SPMV_TRY(clear_y_for_beta0(m));
    const bool warp = m->opt_csr_algo == 3;
    if (warp) return launch_csr_warp(m, c.lanes > 0 ? c.lanes : 8);
    m->kernel_name = c.lanes == 0 ? "csr_stream_kernel<product>" : "csr_stream_kernel<direct>";
    return m->off64 ? launch_csr_t<int64_t>(m, c) : launch_csr_t<uint32_t>(m, c);
}

}  // namespace spmvb200
