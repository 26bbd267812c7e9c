// kernels_csr_flat.cu -- CSR SpMV over the flat non-zero stream:  y += A*x.
//
// Same job as csr_stream_kernel (kernels_csr.cu; reference matrix/csr-matrix-spmv.cpp:21-33, 63-76).
// The work is cut by NON-ZEROS, never by rows: a warp owns 128 consecutive stored entries, four per
// lane, whatever rows they belong to, so every warp streams the same number of bytes for stencils
// and power-law matrices alike.
//
//   * Lane l fetches entries 4l..4l+3 of the span with one 128-bit load (column indices) and one
//     256-bit load (values): 1.5 KB in flight per warp in registers, no shared-memory staging, no
//     CTA barrier, 12+ warps per scheduler resident.
//   * The row of every entry comes from span metadata built once with the matrix (32 B per 128
//     entries): the row of the span's first entry and a 128-bit mask of the entries at which a new row
//     starts, so row(e) = first_row + popcount(mask bits 1..e) -- a handful of integer instructions,
//     no dependent load behind the streamed data.  The kernel therefore reads the metadata INSTEAD of
//     row_ptr (0.25 B per entry instead of 4 B per row); row_ptr stays resident for export, partition
//     and the other kernels.  Empty rows (several rows start at one entry, which a bit cannot say) are
//     handled by numbering only the NON-EMPTY rows in the metadata and translating back through a row map
//     (int32 per non-empty row) where a sum leaves for y.  ("csr.rowptr_path" = 1 selects the older
//     MASK = false path instead: the rows that start inside the span are read from row_ptr, scattered as
//     "row - first_row" into a warp-private line of shared memory (atomicMax) and turned into row numbers
//     by an inclusive max-scan -- a dependent trip through row_ptr.)
//     All of this reads immutable matrix data only and runs before griddepcontrol.wait.
//   * x is gathered through the read-only path, four gathers in flight per lane.
//   * The products are summed per run of equal rows by warp_segmented_add4 (segreduce.cuh): serially
//     in the lane, one segmented scan across the warp, one fp64 reduction (RED.ADD.F64) per run end.
//
// Arithmetic order: products of a row are added left to right inside a lane and lane sums are
// combined in scan order, so rows longer than 4 entries differ from the reference's strictly
// sequential sum in the last bits, within BASELINE.json's |y - y_ref| <= 1e-12 * sum_j |a_ij x_j|.
// ("csr.algo" = 1 with "csr.lanes" = 1 keeps the bit-identical order for rows inside a tile.)
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "segreduce.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// E = entries per lane (4 or 8): a span is 32*E entries, its metadata 4 + E ints ({first row, -, -, -,
// E mask words}).  E = 8 halves the number of warps and of cross-lane reductions per entry and doubles the
// bytes a warp has in flight, which pays on small matrices whose kernels last a few warp lifetimes
// (2-D 5-point 1000^2: 13.4-14.3 us with E = 4); E = 4 keeps 32 registers and full occupancy.
template <int E>
struct FlatShape {
    static constexpr int span = 32 * E;
    static constexpr int stride = 4 + E;  // ints of metadata per span
};

// meta[stride*w + 0] = largest r with row_ptr[r] <= span*w;  meta[stride*w + 4 ..] = bit e set iff a row starts
// at entry span*w + e, e = 1..span-1 (a row starting at e = 0 is the first row itself).
// cidx (optional): number of non-empty rows before row r; then the metadata holds that number instead of r.
template <typename OffT>
__global__ void csr_flat_first_row_kernel(int64_t rows, int64_t nspans, int span, int stride, const OffT * __restrict__ rp,
                                          const int32_t * __restrict__ cidx, int32_t * __restrict__ meta)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nspans) return;
    const int64_t target = w * span;
    int64_t lo = 0, hi = rows - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if ((int64_t)rp[mid] <= target) lo = mid; else hi = mid - 1;
    }
    meta[(int64_t)stride * w] = cidx ? cidx[lo] : (int32_t)lo;
}

template <typename OffT>
__global__ void csr_nonempty_flag_kernel(int64_t rows, const OffT * __restrict__ rp, int32_t * __restrict__ flag)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        flag[r] = (int64_t)rp[r + 1] > (int64_t)rp[r] ? 1 : 0;
}

template <typename OffT>
__global__ void csr_rowmap_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ cidx,
                                  int32_t * __restrict__ rowmap)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        if ((int64_t)rp[r + 1] > (int64_t)rp[r]) rowmap[cidx[r]] = (int32_t)r;
}

template <typename OffT>
__global__ void csr_flat_mask_kernel(int64_t rows, int span, int stride, const OffT * __restrict__ rp, int32_t * __restrict__ meta,
                                     int * __restrict__ has_empty)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = (int64_t)rp[r];
        if ((int64_t)rp[r + 1] == p) { *has_empty = 1; continue; }
        const int e = (int)(p % span);
        if (e) atomicOr(reinterpret_cast<unsigned *>(meta) + (int64_t)stride * (p / span) + 4 + (e >> 5), 1u << (e & 31));
    }
}

template <typename OffT>
static int csr_build_flat_meta_t(Matrix * m, int span, int stride)
{
    cudaStream_t s = m->stream;
    const OffT * rp = (const OffT *)m->rp;
    const int64_t nspans = (m->stored + span - 1) / span;
    SPMV_TRY(dev_alloc((Matrix *)nullptr, &m->flat_meta, (int64_t)stride * (nspans + 1)));
    m->flat_span = span;
    SPMV_CUDA(cudaMemsetAsync(m->flat_meta, 0, sizeof(int32_t) * (size_t)stride * (size_t)(nspans + 1), s));
    Scratch<int> flag;
    SPMV_TRY(flag.alloc(1));
    SPMV_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
    csr_flat_mask_kernel<OffT><<<grid_for(m->rows, m->sm_count), 256, 0, s>>>(m->rows, span, stride, rp, m->flat_meta, flag.p);
    SPMV_CUDA(cudaGetLastError());
    int has_empty = 0;
    SPMV_CUDA(cudaMemcpyAsync(&has_empty, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    m->flat_has_empty = has_empty != 0;
    Scratch<int32_t> nonempty, cidx;
    if (m->flat_has_empty && !m->flat_rowmap) {  // number the non-empty rows: cidx[r] = how many lie before row r
        Scratch<unsigned char> tmp;
        SPMV_TRY(nonempty.alloc(m->rows + 1)); SPMV_TRY(cidx.alloc(m->rows + 1));
        SPMV_CUDA(cudaMemsetAsync(nonempty.p + m->rows, 0, sizeof(int32_t), s));
        csr_nonempty_flag_kernel<OffT><<<grid_for(m->rows, m->sm_count), 256, 0, s>>>(m->rows, rp, nonempty.p);
        SPMV_CUDA(cudaGetLastError());
        size_t tb = 0;
        SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nonempty.p, cidx.p, m->rows + 1, s));
        SPMV_TRY(tmp.alloc((int64_t)tb));
        SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, nonempty.p, cidx.p, m->rows + 1, s));
        int32_t count = 0;
        SPMV_CUDA(cudaMemcpyAsync(&count, cidx.p + m->rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
        SPMV_TRY(dev_alloc((Matrix *)nullptr, &m->flat_rowmap, (int64_t)count + 1));
        csr_rowmap_kernel<OffT><<<grid_for(m->rows, m->sm_count), 256, 0, s>>>(m->rows, rp, cidx.p, m->flat_rowmap);
        SPMV_CUDA(cudaGetLastError());
    } else if (m->flat_has_empty) {  // span size changed: the numbering is needed again for the first-row entries
        Scratch<unsigned char> tmp;
        SPMV_TRY(nonempty.alloc(m->rows + 1)); SPMV_TRY(cidx.alloc(m->rows + 1));
        SPMV_CUDA(cudaMemsetAsync(nonempty.p + m->rows, 0, sizeof(int32_t), s));
        csr_nonempty_flag_kernel<OffT><<<grid_for(m->rows, m->sm_count), 256, 0, s>>>(m->rows, rp, nonempty.p);
        size_t tb = 0;
        SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nonempty.p, cidx.p, m->rows + 1, s));
        SPMV_TRY(tmp.alloc((int64_t)tb));
        SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, nonempty.p, cidx.p, m->rows + 1, s));
        SPMV_CUDA(cudaStreamSynchronize(s));
    }
    // the MASK = false path ("csr.rowptr_path") wants real row numbers in the metadata, the mask path the
    // numbering of the non-empty rows
    m->flat_meta_compressed = m->flat_has_empty && !m->opt_csr_rowptr_path;
    csr_flat_first_row_kernel<OffT><<<(unsigned)((nspans + 255) / 256), 256, 0, s>>>(
        m->rows, nspans, span, stride, rp, m->flat_meta_compressed ? cidx.p : nullptr, m->flat_meta);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

static int csr_build_flat_meta(Matrix * m, int span, int stride)
{
    const bool want_compressed = m->flat_has_empty && !m->opt_csr_rowptr_path;
    if (m->flat_meta && m->flat_span == span && m->flat_meta_compressed == want_compressed) return 0;
    if (m->flat_meta) {
        cudaFree(m->flat_meta);
        m->flat_meta = nullptr;
    }
    return m->off64 ? csr_build_flat_meta_t<int64_t>(m, span, stride) : csr_build_flat_meta_t<uint32_t>(m, span, stride);
}

// PROBE: 0 = y += A*x; 1 = "regular traffic": y_i += sum_k a_k, the matrix values streamed, no gather
// (csr_spmv_inner_loop_regular_traffic, csr-matrix-spmv.cpp:35-47); 2 = "irregular traffic": y_i += sum_k
// x[j_k], the gather alone without the values (:49-61).  The reference keeps the two as diagnostics; here
// they split a kernel's time into its streaming and its gather part.
template <typename OffT, int WARPS, int E, bool MASK, int PROBE>
__global__ void __launch_bounds__(WARPS * 32)
csr_flat_kernel(int64_t rows, int64_t stored, int64_t nspans, int independent, const OffT * __restrict__ rp,
                const int32_t * __restrict__ col, const double * __restrict__ val,
                const int32_t * __restrict__ meta, const int32_t * __restrict__ rowmap, const double * __restrict__ x,
                double * __restrict__ y, double alpha)
{
    constexpr int SPAN = FlatShape<E>::span, STRIDE = FlatShape<E>::stride;
    __shared__ __align__(16) int32_t smark[MASK ? 1 : WARPS][SPAN];

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * WARPS + warp;
    if (w >= nspans) return;  // no CTA-wide barrier below
    const int64_t kw = w * SPAN;
    const int64_t kend = min(kw + (int64_t)SPAN, stored);
    const int64_t k0 = kw + E * lane;
    const uint64_t pol = policy_evict_first();

    // the span in flight (arrays are padded past `stored` with (0, 0.0)): E/4 128-bit and E/4 256-bit loads per lane
    int c[E];
    double a[E];
#pragma unroll
    for (int q = 0; q < E / 4; ++q) {
        int4 c4 = make_int4(0, 0, 0, 0);
        if (PROBE != 1) c4 = ldg_stream_i4(col + k0 + 4 * q, pol);
        c[4 * q] = c4.x; c[4 * q + 1] = c4.y; c[4 * q + 2] = c4.z; c[4 * q + 3] = c4.w;
        double a4[4] = {1.0, 1.0, 1.0, 1.0};
        if (PROBE != 2) ldg_stream_d4(val + k0 + 4 * q, a4);
        a[4 * q] = a4[0]; a[4 * q + 1] = a4[1]; a[4 * q + 2] = a4[2]; a[4 * q + 3] = a4[3];
    }
    int r[E];
    const int r_lo = __ldg(meta + (int64_t)STRIDE * w);
    if (MASK) {
        unsigned mk[E];
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
            const int4 t = __ldg(reinterpret_cast<const int4 *>(meta + (int64_t)STRIDE * w + 4) + q);
            mk[4 * q] = (unsigned)t.x; mk[4 * q + 1] = (unsigned)t.y; mk[4 * q + 2] = (unsigned)t.z; mk[4 * q + 3] = (unsigned)t.w;
        }
        const int word = (E * lane) >> 5, sh = (E * lane) & 31;
        unsigned mw = mk[0];
        int below = 0;
#pragma unroll
        for (int q = 1; q < E; ++q) {
            if (word >= q) { below += __popc(mk[q - 1]); mw = mk[q]; }
        }
        r[0] = r_lo + below + __popc(mw & ((2u << sh) - 1u));  // rows started at entries 1 .. E*lane
#pragma unroll
        for (int j = 1; j < E; ++j) r[j] = r[j - 1] + (int)((mw >> (sh + j)) & 1u);  // sh + j <= 31: sh is a multiple of E
    } else {
        // marks: mark[k - kw] = (last row that starts at entry k) - r_lo, 0 where no row starts
        int32_t * mark = smark[warp];
#pragma unroll
        for (int q = 0; q < E / 4; ++q) *reinterpret_cast<int4 *>(mark + E * lane + 4 * q) = make_int4(0, 0, 0, 0);
        __syncwarp();
        for (int64_t rb = (int64_t)r_lo + 1; rb < rows; rb += 32) {
            const int64_t rr = rb + lane;
            const int64_t p = rr < rows ? (int64_t)rp[rr] : LLONG_MAX;  // rows above r_lo start after kw
            if (p < kend) atomicMax(mark + (int)(p - kw), (int)(rr - r_lo));
            if (__shfl_sync(0xffffffffu, p >= kend ? 1 : 0, 31)) break;  // row_ptr is monotone
        }
        __syncwarp();
        int o[E];
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
            const int4 t = *reinterpret_cast<const int4 *>(mark + E * lane + 4 * q);
            o[4 * q] = t.x; o[4 * q + 1] = t.y; o[4 * q + 2] = t.z; o[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int j = 1; j < E; ++j) o[j] = max(o[j - 1], o[j]);
        int run = o[E - 1];  // inclusive max-scan of the lanes' last marks
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, run, d);
            if (lane >= d) run = max(run, t);
        }
        int before = __shfl_up_sync(0xffffffffu, run, 1);
        if (lane == 0) before = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) r[j] = r_lo + max(before, o[j]);
    }
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (k0 + j >= kend) r[j] = -1;

    // Everything above reads only the immutable matrix; x and y may come from the previous launch.
    if (!independent) asm volatile("griddepcontrol.wait;" ::: "memory");
    double p[E];
#pragma unroll
    for (int j = 0; j < E; ++j) p[j] = PROBE != 1 ? ldx(x + c[j]) : 1.0;
#pragma unroll
    for (int j = 0; j < E; ++j) p[j] = __dmul_rn(a[j], p[j]);
    warp_segmented_add<E>(lane, r, p, y, alpha, MASK ? rowmap : nullptr);
}

template <typename OffT, int WARPS, int E>
static int launch_flat_variant(Matrix * m)
{
    SPMV_TRY(csr_build_flat_meta(m, FlatShape<E>::span, FlatShape<E>::stride));
    if (m->dry_run) return 0;
    const int64_t nspans = (m->stored + FlatShape<E>::span - 1) / FlatShape<E>::span;
    const int64_t grid = (nspans + WARPS - 1) / WARPS;
    if (grid > INT_MAX) return fail(SPMVB200_ERR_OVERFLOW, "CSR matrix too large for one launch");
    SPMV_TRY(clear_y_for_beta0(m));
    const RunMode rm = run_mode(m);
    const bool rowptr_path = m->flat_has_empty && !m->flat_meta_compressed;
    auto kernel = rowptr_path ? csr_flat_kernel<OffT, WARPS, E, false, 0> : csr_flat_kernel<OffT, WARPS, E, true, 0>;
    if (WARPS == 4 && E == 4 && m->opt_csr_probe == 1)
        kernel = rowptr_path ? csr_flat_kernel<OffT, 4, 4, false, 1> : csr_flat_kernel<OffT, 4, 4, true, 1>;
    else if (WARPS == 4 && E == 4 && m->opt_csr_probe == 2)
        kernel = rowptr_path ? csr_flat_kernel<OffT, 4, 4, false, 2> : csr_flat_kernel<OffT, 4, 4, true, 2>;
    else if (m->opt_csr_probe != 0)
        return fail(SPMVB200_ERR_INVALID, "csr.probe must be 0, 1 or 2 (with csr.threads 128 and csr.entries 4)");
    SPMV_CUDA(launch_kernel(kernel, (unsigned)grid, WARPS * 32u, 0, m->stream, rm.pdl, m->rows,
                            m->stored, nspans, rm.independent, (const OffT *)m->rp, (const int32_t *)m->col,
                            (const double *)m->val, (const int32_t *)m->flat_meta,
                            (const int32_t *)(m->flat_meta_compressed ? m->flat_rowmap : nullptr), (const double *)m->x, m->y,
                            m->alpha));
    count_launch();
    return 0;
}

template <typename OffT>
static int launch_flat_off(Matrix * m, int threads, int entries)
{
    if (entries == 4) {
        if (threads == 64) return launch_flat_variant<OffT, 2, 4>(m);
        if (threads == 128) return launch_flat_variant<OffT, 4, 4>(m);
        if (threads == 256) return launch_flat_variant<OffT, 8, 4>(m);
    } else if (entries == 8) {
        if (threads == 64) return launch_flat_variant<OffT, 2, 8>(m);
        if (threads == 128) return launch_flat_variant<OffT, 4, 8>(m);
        if (threads == 256) return launch_flat_variant<OffT, 8, 8>(m);
    } else {
        return fail(SPMVB200_ERR_INVALID, "csr.entries must be 4 or 8");
    }
    return fail(SPMVB200_ERR_INVALID, "csr.threads must be 64, 128 or 256 for the flat kernel");
}

int launch_csr_flat(Matrix * m)
{
    const int threads = (int)(m->opt_csr_threads ? m->opt_csr_threads : 128);
    // the probes are built for 4 entries per lane only
    const int entries = m->opt_csr_probe ? 4 : (int)(m->opt_csr_entries ? m->opt_csr_entries : 4);
    m->kernel_name = m->opt_csr_probe == 1 ? "csr_flat_kernel<regular traffic>"
                     : m->opt_csr_probe == 2 ? "csr_flat_kernel<irregular traffic>" : "csr_flat_kernel";
    return m->off64 ? launch_flat_off<int64_t>(m, threads, entries) : launch_flat_off<uint32_t>(m, threads, entries);
}

}  // namespace spmvb200
