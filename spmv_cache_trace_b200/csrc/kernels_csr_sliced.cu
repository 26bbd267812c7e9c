// kernels_csr_sliced.cu -- CSR SpMV with a lane-per-row mapping on a slice-interleaved copy:  y += A*x.
//
// Same job as csr_flat_kernel (reference matrix/csr-matrix-spmv.cpp:21-33, 63-76) for matrices whose rows
// are long and regular (wide stencils).  There the flat kernel is bound by the L1, not by HBM: a lane's
// four consecutive entries of a 27-wide row gather x from ~17 sectors per instruction (l1tex 99 % busy,
// profiles/r01_ncu_c5s_csr_flat.txt), whereas the ELL kernel, whose 32 lanes hold 32 consecutive ROWS of
// one slot, touches 8 full sectors and runs the same matrix at 6.76 TB/s against 5.67 TB/s.
//
// This kernel gives CSR that mapping without ELL's padding.  The builder keeps a second copy of
// column_index / value in which the entries of every slice of 32 consecutive rows are stored SLOT-MAJOR:
// first the first entries of all rows of the slice that have one (ascending row), then the second
// entries, and so on.  A slice holds exactly the entries it holds in plain CSR, so row_ptr[32 s] is also
// the slice's offset in the copy, the copy has `stored` entries, and no byte of padding exists.  Lane i
// of a warp owns row 32 s + i; for slot l the lanes whose row is longer than l read consecutive
// addresses (position = slice offset + entries of earlier slots + rank of the lane among the active
// ones, from one ballot and two popcounts), gather x from consecutive rows -- adjacent columns for a
// banded matrix -- and add the product to their row's sum strictly left to right with separate
// multiply/add roundings: bit-identical to the reference's loop for every row.
//
// Chosen automatically when the mean row length is at least 10 and the longest row at most twice the
// mean (lanes idle while the longest row of their slice finishes); "csr.algo" = 5 forces it.
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// One warp per slice: copy the slice's entries from row-major to slot-major order.
template <typename OffT>
__global__ void csr_slice_fill_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ col,
                                      const double * __restrict__ val, int32_t * __restrict__ scol, double * __restrict__ sval)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // row (blockDim is a multiple of 32)
    if (i - lane >= rows) return;
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    int64_t pos = __shfl_sync(0xffffffffu, lo, 0);
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const int64_t p = pos + __popc(mask & ((1u << lane) - 1u));
            scol[p] = col[lo + l];
            sval[p] = val[lo + l];
        }
        pos += __popc(mask);
    }
}

// U = slots whose matrix loads are in flight together
// The inverse of csr_slice_fill_kernel.
template <typename OffT>
__global__ void csr_slice_unfill_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ scol,
                                        const double * __restrict__ sval, int32_t * __restrict__ col, double * __restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i - lane >= rows) return;
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    int64_t pos = __shfl_sync(0xffffffffu, lo, 0);
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const int64_t p = pos + __popc(mask & ((1u << lane) - 1u));
            col[lo + l] = scol[p];
            val[lo + l] = sval[p];
        }
        pos += __popc(mask);
    }
}

int csr_ensure_row_major(Matrix * m)
{
    if (m->format != SPMVB200_CSR || (m->col && m->val)) return 0;
    if (!m->slice_col || !m->slice_val) return fail(SPMVB200_ERR_INVALID, "CSR matrix holds neither copy of its entries");
    int rc = alloc_streamed(m, &m->col, m->stored);
    if (rc == 0) rc = alloc_streamed(m, &m->val, m->stored);
    if (rc) {
        if (m->col) cudaFree(m->col);
        m->col = nullptr;
        m->val = nullptr;
        return rc;
    }
    const unsigned grid = (unsigned)((m->rows + 127) / 128);
    if (m->off64) csr_slice_unfill_kernel<int64_t><<<grid, 128, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, m->slice_col, m->slice_val, m->col, m->val);
    else csr_slice_unfill_kernel<uint32_t><<<grid, 128, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, m->slice_col, m->slice_val, m->col, m->val);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

template <typename OffT, int U, int THREADS, bool PUSH = false>
__global__ void __launch_bounds__(THREADS, (U <= 4 ? 2048 : 1024) / THREADS)
csr_sliced_kernel(int64_t row0, int64_t rows, int independent, int store, double alpha, const OffT * __restrict__ rp,
                  const int32_t * __restrict__ scol, const double * __restrict__ sval, const double * __restrict__ x,
                  double * __restrict__ y, const double * __restrict__ y_in_host, double * __restrict__ y_out_host,
                  double * push0, int64_t push0_lo, int64_t push0_hi, double * push1, int64_t push1_lo, int64_t push1_hi)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    // rows [row0, rows) of the matrix (row0 a multiple of 32: a warp is a slice)
    const int64_t i = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i - lane >= rows) return;  // whole warp past the end
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    int64_t pos = __shfl_sync(0xffffffffu, lo, 0);  // offset of the slice = row_ptr of its first row
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    const unsigned below = (1u << lane) - 1u;
    double z = 0.0;
    bool waited = independent != 0;
    for (int l0 = 0; l0 < maxlen; l0 += U) {
        int c[U];
        double a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {  // warp-uniform trip count: every lane takes part in the ballots
            const bool active = len > l0 + u;
            const unsigned mask = __ballot_sync(0xffffffffu, active);
            const int64_t p = pos + __popc(mask & below);
            // Plain read-only loads (L1 allocation, normal L2 policy), unlike the other kernels' streams: a
            // slot's 128 / 256 B of a slice start wherever the previous slot ended, so consecutive requests share
            // sectors, and with L1::no_allocate + L2 evict-first the shared sectors were fetched from DRAM twice
            // (6.06 GB read for 5.73 GB on 27-point 256^3; 512^3: 7.31 -> 7.00 ms with plain loads).
            c[u] = active ? __ldg(scol + p) : 0;
            a[u] = active ? __ldg(sval + p) : 0.0;
            pos += __popc(mask);
        }
        if (!waited) {  // the matrix is immutable; x and y may come from the previous launch
            asm volatile("griddepcontrol.wait;" ::: "memory");
            waited = true;
        }
        double xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = len > l0 + u ? ldx(x + c[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (len > l0 + u) z = __dadd_rn(z, __dmul_rn(a[u], xv[u]));
    }
    if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
    // Zero-copy form (spmvb200_spmv_host with pinned buffers): y_old is read from and y_new written to mapped HOST
    // memory by this kernel -- a warp moves 256 contiguous bytes each way -- so the two PCIe directions run at once
    // and y never takes a separate trip through the copy engine.
    if (y_out_host) {
        if (i < rows) {
            const double yo = y_in_host ? __ldcs(y_in_host + i) : 0.0;
            __stcs(y_out_host + i, __dadd_rn(yo, __dmul_rn(alpha, z)));
        }
        return;
    }
    // a lane owns its whole row: y = alpha*A*x is a plain store (no clearing pass, no read of y)
    if (store == 1) {
        if (i < rows) {
            const double v = __dmul_rn(alpha, z);
            y[i] = v;
            // Fused halo push of the row-partitioned mode: the rows a neighbouring rank's rows reference are stored into
            // that rank's x buffer as well -- a peer-mapped pointer, the store travels over NVLink -- so the exchange of
            // the next step has nothing left to copy.  Consecutive lanes write consecutive addresses (256 B per warp).
            if (PUSH) {
                if (i >= push0_lo && i < push0_hi) push0[i] = v;
                if (i >= push1_lo && i < push1_hi) push1[i] = v;
            }
        }
    }
    else if (store == 2) {  // y += ...: plain read-modify-write by the lane that owns the row (the launch is ordered)
        if (len > 0) y[i] = __dadd_rn(y[i], __dmul_rn(alpha, z));
    }
    else if (len > 0) red_add_f64(y + i, __dmul_rn(alpha, z));
}

static int csr_drop_row_major(Matrix * m)
{
    if (!m->opt_csr_drop || !m->col || !m->slice_col) return 0;
    SPMV_CUDA(cudaStreamSynchronize(m->stream));  // the fill kernel reads them
    const int64_t cap = round_up(m->stored, 4096) + kPadEntries;
    cudaFree(m->col);
    cudaFree(m->val);
    m->col = nullptr;
    m->val = nullptr;
    m->device_bytes -= cap * 12;
    return 0;
}

static int csr_build_sliced(Matrix * m)
{
    if (m->slice_col && m->slice_val) return csr_drop_row_major(m);
    int rc = alloc_streamed(m, &m->slice_col, m->stored);
    if (rc == 0) rc = alloc_streamed(m, &m->slice_val, m->stored);
    if (rc) {  // leave nothing half-built behind
        if (m->slice_col) cudaFree(m->slice_col);
        m->slice_col = nullptr;
        m->slice_val = nullptr;
        return rc;
    }
    const unsigned grid = (unsigned)((m->rows + 127) / 128);
    if (m->off64) csr_slice_fill_kernel<int64_t><<<grid, 128, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, m->col, m->val, m->slice_col, m->slice_val);
    else csr_slice_fill_kernel<uint32_t><<<grid, 128, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, m->col, m->val, m->slice_col, m->slice_val);
    SPMV_CUDA(cudaGetLastError());
    m->aux_dirty = true;
    return csr_drop_row_major(m);
}

// Largest column referenced by each of `chunks` equal row chunks (chunk = rows_per_chunk rows, a multiple of 32), from the
// slot-major copy: a slice's entries are contiguous, [rp[32 s], rp[32 s + 32)).  One warp per slice.
template <typename OffT>
__global__ void csr_chunk_colmax_kernel(int64_t rows, int64_t rows_per_chunk, const OffT * __restrict__ rp,
                                        const int32_t * __restrict__ scol, int * __restrict__ colmax)
{
    const int lane = threadIdx.x & 31;
    const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r0 = slice * 32;
    if (r0 >= rows) return;
    const int64_t lo = (int64_t)rp[r0], hi = (int64_t)rp[min(r0 + 32, rows)];
    int best = -1;
    for (int64_t k = lo + lane; k < hi; k += 32) best = max(best, __ldg(scol + k));
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0 && best >= 0) atomicMax(colmax + r0 / rows_per_chunk, best);
}

int csr_chunk_colmax(Matrix * m, int64_t rows_per_chunk, int chunks, int * host_out)
{
    Scratch<int> d;
    SPMV_TRY(d.alloc(chunks));
    SPMV_CUDA(cudaMemsetAsync(d.p, 0xff, sizeof(int) * (size_t)chunks, m->stream));
    const int64_t warps = (m->rows + 31) / 32;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    if (m->off64) csr_chunk_colmax_kernel<int64_t><<<grid, 256, 0, m->stream>>>(m->rows, rows_per_chunk, (const int64_t *)m->rp, m->slice_col, d.p);
    else csr_chunk_colmax_kernel<uint32_t><<<grid, 256, 0, m->stream>>>(m->rows, rows_per_chunk, (const uint32_t *)m->rp, m->slice_col, d.p);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaMemcpyAsync(host_out, d.p, sizeof(int) * (size_t)chunks, cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int launch_csr_sliced(Matrix * m)
{
    SPMV_TRY(csr_build_sliced(m));
    m->kernel_name = "csr_sliced_kernel";
    if (m->dry_run) return 0;
    const int threads = (int)(m->opt_csr_threads ? m->opt_csr_threads : 128);
    // optional row range (the pipelined host-buffer path runs the matrix in row chunks)
    const int64_t row0 = m->range_end > 0 ? m->range_begin : 0;
    const int64_t row1 = m->range_end > 0 ? m->range_end : m->rows;
    const int64_t grid = (row1 - row0 + threads - 1) / threads;
    if (grid > INT_MAX) return fail(SPMVB200_ERR_OVERFLOW, "CSR matrix too large for one launch");
    if (grid <= 0) return 0;
    const RunMode rm = run_mode(m);
    const int store = (m->run_beta0 && !m->host_y_out) ? 1 : (m->run_rmw && !m->host_y_out && !rm.independent) ? 2 : 0;
    m->run_beta0 = false;
    const int batch = (int)(m->opt_csr_batch ? m->opt_csr_batch : 4);
#define SPMV_SLICED(OFF, UU, TT)                                                                                          \
    SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, UU, TT>, (unsigned)grid, (unsigned)TT, 0, m->stream, rm.pdl, row0, row1,  \
                            rm.independent, store, m->alpha, (const OFF *)m->rp, (const int32_t *)m->slice_col,          \
                            (const double *)m->slice_val, (const double *)m->x, m->y, (const double *)m->host_y_in,      \
                            m->host_y_out, m->push_y[0], m->push_lo[0], m->push_hi[0], m->push_y[1], m->push_lo[1],       \
                            m->push_hi[1]))
#define SPMV_SLICED_P(OFF)                                                                                                 \
    SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, 4, 128, true>, (unsigned)grid, 128u, 0, m->stream, rm.pdl, row0, row1,  \
                            rm.independent, store, m->alpha, (const OFF *)m->rp, (const int32_t *)m->slice_col,          \
                            (const double *)m->slice_val, (const double *)m->x, m->y, (const double *)m->host_y_in,      \
                            m->host_y_out, m->push_y[0], m->push_lo[0], m->push_hi[0], m->push_y[1], m->push_lo[1],       \
                            m->push_hi[1]))
#define SPMV_SLICED_T(UU, TT)                                             \
    do {                                                                  \
        if (m->off64) SPMV_SLICED(int64_t, UU, TT);                       \
        else SPMV_SLICED(uint32_t, UU, TT);                               \
    } while (0)
    if (m->push_y[0] || m->push_y[1]) {  // fused halo push: its own instantiation, so that the plain kernel keeps its 32 registers
        if (threads != 128 || batch != 4 || store != 1) return fail(SPMVB200_ERR_UNSUPPORTED, "halo push needs the default sliced kernel in store mode");
        if (m->off64) SPMV_SLICED_P(int64_t);
        else SPMV_SLICED_P(uint32_t);
    } else if (threads == 128) {
        if (batch == 4) SPMV_SLICED_T(4, 128);
        else if (batch == 8) SPMV_SLICED_T(8, 128);
        else if (batch == 2) SPMV_SLICED_T(2, 128);
        else return fail(SPMVB200_ERR_INVALID, "csr.batch must be 2, 4 or 8");
    } else if (threads == 256 && batch == 4) {
        SPMV_SLICED_T(4, 256);
    } else if (threads == 512 && batch == 4) {
        SPMV_SLICED_T(4, 512);
    } else {
        return fail(SPMVB200_ERR_INVALID, "sliced kernel: csr.threads 128 (csr.batch 2|4|8) or 256|512 (csr.batch 4)");
    }
#undef SPMV_SLICED_T
#undef SPMV_SLICED_P
#undef SPMV_SLICED
    count_launch();
    return 0;
}

}  // namespace spmvb200
