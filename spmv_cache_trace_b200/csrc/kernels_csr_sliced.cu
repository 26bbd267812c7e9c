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
//
// INDEX RUNS.  In the slot-major order the 32 column indices of a slot are, for a banded matrix, one arithmetic run:
// lane i holds column c0 + i (a stencil's rows are shifted copies of each other).  The builder detects that per
// (slice, slot) -- all active lanes satisfy col == base + lane -- and stores ONE int32 for such a slot instead of up to
// 32; a flag word per slice (bit l = slot l is a run; slots >= 32 are never runs) and the slice's offset in the
// compressed column stream tell the kernel which form a slot has.  Values are not touched, every row is still summed
// left to right from the same numbers: bit-identical results.  27-point 512^3: 88 % of the slots are runs (the rest are
// the slices that contain a grid-boundary row), the column stream shrinks from 14.4 GB to about 2 GB, and the kernel
// moves ~34 GB instead of 46 GB per product -- it is HBM-bound, so that is the speed-up.  The algorithmic bytes of the
// metric keep counting 4 B per stored column index like the reference's csr_matrix::size(); the measured DRAM traffic is
// reported beside them.  Used when the compressed stream is at most 3/4 of the plain one ("csr.index_runs": 0 auto,
// 1 always, -1 never); a matrix without runs (R-MAT) keeps the plain slot-major copy.
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

// One warp per slice: which slots are index runs (flag bit l) and how many int32 the slice's compressed column stream takes.
template <typename OffT>
__global__ void csr_slice_scan_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ col,
                                      uint32_t * __restrict__ flags, OffT * __restrict__ clen)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // row (blockDim is a multiple of 32)
    if (i - lane >= rows) return;
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    uint32_t fl = 0;
    int64_t n = 0;
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        const int c = active ? col[lo + l] : 0;
        const int first = __ffs(mask) - 1;
        const int base = __shfl_sync(0xffffffffu, c, first) - first;  // column lane 0 would hold
        const bool fits = !active || c == base + lane;
        const bool run = l < 32 && __popc(mask) >= 2 && __all_sync(0xffffffffu, fits);
        if (run) fl |= 1u << l;
        n += run ? 1 : __popc(mask);
    }
    if (lane == 0) {
        flags[i >> 5] = fl;
        clen[i >> 5] = (OffT)n;
    }
}

// One warp per slice: copy the slice's entries from row-major to slot-major order.  With `flags` the column indices go
// to the compressed stream (one int32 = the column of lane 0 for a run slot), at the slice's offset cofs[slice].
template <typename OffT>
__global__ void csr_slice_fill_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ col,
                                      const double * __restrict__ val, int32_t * __restrict__ scol, double * __restrict__ sval,
                                      const uint32_t * __restrict__ flags, const OffT * __restrict__ cofs)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // row (blockDim is a multiple of 32)
    if (i - lane >= rows) return;
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    int64_t pos = __shfl_sync(0xffffffffu, lo, 0);
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    uint32_t fl = flags ? flags[i >> 5] : 0u;
    int64_t cpos = flags ? (int64_t)cofs[i >> 5] : pos;
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        const bool run = fl & 1u;
        fl >>= 1;
        if (active) {
            sval[pos + rank] = val[lo + l];
            const int c = col[lo + l];
            if (!run) scol[cpos + rank] = c;
            else if (rank == 0) scol[cpos] = c - lane;
        }
        pos += __popc(mask);
        cpos += run ? 1 : __popc(mask);
    }
}

// The inverse of csr_slice_fill_kernel.
template <typename OffT>
__global__ void csr_slice_unfill_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ scol,
                                        const double * __restrict__ sval, int32_t * __restrict__ col, double * __restrict__ val,
                                        const uint32_t * __restrict__ flags, const OffT * __restrict__ cofs)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i - lane >= rows) return;
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    int64_t pos = __shfl_sync(0xffffffffu, lo, 0);
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    uint32_t fl = flags ? flags[i >> 5] : 0u;
    int64_t cpos = flags ? (int64_t)cofs[i >> 5] : pos;
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        const bool run = fl & 1u;
        fl >>= 1;
        if (active) {
            col[lo + l] = run ? scol[cpos] + lane : scol[cpos + rank];
            val[lo + l] = sval[pos + rank];
        }
        pos += __popc(mask);
        cpos += run ? 1 : __popc(mask);
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
    if (m->off64) csr_slice_unfill_kernel<int64_t><<<grid, 128, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, m->slice_col, m->slice_val, m->col, m->val, m->slice_flags, (const int64_t *)m->slice_cofs);
    else csr_slice_unfill_kernel<uint32_t><<<grid, 128, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, m->slice_col, m->slice_val, m->col, m->val, m->slice_flags, (const uint32_t *)m->slice_cofs);
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

template <typename OffT, int U, int THREADS, bool PUSH = false, bool RUNS = false, int REGS = (U <= 4 ? 32 : 64)>
__global__ void __launch_bounds__(THREADS, 65536 / REGS / THREADS)
csr_sliced_kernel(int64_t row0, int64_t rows, int independent, int store, double alpha, const OffT * __restrict__ rp,
                  const int32_t * __restrict__ scol, const double * __restrict__ sval, const double * __restrict__ x,
                  double * __restrict__ y, const double * __restrict__ y_in_host, double * __restrict__ y_out_host,
                  double * push0, int64_t push0_lo, int64_t push0_hi, double * push1, int64_t push1_lo, int64_t push1_hi,
                  const uint32_t * __restrict__ sflags, const OffT * __restrict__ scofs)
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
    // index runs: the slice's flag word (bit l: slot l is stored as ONE column, lane i holds column + i) and its place
    // in the compressed column stream; cq counts the int32 consumed so far
    uint32_t fl = 0;
    OffT cq = 0;  // position in the compressed column stream (32-bit whenever row_ptr is)
    if (RUNS) {
        fl = __ldg(sflags + (i >> 5));
        cq = __ldg(scofs + (i >> 5));
    }
    // The common case of a banded matrix: all 32 rows of the slice have the same length and every slot is a run.  Then
    // slot l's values are the 32 doubles at pos + 32 l, its column is stream[l] + lane, and ONE coalesced load brings the
    // whole column stream of the slice (<= 32 int32) into the warp: the gathers no longer wait for a column load, so the
    // value loads and the gathers of a batch are in flight together -- one memory round trip per batch instead of two.
    if (RUNS && maxlen <= 32 && fl == (0xffffffffu >> (32 - maxlen)) && __all_sync(0xffffffffu, len == maxlen)) {
        const int32_t cb = lane < maxlen ? __ldg(scol + (cq + (OffT)lane)) : 0;
        const double * sv = sval + pos + lane;
        for (int l0 = 0; l0 < maxlen; l0 += U) {
            double a[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = l0 + u < maxlen ? __ldg(sv + 32 * (l0 + u)) : 0.0;
            if (!waited) {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                waited = true;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int32_t base = __shfl_sync(0xffffffffu, cb, (l0 + u) & 31);
                xv[u] = l0 + u < maxlen ? ldx(x + base + lane) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (l0 + u < maxlen) z = __dadd_rn(z, __dmul_rn(a[u], xv[u]));
        }
    } else
    for (int l0 = 0; l0 < maxlen; l0 += U) {
        int c[U];
        double a[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {  // warp-uniform trip count: every lane takes part in the ballots
            const bool active = len > l0 + u;
            const unsigned mask = __ballot_sync(0xffffffffu, active);
            const int64_t p = pos + __popc(mask & below);
            if (RUNS) {
                const bool run = fl & 1u;  // warp-uniform
                fl >>= 1;
                const int32_t raw = active ? __ldg(scol + (cq + (OffT)(run ? 0 : __popc(mask & below)))) : 0;
                c[u] = run ? raw + lane : raw;
                a[u] = active ? __ldg(sval + p) : 0.0;
                pos += __popc(mask);
                cq += (OffT)(run ? 1 : __popc(mask));
                continue;
            }
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

template <typename OffT>
static int csr_scan_index_runs(Matrix * m, int64_t nslices, int64_t * total)
{
    OffT * cofs = (OffT *)m->slice_cofs;
    SPMV_CUDA(cudaMemsetAsync(cofs, 0, sizeof(OffT) * (size_t)(nslices + 1), m->stream));
    const unsigned grid = (unsigned)((m->rows + 127) / 128);
    csr_slice_scan_kernel<OffT><<<grid, 128, 0, m->stream>>>(m->rows, (const OffT *)m->rp, m->col, m->slice_flags, cofs);
    SPMV_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cofs, cofs, nslices + 1, m->stream));
    Scratch<unsigned char> tmp;
    SPMV_TRY(tmp.alloc((int64_t)tmp_bytes));
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cofs, cofs, nslices + 1, m->stream));
    OffT last = 0;
    SPMV_CUDA(cudaMemcpyAsync(&last, cofs + nslices, sizeof(OffT), cudaMemcpyDeviceToHost, m->stream));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    *total = (int64_t)last;
    return 0;
}

static void csr_free_index_runs(Matrix * m)
{
    const int64_t nslices = (m->rows + 31) / 32;
    if (m->slice_flags) { cudaFree(m->slice_flags); m->device_bytes -= 4 * nslices; }
    if (m->slice_cofs) { cudaFree(m->slice_cofs); m->device_bytes -= (m->off64 ? 8 : 4) * (nslices + 1); }
    m->slice_flags = nullptr;
    m->slice_cofs = nullptr;
    m->slice_runs = false;
}

static int csr_build_sliced(Matrix * m)
{
    if (m->slice_col && m->slice_val) return csr_drop_row_major(m);
    const int64_t nslices = (m->rows + 31) / 32;
    int64_t ccount = m->stored;
    int rc = alloc_streamed(m, &m->slice_val, m->stored);
    // index runs ("csr.index_runs": 0 auto, 1 always, -1 never): one int32 per (slice, slot) whose columns are base + lane
    if (rc == 0 && m->opt_csr_index_runs >= 0) {
        rc = dev_alloc(m, &m->slice_flags, nslices);
        if (rc == 0) {
            if (m->off64) rc = dev_alloc(m, (int64_t **)&m->slice_cofs, nslices + 1);
            else rc = dev_alloc(m, (uint32_t **)&m->slice_cofs, nslices + 1);
        }
        int64_t total = 0;
        if (rc == 0) rc = m->off64 ? csr_scan_index_runs<int64_t>(m, nslices, &total) : csr_scan_index_runs<uint32_t>(m, nslices, &total);
        if (rc == 0 && (m->opt_csr_index_runs >= 1 || 4 * total <= 3 * m->stored)) {
            m->slice_runs = true;
            ccount = total;
        } else {
            csr_free_index_runs(m);  // not worth it (or no room for the tables): the plain slot-major copy
            if (rc) { cudaGetLastError(); rc = 0; }
        }
    }
    if (rc == 0) rc = alloc_streamed(m, &m->slice_col, ccount);
    if (rc) {  // leave nothing half-built behind
        if (m->slice_col) cudaFree(m->slice_col);
        if (m->slice_val) cudaFree(m->slice_val);
        m->slice_col = nullptr;
        m->slice_val = nullptr;
        csr_free_index_runs(m);
        return rc;
    }
    m->slice_ccount = ccount;
    const unsigned grid = (unsigned)((m->rows + 127) / 128);
    if (m->off64) csr_slice_fill_kernel<int64_t><<<grid, 128, 0, m->stream>>>(m->rows, (const int64_t *)m->rp, m->col, m->val, m->slice_col, m->slice_val, m->slice_flags, (const int64_t *)m->slice_cofs);
    else csr_slice_fill_kernel<uint32_t><<<grid, 128, 0, m->stream>>>(m->rows, (const uint32_t *)m->rp, m->col, m->val, m->slice_col, m->slice_val, m->slice_flags, (const uint32_t *)m->slice_cofs);
    SPMV_CUDA(cudaGetLastError());
    m->aux_dirty = true;
    return csr_drop_row_major(m);
}

int csr_drop_sliced(Matrix * m)
{
    if (!m->slice_col && !m->slice_val) return 0;
    SPMV_TRY(csr_ensure_row_major(m));
    SPMV_CUDA(cudaStreamSynchronize(m->stream));
    stream_synced(m->stream);
    m->device_bytes -= (round_up(m->slice_ccount, 4096) + kPadEntries) * 4 + (round_up(m->stored, 4096) + kPadEntries) * 8;
    cudaFree(m->slice_col);
    cudaFree(m->slice_val);
    m->slice_col = nullptr;
    m->slice_val = nullptr;
    m->slice_ccount = 0;
    csr_free_index_runs(m);
    m->host_chunks = 0;  // the per-chunk column spans are recomputed from the next copy
    return 0;
}

// Largest column referenced by each of `chunks` equal row chunks (chunk = rows_per_chunk rows, a multiple of 32), from the
// slot-major copy: a slice's entries are contiguous, [rp[32 s], rp[32 s + 32)).  One warp per slice.
template <typename OffT>
__global__ void csr_chunk_colmax_kernel(int64_t rows, int64_t rows_per_chunk, const OffT * __restrict__ rp,
                                        const int32_t * __restrict__ scol, const uint32_t * __restrict__ flags,
                                        const OffT * __restrict__ cofs, int * __restrict__ colmax)
{
    const int lane = threadIdx.x & 31;
    const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r0 = slice * 32;
    if (r0 >= rows) return;
    int best = -1;
    if (!flags) {
        const int64_t lo = (int64_t)rp[r0], hi = (int64_t)rp[min(r0 + 32, rows)];
        for (int64_t k = lo + lane; k < hi; k += 32) best = max(best, __ldg(scol + k));
    } else {  // compressed column stream: walk the slots like the SpMV kernel does
        const int64_t i = r0 + lane;
        const int len = i < rows ? (int)min((int64_t)rp[i + 1] - (int64_t)rp[i], (int64_t)INT_MAX) : 0;
        const int maxlen = __reduce_max_sync(0xffffffffu, len);
        uint32_t fl = flags[slice];
        int64_t cpos = (int64_t)cofs[slice];
        for (int l = 0; l < maxlen; ++l) {
            const bool active = len > l;
            const unsigned mask = __ballot_sync(0xffffffffu, active);
            const bool run = fl & 1u;
            fl >>= 1;
            if (active) best = max(best, run ? __ldg(scol + cpos) + lane : __ldg(scol + cpos + __popc(mask & ((1u << lane) - 1u))));
            cpos += run ? 1 : __popc(mask);
        }
    }
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
    if (m->off64) csr_chunk_colmax_kernel<int64_t><<<grid, 256, 0, m->stream>>>(m->rows, rows_per_chunk, (const int64_t *)m->rp, m->slice_col, m->slice_flags, (const int64_t *)m->slice_cofs, d.p);
    else csr_chunk_colmax_kernel<uint32_t><<<grid, 256, 0, m->stream>>>(m->rows, rows_per_chunk, (const uint32_t *)m->rp, m->slice_col, m->slice_flags, (const uint32_t *)m->slice_cofs, d.p);
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
#define SPMV_SLICED_ARGS(OFF)                                                                                             \
    row0, row1, rm.independent, store, m->alpha, (const OFF *)m->rp, (const int32_t *)m->slice_col,                       \
        (const double *)m->slice_val, (const double *)m->x, m->y, (const double *)m->host_y_in, m->host_y_out,            \
        m->push_y[0], m->push_lo[0], m->push_hi[0], m->push_y[1], m->push_lo[1], m->push_hi[1],                            \
        (const uint32_t *)m->slice_flags, (const OFF *)m->slice_cofs
#define SPMV_SLICED(OFF, UU, TT)                                                                                          \
    do {                                                                                                                  \
        if (m->slice_runs && UU == 4 && TT == 128 && m->opt_csr_regs == 40)                                               \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, UU, TT, false, true, (UU == 4 && TT == 128 ? 40 : 32)>, (unsigned)grid, (unsigned)TT, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
        else if (m->slice_runs && UU == 4 && TT == 128 && m->opt_csr_regs == 48)                                          \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, UU, TT, false, true, (UU == 4 && TT == 128 ? 48 : 32)>, (unsigned)grid, (unsigned)TT, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
        else if (m->slice_runs)                                                                                           \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, UU, TT, false, true>, (unsigned)grid, (unsigned)TT, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
        else                                                                                                              \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, UU, TT, false, false>, (unsigned)grid, (unsigned)TT, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
    } while (0)
#define SPMV_SLICED_P(OFF)                                                                                                 \
    do {                                                                                                                  \
        if (m->slice_runs)                                                                                                \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, 4, 128, true, true>, (unsigned)grid, 128u, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
        else                                                                                                              \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, 4, 128, true, false>, (unsigned)grid, 128u, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
    } while (0)
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
#undef SPMV_SLICED_ARGS
    count_launch();
    return 0;
}

}  // namespace spmvb200
