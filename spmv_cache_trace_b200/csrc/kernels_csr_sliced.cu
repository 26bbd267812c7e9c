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
// INDEX RUNS ("diagonal" slices).  The rows of a banded matrix are shifted copies of each other: in a 32-row slice the
// entries lie on a few diagonals, i.e. the set of distinct offsets d = column - row is small.  For such a slice the
// builder lines the slots up by OFFSET instead of by position in the row: slot k holds the slice's k-th smallest offset,
// a 32-bit mask says which rows have an entry there, and the 32 column indices of the slot are ONE number -- lane i's
// column is base_k + i.  The column stream of the slice is then 8 B per slot {base, mask} -- 4 B, the base alone, when every
// slot holds all 32 rows ("dense": the interior of a grid line) -- instead of 4 B per entry, the kernel needs no row_ptr
// (the masks say which lanes are active) and no ballot, and because all descriptors of a slice arrive with one coalesced
// load, the gathers of x never wait for a column load: value loads and gathers of a batch are in flight together (one
// memory round trip per batch instead of two).  Rows at a grid boundary, which lack some neighbours, are simply holes in
// the masks, so every slice of a stencil qualifies whatever the grid's line length is.
// Values are not touched and a row's entries still arrive in ascending column order: every row is summed left to right
// from the same numbers, bit-identical to the reference.  A slice qualifies when no (row, column) pair occurs twice and
// its descriptors take at most 3/4 of the bytes of its explicit indices; other slices (R-MAT) keep explicit indices
// in the position-major order described above.  Per slice: a flag word (bit 31: diagonal form, bit 30: dense, low bits:
// slot count) and the offset of its part of the column stream, fused with the slice's value offset into one 16-byte record
// for the SpMV kernel.  27-point 512^3: the column stream shrinks from 14.4 GB to 0.5 GB and row_ptr is not read
// (33.6 GB moved per product instead of 49.2 GB, 7.40 -> 5.12 ms; profiles/r03_traffic_c5_csr.csv).  The algorithmic bytes of the metric
// keep counting 4 B per stored column index like the reference's csr_matrix::size(); the measured DRAM traffic is
// reported beside them.  Used when the whole stream shrinks to at most 3/4 ("csr.index_runs": 0 auto, 1 always, -1 never).
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <climits>

namespace spmvb200 {

using namespace ptx;

constexpr uint32_t kSliceDiagonal = 0x80000000u;  // flag word: the slice is stored in diagonal form, low 24 bits = slots
constexpr uint32_t kSliceDense = 0x40000000u;     // ... and every slot holds all 32 rows: the descriptors are bases only
constexpr uint32_t kSliceSlots = 0x00ffffffu;
constexpr int kNoOffset = INT_MAX;

// The offset (column - row) of a lane's next entry, or kNoOffset when its row is exhausted.
__device__ __forceinline__ int next_offset(const int32_t * __restrict__ col, int64_t lo, int k, int len, int64_t row)
{
    return k < len ? (int)((int64_t)col[lo + k] - row) : kNoOffset;
}

// One warp per slice: does it qualify for the diagonal form, how many slots, how many int32 of the column stream.
// The lanes' rows are merged by offset: every round takes the smallest pending offset and the lanes that hold it.
template <typename OffT>
__global__ void csr_slice_scan_kernel(int64_t rows, const OffT * __restrict__ rp, const int32_t * __restrict__ col,
                                      uint32_t * __restrict__ flags, OffT * __restrict__ clen)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // row (blockDim is a multiple of 32)
    if (i - lane >= rows) return;
    const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
    const int len = (int)min(hi - lo, (int64_t)INT_MAX);
    const int64_t entries = (i - lane + 32 <= rows ? (int64_t)rp[i - lane + 32] : (int64_t)rp[rows]) - (int64_t)rp[i - lane];
    const int64_t limit = min(entries * 3 / 8, (int64_t)0x00ffffff);  // 8 B per slot <= 3/4 of 4 B per entry
    int k = 0;
    int64_t slots = 0;
    bool ok = entries > 0, dense = true;
    while (ok) {
        const int d = next_offset(col, lo, k, len, i);
        const int dmin = __reduce_min_sync(0xffffffffu, d);
        if (dmin == kNoOffset) break;
        bool dup = false;
        if (d == dmin) {
            ++k;
            dup = next_offset(col, lo, k, len, i) == dmin;  // the same (row, column) twice: two entries for one slot
        }
        if (!__all_sync(0xffffffffu, d == dmin)) dense = false;
        if (++slots > limit || __any_sync(0xffffffffu, dup)) ok = false;
    }
    if (lane == 0) {
        flags[i >> 5] = ok ? (kSliceDiagonal | (dense ? kSliceDense : 0u) | (uint32_t)slots) : 0u;
        // even lengths: the {base, mask} descriptors are read as 8-byte pairs
        clen[i >> 5] = (OffT)(!ok ? ((entries + 1) & ~(int64_t)1) : dense ? ((slots + 1) & ~(int64_t)1) : 2 * slots);
    }
}

// One warp per slice: copy the slice's entries from row-major to slot-major order -- position-major with explicit column
// indices, or (flag word) offset-major with one {base, mask} descriptor per slot.  Without `flags` the column indices go
// where the values go (the plain copy: scol and sval are parallel arrays).
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
    const unsigned below = (1u << lane) - 1u;
    const uint32_t fl = flags ? flags[i >> 5] : 0u;
    int64_t cpos = flags ? (int64_t)cofs[i >> 5] : pos;
    if (fl & kSliceDiagonal) {
        int k = 0;
        for (;;) {
            const int d = next_offset(col, lo, k, len, i);
            const int dmin = __reduce_min_sync(0xffffffffu, d);
            if (dmin == kNoOffset) break;
            const bool active = d == dmin;
            const unsigned mask = __ballot_sync(0xffffffffu, active);
            if (active) sval[pos + __popc(mask & below)] = val[lo + k++];
            if (lane == 0) {
                scol[cpos] = (int32_t)((i + dmin));  // the column lane 0 would hold: lane j's column is this + j
                if (!(fl & kSliceDense)) scol[cpos + 1] = (int32_t)mask;
            }
            pos += __popc(mask);
            cpos += (fl & kSliceDense) ? 1 : 2;
        }
        return;
    }
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        const int rank = __popc(mask & below);
        if (active) {
            sval[pos + rank] = val[lo + l];
            scol[cpos + rank] = col[lo + l];
        }
        pos += __popc(mask);
        cpos += __popc(mask);
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
    const unsigned below = (1u << lane) - 1u;
    const uint32_t fl = flags ? flags[i >> 5] : 0u;
    int64_t cpos = flags ? (int64_t)cofs[i >> 5] : pos;
    if (fl & kSliceDiagonal) {
        const int slots = (int)(fl & kSliceSlots);
        int k = 0;
        for (int t = 0; t < slots; ++t) {
            const int base = (fl & kSliceDense) ? scol[cpos + t] : scol[cpos + 2 * t];
            const unsigned mask = (fl & kSliceDense) ? 0xffffffffu : (unsigned)scol[cpos + 2 * t + 1];
            if ((mask >> lane) & 1u) {
                col[lo + k] = base + lane;
                val[lo + k] = sval[pos + __popc(mask & below)];
                ++k;
            }
            pos += __popc(mask);
        }
        return;
    }
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    for (int l = 0; l < maxlen; ++l) {
        const bool active = len > l;
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        const int rank = __popc(mask & below);
        if (active) {
            col[lo + l] = scol[cpos + rank];
            val[lo + l] = sval[pos + rank];
        }
        pos += __popc(mask);
        cpos += __popc(mask);
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

// What the SpMV kernel needs to know about a slice, in ONE 16- or 32-byte record (one broadcast load at the start of a warp's
// life instead of three dependent ones): the flag word, the slice's place in the column stream, and its place in the values
// (= row_ptr of its first row).
template <typename OffT>
struct SliceMeta {
    OffT flags, cofs, vofs, pad;
};

__device__ __forceinline__ void load_slice_meta(const SliceMeta<uint32_t> * p, uint32_t & fl, int64_t & cofs, int64_t & vofs)
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    fl = v.x; cofs = v.y; vofs = v.z;
}
__device__ __forceinline__ void load_slice_meta(const SliceMeta<int64_t> * p, uint32_t & fl, int64_t & cofs, int64_t & vofs)
{
    const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(p)), b = __ldg(reinterpret_cast<const longlong2 *>(p) + 1);
    fl = (uint32_t)a.x; cofs = a.y; vofs = b.x;
}

template <typename OffT>
__global__ void csr_slice_meta_kernel(int64_t nslices, const OffT * __restrict__ rp, const uint32_t * __restrict__ flags,
                                      const OffT * __restrict__ cofs, SliceMeta<OffT> * __restrict__ meta)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nslices) meta[s] = SliceMeta<OffT>{(OffT)flags[s], cofs[s], rp[32 * s], 0};
}

template <typename OffT, int U, int THREADS, bool PUSH = false, bool RUNS = false, int REGS = (U <= 4 ? 32 : 64)>
__global__ void __launch_bounds__(THREADS, 65536 / REGS / THREADS)
csr_sliced_kernel(int64_t row0, int64_t rows, int independent, int store, double alpha, const OffT * __restrict__ rp,
                  const int32_t * __restrict__ scol, const double * __restrict__ sval, const double * __restrict__ x,
                  double * __restrict__ y, const double * __restrict__ y_in_host, double * __restrict__ y_out_host,
                  double * push0, int64_t push0_lo, int64_t push0_hi, double * push1, int64_t push1_lo, int64_t push1_hi,
                  const SliceMeta<OffT> * __restrict__ smeta)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    // rows [row0, rows) of the matrix (row0 a multiple of 32: a warp is a slice)
    const int64_t i = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i - lane >= rows) return;  // whole warp past the end
    const unsigned below = (1u << lane) - 1u;
    double z = 0.0;
    bool waited = independent != 0;
    int len = 0;  // > 0: the row has entries (diagonal form: 1 stands for "some")
    uint32_t fl = 0;
    int64_t cofs = 0, vofs = 0;
    if (RUNS) load_slice_meta(smeta + (i >> 5), fl, cofs, vofs);
    if (RUNS && (fl & kSliceDiagonal)) {
        // Diagonal form: slot k of the slice = its k-th smallest offset (column - row); descriptor {base, mask}: lane j has an
        // entry iff bit j of mask is set, and its column is base + j.  Values in slot order, active lanes ascending.
        const int slots = (int)(fl & kSliceSlots);
        const double * sv = sval + vofs;  // the slice's entries start at row_ptr of its first row
        unsigned seen = 0;
        if (fl & kSliceDense) {
            // every slot holds all 32 rows (the interior of a grid line): descriptors are bases only, slot t's values are
            // the 32 doubles at 32 t
            const int32_t * __restrict__ bases = scol + cofs;
            sv += lane;
            for (int t0 = 0; t0 < slots; t0 += 32) {
                const int32_t mine = t0 + lane < slots ? __ldg(bases + t0 + lane) : 0;
                const int n = min(32, slots - t0);
                for (int l0 = 0; l0 < n; l0 += U) {
                    double a[U], xv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) a[u] = l0 + u < n ? __ldg(sv + 32 * (l0 + u)) : 0.0;
                    if (!waited) {
                        asm volatile("griddepcontrol.wait;" ::: "memory");
                        waited = true;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int32_t base = __shfl_sync(0xffffffffu, mine, (l0 + u) & 31);
                        xv[u] = l0 + u < n ? ldx(x + base + lane) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (l0 + u < n) z = __dadd_rn(z, __dmul_rn(a[u], xv[u]));
                }
                sv += 32 * n;
            }
            seen = 1;
        } else {
        const int2 * __restrict__ desc = reinterpret_cast<const int2 *>(scol + cofs);
        for (int t0 = 0; t0 < slots; t0 += 32) {
            const int2 mine = t0 + lane < slots ? __ldg(desc + t0 + lane) : make_int2(0, 0);  // all descriptors: one coalesced load
            const int n = min(32, slots - t0);
            for (int l0 = 0; l0 < n; l0 += U) {
                double a[U], xv[U];
                int c[U];
                unsigned act = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {  // slots beyond n read descriptor {0, 0} of an idle lane or a real one: masked off
                    const unsigned mask = l0 + u < n ? (unsigned)__shfl_sync(0xffffffffu, mine.y, (l0 + u) & 31) : 0u;
                    c[u] = __shfl_sync(0xffffffffu, mine.x, (l0 + u) & 31) + lane;
                    const bool active = (mask >> lane) & 1u;
                    act |= (active ? 1u : 0u) << u;
                    a[u] = active ? __ldg(sv + __popc(mask & below)) : 0.0;
                    sv += __popc(mask);
                }
                if (!waited) {  // the matrix is immutable; x and y may come from the previous launch
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    waited = true;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) xv[u] = (act >> u) & 1u ? ldx(x + c[u]) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if ((act >> u) & 1u) z = __dadd_rn(z, __dmul_rn(a[u], xv[u]));
                seen |= act;
            }
        }
        }
        len = seen ? 1 : 0;
    } else {
        const int64_t lo = i < rows ? (int64_t)rp[i] : 0, hi = i < rows ? (int64_t)rp[i + 1] : 0;
        len = (int)min(hi - lo, (int64_t)INT_MAX);
        int64_t pos = __shfl_sync(0xffffffffu, lo, 0);  // offset of the slice = row_ptr of its first row
        // with index runs the explicit columns of this slice sit at its offset in the column stream, else beside the values
        int64_t cdelta = 0;
        if (RUNS) cdelta = cofs - pos;
        const int maxlen = __reduce_max_sync(0xffffffffu, len);
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
                c[u] = active ? __ldg(scol + (RUNS ? p + cdelta : p)) : 0;
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
    if (m->slice_meta) { cudaFree(m->slice_meta); m->device_bytes -= (m->off64 ? 32 : 16) * nslices; }
    m->slice_flags = nullptr;
    m->slice_cofs = nullptr;
    m->slice_meta = nullptr;
    m->slice_runs = false;
}

static int csr_build_sliced(Matrix * m)
{
    if (m->slice_col && m->slice_val) return csr_drop_row_major(m);
    const int64_t nslices = (m->rows + 31) / 32;
    int64_t ccount = m->stored;
    int rc = alloc_streamed(m, &m->slice_val, m->stored);
    // index runs ("csr.index_runs": 0 auto, 1 always, -1 never): one int32 per (slice, slot) whose columns are base + lane
    // (32-bit stream offsets must hold stored + one pad per slice in the worst case)
    const bool offsets_fit = m->off64 || m->stored + nslices < ((int64_t)1 << 32);
    if (rc == 0 && m->opt_csr_index_runs >= 0 && offsets_fit) {
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
    if (m->slice_runs) {  // the fused per-slice records the SpMV kernel reads
        const size_t bytes = (m->off64 ? 32 : 16) * (size_t)nslices;
        if (cudaMalloc(&m->slice_meta, bytes ? bytes : 16) != cudaSuccess) {
            cudaGetLastError();
            return fail(SPMVB200_ERR_NOMEM, "out of device memory for the slice records");
        }
        m->device_bytes += (int64_t)bytes;
        const unsigned g2 = (unsigned)((nslices + 255) / 256);
        if (m->off64) csr_slice_meta_kernel<int64_t><<<g2, 256, 0, m->stream>>>(nslices, (const int64_t *)m->rp, m->slice_flags, (const int64_t *)m->slice_cofs, (SliceMeta<int64_t> *)m->slice_meta);
        else csr_slice_meta_kernel<uint32_t><<<g2, 256, 0, m->stream>>>(nslices, (const uint32_t *)m->rp, m->slice_flags, (const uint32_t *)m->slice_cofs, (SliceMeta<uint32_t> *)m->slice_meta);
        SPMV_CUDA(cudaGetLastError());
    }
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
    } else if (flags[slice] & kSliceDiagonal) {  // descriptors {base, mask}: the largest column of a slot is base + highest lane
        const int slots = (int)(flags[slice] & kSliceSlots);
        const bool dense = flags[slice] & kSliceDense;
        const int64_t cpos = (int64_t)cofs[slice];
        for (int t = lane; t < slots; t += 32) {
            const unsigned mask = dense ? 0xffffffffu : (unsigned)__ldg(scol + cpos + 2 * t + 1);
            if (mask) best = max(best, __ldg(scol + cpos + (dense ? t : 2 * t)) + 31 - __clz(mask));
        }
    } else {  // explicit columns at the slice's place in the column stream
        const int64_t n = (int64_t)rp[min(r0 + 32, rows)] - (int64_t)rp[r0], cpos = (int64_t)cofs[slice];
        for (int64_t k = lane; k < n; k += 32) best = max(best, __ldg(scol + cpos + k));
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
        (const SliceMeta<OFF> *)m->slice_meta
#define SPMV_SLICED(OFF, UU, TT)                                                                                          \
    do {                                                                                                                  \
        if (m->slice_runs && UU == 4 && TT == 128 && m->opt_csr_regs == 40)                                               \
            SPMV_CUDA(launch_kernel(csr_sliced_kernel<OFF, UU, TT, false, true, (UU == 4 && TT == 128 ? 40 : 32)>, (unsigned)grid, (unsigned)TT, 0, m->stream, rm.pdl, SPMV_SLICED_ARGS(OFF))); \
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
