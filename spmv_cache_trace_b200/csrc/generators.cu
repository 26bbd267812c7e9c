// generators.cu -- synthetic matrices of BASELINE.json, generated directly in device memory.
//
// The reference has no generators (its inputs are Matrix Market files); these exist because the
// large configurations (2^31 and 3.6e9 non-zeros) exceed both the reference's int32 sizes and host
// RAM.  Both generators are deterministic functions of their parameters and have a numpy
// restatement in tests/generators_ref.py that the GPU tests compare against bit for bit.
//
//   stencil: rows [row_begin, row_end) of the 5-point (2D), 7-point or 27-point (3D) operator on an
//            nx*ny*nz grid, x fastest; diagonal 4 / 6 / 26, neighbours -1; columns ascending.
//   rmat:    edge_factor * 2^scale draws; at each of `scale` levels the quadrant is chosen from
//            u = (splitmix64(seed + e*G + level*H) >> 11) * 2^-53 against (a, a+b, a+b+c);
//            duplicates removed; value(i, j) = 2*((splitmix64(key ^ V) >> 11) * 2^-53) - 1.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

namespace spmvb200 {

int store_offsets(Matrix * m, const int64_t * d_rp64, int64_t rows, int64_t stored);

// ---------------------------------------------------------------------------------------------
// stencils
// ---------------------------------------------------------------------------------------------

struct Grid {
    int64_t nx, ny, nz;
    int kind;
};

__device__ __forceinline__ int span(int64_t i, int64_t n) { return 1 + (i > 0) + (i + 1 < n); }

__device__ __forceinline__ int stencil_count(const Grid g, int64_t r)
{
    const int64_t ix = r % g.nx, iy = (r / g.nx) % g.ny, iz = r / (g.nx * g.ny);
    const int cx = span(ix, g.nx), cy = span(iy, g.ny), cz = span(iz, g.nz);
    if (g.kind == SPMVB200_STENCIL_3D27) return cx * cy * cz;
    if (g.kind == SPMVB200_STENCIL_3D7) return 1 + (cx - 1) + (cy - 1) + (cz - 1);
    return 1 + (cx - 1) + (cy - 1);
}

__global__ void stencil_count_kernel(Grid g, int64_t rb, int64_t nrows, int64_t * counts)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= nrows; t += (int64_t)gridDim.x * blockDim.x)
        counts[t] = t < nrows ? stencil_count(g, rb + t) : 0;
}

__global__ void stencil_fill_kernel(Grid g, int64_t rb, int64_t nrows, const int64_t * rp, int32_t * col, double * val)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = rb + t;
        const int64_t ix = r % g.nx, iy = (r / g.nx) % g.ny, iz = r / (g.nx * g.ny);
        int64_t k = rp[t];
        if (g.kind == SPMVB200_STENCIL_3D27) {
            for (int dz = -1; dz <= 1; ++dz) {
                if (iz + dz < 0 || iz + dz >= g.nz) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    if (iy + dy < 0 || iy + dy >= g.ny) continue;
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (ix + dx < 0 || ix + dx >= g.nx) continue;
                        col[k] = (int32_t)(r + (dz * g.ny + dy) * g.nx + dx);
                        val[k] = (dx == 0 && dy == 0 && dz == 0) ? 26.0 : -1.0;
                        ++k;
                    }
                }
            }
        } else {
            const bool three = g.kind == SPMVB200_STENCIL_3D7;
            if (three && iz > 0) { col[k] = (int32_t)(r - g.nx * g.ny); val[k++] = -1.0; }
            if (iy > 0) { col[k] = (int32_t)(r - g.nx); val[k++] = -1.0; }
            if (ix > 0) { col[k] = (int32_t)(r - 1); val[k++] = -1.0; }
            col[k] = (int32_t)r; val[k++] = three ? 6.0 : 4.0;
            if (ix + 1 < g.nx) { col[k] = (int32_t)(r + 1); val[k++] = -1.0; }
            if (iy + 1 < g.ny) { col[k] = (int32_t)(r + g.nx); val[k++] = -1.0; }
            if (three && iz + 1 < g.nz) { col[k] = (int32_t)(r + g.nx * g.ny); val[k++] = -1.0; }
        }
    }
}

int gen_stencil(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t rb, int64_t re, Matrix * m)
{
    if (kind < 0 || kind > 2) return fail(SPMVB200_ERR_INVALID, "unknown stencil kind");
    if (kind == SPMVB200_STENCIL_2D5 && nz != 1) return fail(SPMVB200_ERR_INVALID, "2D stencil needs nz == 1");
    cudaStream_t s = m->stream;
    const Grid g{nx, ny, nz, kind};
    const int64_t nrows = re - rb;
    Scratch<int64_t> counts, rp64;
    Scratch<unsigned char> tmp;
    SPMV_TRY(counts.alloc(nrows + 1)); SPMV_TRY(rp64.alloc(nrows + 1));
    stencil_count_kernel<<<grid_for(nrows + 1), 256, 0, s>>>(g, rb, nrows, counts.p);
    SPMV_CUDA(cudaGetLastError());
    size_t tb = 0;
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, rp64.p, nrows + 1, s));
    SPMV_TRY(tmp.alloc((int64_t)tb));
    SPMV_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, rp64.p, nrows + 1, s));
    int64_t nnz = 0;
    SPMV_CUDA(cudaMemcpyAsync(&nnz, rp64.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    if (cudaFree(counts.release()) != cudaSuccess) return fail(SPMVB200_ERR_CUDA, "cudaFree");

    m->format = SPMVB200_CSR;
    m->rows = nrows; m->cols = nx * ny * nz; m->nnz = nnz; m->stored = nnz; m->row_offset = rb;
    SPMV_TRY(alloc_streamed(m, &m->col, nnz));
    SPMV_TRY(alloc_streamed(m, &m->val, nnz));
    if (nrows > 0) {
        stencil_fill_kernel<<<grid_for(nrows), 256, 0, s>>>(g, rb, nrows, rp64.p, m->col, m->val);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_TRY(store_offsets(m, rp64.p, nrows, nnz));
    SPMV_TRY(csr_build_tiles(m));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// R-MAT
// ---------------------------------------------------------------------------------------------

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void rmat_edges_kernel(int64_t m_edges, int scale, uint64_t seed, double ta, double tb, double tc,
                                  uint64_t * keys)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m_edges; e += (int64_t)gridDim.x * blockDim.x) {
        uint64_t row = 0, col = 0;
        for (int level = 0; level < scale; ++level) {
            const uint64_t h = splitmix64(seed + (uint64_t)e * 0x9E3779B97F4A7C15ull + (uint64_t)level * 0xBF58476D1CE4E5B9ull);
            const double u = (double)(h >> 11) * 0x1.0p-53;
            const int q = u < ta ? 0 : (u < tb ? 1 : (u < tc ? 2 : 3));
            row = (row << 1) | (uint64_t)(q >> 1);
            col = (col << 1) | (uint64_t)(q & 1);
        }
        keys[e] = (row << 32) | col;
    }
}

__global__ void key_lower_bound_kernel(int64_t rb, int64_t nrows, int64_t n, const uint64_t * keys, int64_t * rp)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= nrows; t += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t target = (uint64_t)(rb + t) << 32;
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1; else hi = mid;
        }
        rp[t] = lo;
    }
}

__global__ void rebase_kernel(int64_t n, int64_t * rp, int64_t first)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        rp[t] -= first;
}

__global__ void rmat_fill_kernel(int64_t n, const uint64_t * keys, int32_t * col, double * val)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[k];
        col[k] = (int32_t)(uint32_t)key;
        const uint64_t h = splitmix64(key ^ 0xD1B54A32D192ED03ull);
        val[k] = 2.0 * ((double)(h >> 11) * 0x1.0p-53) - 1.0;
    }
}

int gen_rmat(int scale, int edge_factor, uint64_t seed, double a, double b, double c, int64_t rb, int64_t re,
             Matrix * m)
{
    cudaStream_t s = m->stream;
    const int64_t n_rows_total = (int64_t)1 << scale;
    const int64_t m_edges = (int64_t)edge_factor << scale;
    const double ta = a, tb = a + b, tc = a + b + c;
    Scratch<uint64_t> keys, sorted;
    Scratch<unsigned char> tmp;
    Scratch<int64_t> nsel, rp64;
    SPMV_TRY(keys.alloc(m_edges)); SPMV_TRY(sorted.alloc(m_edges)); SPMV_TRY(nsel.alloc(1));
    rmat_edges_kernel<<<grid_for(m_edges, m->sm_count * 2), 256, 0, s>>>(m_edges, scale, seed, ta, tb, tc, keys.p);
    SPMV_CUDA(cudaGetLastError());
    size_t tb1 = 0, tb2 = 0;
    SPMV_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb1, keys.p, sorted.p, m_edges, 0, 32 + scale, s));
    SPMV_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, sorted.p, keys.p, nsel.p, m_edges, s));
    SPMV_TRY(tmp.alloc((int64_t)std::max(tb1, tb2)));
    SPMV_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb1, keys.p, sorted.p, m_edges, 0, 32 + scale, s));
    SPMV_CUDA(cub::DeviceSelect::Unique(tmp.p, tb2, sorted.p, keys.p, nsel.p, m_edges, s));
    int64_t n_unique = 0;
    SPMV_CUDA(cudaMemcpyAsync(&n_unique, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    cudaFree(sorted.release());
    cudaFree(tmp.release());

    const int64_t nrows = re - rb;
    SPMV_TRY(rp64.alloc(nrows + 1));
    key_lower_bound_kernel<<<grid_for(nrows + 1), 256, 0, s>>>(rb, nrows, n_unique, keys.p, rp64.p);
    SPMV_CUDA(cudaGetLastError());
    int64_t first = 0, last = 0;
    SPMV_CUDA(cudaMemcpyAsync(&first, rp64.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaMemcpyAsync(&last, rp64.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SPMV_CUDA(cudaStreamSynchronize(s));
    const int64_t nnz = last - first;
    if (first) rebase_kernel<<<grid_for(nrows + 1), 256, 0, s>>>(nrows + 1, rp64.p, first);

    m->format = SPMVB200_CSR;
    m->rows = nrows; m->cols = n_rows_total; m->nnz = nnz; m->stored = nnz; m->row_offset = rb;
    SPMV_TRY(alloc_streamed(m, &m->col, nnz));
    SPMV_TRY(alloc_streamed(m, &m->val, nnz));
    if (nnz > 0) {
        rmat_fill_kernel<<<grid_for(nnz, m->sm_count * 2), 256, 0, s>>>(nnz, keys.p + first, m->col, m->val);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_TRY(store_offsets(m, rp64.p, nrows, nnz));
    SPMV_TRY(csr_build_tiles(m));
    SPMV_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace spmvb200
