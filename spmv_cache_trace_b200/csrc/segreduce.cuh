// segreduce.cuh -- y[row] += sum of the products of a run of equal row indices, for a warp that
// holds 32*E consecutive entries, E per lane (lane l: entries E*l .. E*l+E-1).
//
// Each lane sums its own runs serially.  Runs that lie strictly inside a lane are added to y at
// once; the lane's first and last run may continue in the neighbouring lanes, so the sums of the
// last runs go through ONE segmented inclusive scan across the warp (Kogge-Stone with shuffles, cut
// short at the longest chain of single-run lanes), after which every run that ends in a lane is
// added to y by that lane with one fp64 reduction.  Rows < 0 mark entries past the end.
// Used by coo_warp4_kernel (kernels_coo.cu) and csr_flat_kernel (kernels_csr_flat.cu).
#pragma once

#include "ptx.cuh"

namespace spmvb200 {

// alpha scales every sum on its way out (alpha = 1.0 is exact).  E = entries per lane.
// rowmap (optional): the row ids in r are indices into it (the flat CSR kernel numbers only the non-empty
// rows); it is consulted only where a sum leaves for y.
template <int E>
__device__ __forceinline__ void warp_segmented_add(int lane, const int (&r)[E], const double (&p)[E],
                                                   double * __restrict__ y, double alpha = 1.0,
                                                   const int32_t * __restrict__ rowmap = nullptr)
{
    auto red_add_f64 = [&](int row, double v) { ptx::red_add_f64(y + (rowmap ? __ldg(rowmap + row) : row), v); };
    int cur_row = r[0];
    double cur = p[0], head = 0.0;
    bool single = true;  // the lane holds one run only
#pragma unroll
    for (int j = 1; j < E; ++j) {
        if (r[j] == cur_row) {
            cur = __dadd_rn(cur, p[j]);
        } else {
            if (single) { head = cur; single = false; }
            else if (cur_row >= 0) red_add_f64(cur_row, __dmul_rn(alpha, cur));
            cur_row = r[j];
            cur = p[j];
        }
    }
    // across lanes: my first run may continue the previous lane's last run
    const int prev_last = __shfl_up_sync(0xffffffffu, r[E - 1], 1);
    const int next_first = __shfl_down_sync(0xffffffffu, r[0], 1);
    const bool cont = lane > 0 && prev_last == r[0];
    const bool next_cont = lane < 31 && next_first == r[E - 1];
    const unsigned heads = __ballot_sync(0xffffffffu, !(single && cont));
    const int dist = lane - (31 - __clz(heads & (0xffffffffu >> (31 - lane))));
    const int longest = __reduce_max_sync(0xffffffffu, dist);
    double s = cur;  // sum of the last run, then of the chain of single-run lanes that ends here
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        if (d <= longest) {  // warp-uniform
            const double t = __shfl_up_sync(0xffffffffu, s, d);
            if (dist >= d) s = __dadd_rn(s, t);
        }
    }
    const double prev_s = __shfl_up_sync(0xffffffffu, s, 1);
    if (!single && r[0] >= 0) red_add_f64(r[0], __dmul_rn(alpha, cont ? __dadd_rn(prev_s, head) : head));
    if (!next_cont && r[E - 1] >= 0) red_add_f64(r[E - 1], __dmul_rn(alpha, s));
}

__device__ __forceinline__ void warp_segmented_add4(int lane, const int (&r)[4], const double (&p)[4],
                                                    double * __restrict__ y, double alpha = 1.0)
{
    warp_segmented_add<4>(lane, r, p, y, alpha);
}

}  // namespace spmvb200
