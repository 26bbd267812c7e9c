// ptx.cuh -- the few sm_100a instructions the SpMV kernels need that CUDA C++ has no spelling for:
// mbarrier transactions, 1-D bulk-async copies (the TMA engine, UBLKCP in SASS) with an L2
// eviction policy, and non-allocating vector loads for streamed data.
#pragma once

#include <cstdint>

namespace spmvb200 {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void * p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Make barrier initialisation visible to the async proxy (the TMA engine).
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Order earlier generic-proxy accesses to shared memory before later async-proxy ones.
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t * bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t * bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
#ifdef SPMVB200_STREAM_EVICT_NORMAL  // experiment switch (build.py: SPMVB200_CFLAGS=-DSPMVB200_STREAM_EVICT_NORMAL)
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
#else
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}

__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// global -> shared bulk copy, completion counted in bytes on `bar`.
// dst, src 16 B aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// Streamed (read-once) vector loads: read-only path, no L1 allocation, L2 policy `pol`
// (evict-first for matrix data).  On sm_100a the plain ".L2::evict_first" qualifier is only
// accepted on 256-bit loads, so the narrower ones carry an explicit cache-hint policy.
__device__ __forceinline__ int4 ldg_stream_i4(const void * p, uint64_t pol)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int2 ldg_stream_i2(const void * p, uint64_t pol)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int ldg_stream_i1(const void * p, uint64_t pol)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double2 ldg_stream_d2(const void * p, uint64_t pol)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ldg_stream_d1(const void * p, uint64_t pol)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
// 256-bit load (sm_100a): four doubles, 32 B aligned.
__device__ __forceinline__ void ldg_stream_d4(const void * p, double (&v)[4])
{
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
    v[0] = __longlong_as_double((long long)a);
    v[1] = __longlong_as_double((long long)b);
    v[2] = __longlong_as_double((long long)c);
    v[3] = __longlong_as_double((long long)d);
}

// Gather of one x value: the read-only path (ld.global.nc, cached in L1).  PTX defines .nc only for data that nothing
// writes while the kernel runs, which under programmatic dependent launch includes the tail of the PREDECESSOR: the
// library therefore never launches with the PDL attribute when x may have been written by a kernel in flight
// (plan_run, abi.cu).  -DSPMVB200_X_COHERENT (experiment build, tools/pdl_probe.py) makes every gather an ordinary
// coherent load instead, to tell the two explanations of a wrong result under PDL apart.
__device__ __forceinline__ double ldx(const double * p)
{
#ifdef SPMVB200_X_COHERENT
    double r;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
#else
    return __ldg(p);
#endif
}

// Gather of one x value, by cache path (experiment switch of the COO kernels, "coo.xload"):
//   0 read-only path, cached in L1 (ld.global.nc = __ldg)      1 L2 only (ld.global.cg: no L1 line is allocated)
//   2 read-only path, L1::no_allocate                           3 read-only path, L1::evict_last
template <int PATH>
__device__ __forceinline__ double ld_x(const double * p)
{
    double r;
    if (PATH == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if (PATH == 2) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if (PATH == 3) asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else r = __ldg(p);
    return r;
}

// fp64 reduction into global memory without a return value (RED.E.ADD.F64).
__device__ __forceinline__ void red_add_f64(double * p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

}  // namespace ptx
}  // namespace spmvb200
