// lto_cw_common.cuh -- pieces shared by the column-warp throughput kernels
// (lto_direct_cw.cu, lto_indirect_cw.cu): shared-memory mbarriers and branch-free
// reciprocal / reciprocal square root.
#pragma once
#include <cuda_runtime.h>

namespace lto {
namespace cwc {

// ---- mbarrier (shared::cta): producer/consumer hand-off between a state warp and the
// column warps without a CTA-wide barrier.  arrive has release, try_wait acquire semantics.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

// ---- TMA bulk store shared::cta -> global (one elected thread).  The generic-proxy writes of the staging buffer are
// made visible to the async proxy by fence_proxy_async() in every writing thread before the barrier that precedes it.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gptr, unsigned smem_addr, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gptr), "r"(smem_addr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the N most recent bulk stores of this thread have been read out of shared memory
template <int N> __device__ __forceinline__ void bulk_store_wait_read_but() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// try_wait with a suspend-time hint: the hardware parks the thread instead of spinning on the issue port
__device__ __forceinline__ void mbar_wait_parked(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAITP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONEP;\n"
        "bra LAB_WAITP;\n"
        "DONEP:\n"
        "}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}

// ---- Branch-free 1/sqrt(x) and 1/x for normal positive x: hardware seed (MUFU.RSQ64H /
// MUFU.RCP64H, ~2^-20) + one cubically convergent correction (error ~2^-60 before
// rounding).  No slow-path subroutine: keeps a state warp's dependent chain and its
// instruction footprint short.  NaN inputs propagate as NaN (reported through status[]).
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// u^(-1/16) for normal positive u, i.e. (sqrt u)^(-1/8): the step-size factor 0.9 * err^(-1/8) of the order-8 controller straight from
// the SQUARED scaled error, as three square roots of 1/sqrt(u), each one multiply by a reciprocal square root -- no IEEE
// sqrt / division subroutines on the state warp's chain.
__device__ __forceinline__ double inv_sixteenth_root(double u) {
    double z = fast_rsqrt(u);              // u^(-1/2)
    z = z * fast_rsqrt(z);                 // u^(-1/4)
    z = z * fast_rsqrt(z);                 // u^(-1/8)
    return z * fast_rsqrt(z);              // u^(-1/16)
}

}  // namespace cwc
}  // namespace lto
