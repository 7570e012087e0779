// mbarrier / async-copy helpers shared by the tiled kernels (sm_90+ PTX).
#pragma once

#include <cstdint>

namespace torbi_b200 {

typedef unsigned long long u64;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive(u64 *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait until the phase with this parity has completed (a fresh barrier passes parity 1)
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Ring synchronisation of the tiled / streamed-split kernels.  Shipped build: mbarriers -- every thread's
// cp.async copies arrive on full[slot] when they land, every warp arrives on empty[slot] when it has read
// the stage, so warps may drift apart by up to a stage.  -DTORBI_RING_SYNCTHREADS (debug build of
// tools/sanitize.sh only) replaces all four by cp.async.wait_all + __syncthreads at the same program
// points: the same copy / read schedule with synchronisation compute-sanitizer's racecheck can model
// (it does not track non-bulk cp.async completion through mbarriers and reports those pairs as hazards).
#ifdef TORBI_RING_SYNCTHREADS
__device__ __forceinline__ void ring_full_arrive(u64 *) {}
__device__ __forceinline__ void ring_full_wait(u64 *, unsigned) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
}
__device__ __forceinline__ void ring_empty_arrive(u64 *) {}
__device__ __forceinline__ void ring_empty_wait(u64 *, unsigned) { __syncthreads(); }
#else
__device__ __forceinline__ void ring_full_arrive(u64 *bar) { cp_async_arrive(bar); }
__device__ __forceinline__ void ring_full_wait(u64 *bar, unsigned parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void ring_empty_arrive(u64 *bar) { mbar_arrive(bar); }
__device__ __forceinline__ void ring_empty_wait(u64 *bar, unsigned parity) { mbar_wait(bar, parity); }
#endif

}  // namespace torbi_b200
