// Device self-tests exported through the C ABI (used by tests/, not by the decode path).
#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

namespace {

// every fp32 bit pattern: the shortcut form of the eps floor against the plain formula
__global__ void selftest_conversion_kernel(unsigned long long *mismatches) {
    const uint64_t n = 1ull << 32;
    unsigned long long bad = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)i);
        if (__float_as_uint(eps_floor(x)) != __float_as_uint(eps_floor_reference(x))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace

cudaError_t launch_selftest_conversion(unsigned long long *mismatches, int sm_count, cudaStream_t stream) {
    selftest_conversion_kernel<<<sm_count * 8, 256, 0, stream>>>(mismatches);
    return cudaGetLastError();
}

}  // namespace torbi_b200
