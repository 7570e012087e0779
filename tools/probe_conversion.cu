// Exhaustive probe (all 2^32 fp32 bit patterns): where is the reference's eps floor
//     f(x) = logf(expf(x) + FLT_MIN)            (core.py:193-197, CUDA math library, no fast-math)
// the identity?  Prints, per sign and binade, how many inputs have f(x) != x (bitwise), so the
// kernels can skip the two transcendental calls on a range PROVEN to be the identity.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/probe_conversion tools/probe_conversion.cu
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(unsigned long long *fail) {   // fail[sign][exponent]
    const uint64_t n = 1ull << 32;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)i;
        const float x = __uint_as_float(b);
        const float y = logf(expf(x) + FLT_MIN);
        if (__float_as_uint(y) != b) atomicAdd(&fail[(b >> 31) * 256 + ((b >> 23) & 255)], 1ull);
    }
}

int main() {
    unsigned long long *d, h[512];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0, sizeof(h));
    probe<<<148 * 8, 256>>>(d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int s = 0; s < 2; ++s)
        for (int e = 100; e < 140; ++e)
            printf("%s binade [2^%d, 2^%d): %llu of 8388608 inputs differ\n", s ? "negative" : "positive",
                   e - 127, e - 126, h[s * 256 + e]);
    return cudaDeviceSynchronize() != cudaSuccess;
}
