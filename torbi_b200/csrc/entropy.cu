// Framewise normalised entropy for the chunked long-sequence mode: restates `entropy` of
// /root/reference/torbi/chunk.py:81-85 on the device,
//     H[t] = -( sum_s exp(x[t][s]) * x[t][s] ) / log(states),
// applied -- like the reference's Dataset (data/dataset.py:21-23) -- to the observation exactly as
// stored (log-probabilities in the intended use).  The reference evaluates it with torch CPU ops on
// the (states, frames) transpose and then walks the frames in a Python loop (chunk.py:66-76).
//
// One warp per frame, one pass over HBM (4 * states bytes per frame, read once: HBM-bound).  The
// sum runs in a FIXED order (lane-strided float4 groups ascending, then a butterfly across lanes),
// so the result is deterministic; it is not bit-identical to torch's CPU reduction, whose order is
// an implementation detail of the vectorised kernel.  Exactness of the split DECISIONS is kept one
// level up: alongside H the kernel returns A[t] = sum_s |exp(x) * x| / log(states), which bounds the
// rounding difference between any two fp32 summation orders (<= ~2 * states * 2^-24 * A), and the
// host re-evaluates with the reference's own ops the few frames whose H lies within that margin of
// the threshold (torbi_b200/chunk.py).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

namespace {

constexpr int kEntropyThreads = 256;

template <int kDtype>
__global__ void __launch_bounds__(kEntropyThreads)
entropy_kernel(const void *__restrict__ observation, int64_t rows, int S, float log_states,
               float2 *__restrict__ out) {
    const int lane = threadIdx.x % kWarp;
    const int64_t warps = (int64_t)gridDim.x * (kEntropyThreads / kWarp);
    const bool vec = (S % 4) == 0 && (reinterpret_cast<uintptr_t>(observation) & 15) == 0;
    for (int64_t t = (int64_t)blockIdx.x * (kEntropyThreads / kWarp) + threadIdx.x / kWarp; t < rows; t += warps) {
        float sum = 0.f, mag = 0.f;
        auto term = [&](float x) {
            // exp(x) * x, as chunk.py:83 (x = -inf gives 0 * -inf = NaN there too; outside the intended domain)
            const float v = expf(x) * x;
            sum += v;
            mag += fabsf(v);
        };
        if (kDtype == 0 && vec) {
            const float4 *row = reinterpret_cast<const float4 *>(static_cast<const float *>(observation) + t * S);
            for (int q = lane; q < S / 4; q += kWarp) {
                const float4 x = __ldcs(row + q);            // read once: streaming
                term(x.x); term(x.y); term(x.z); term(x.w);
            }
        } else if (kDtype != 0 && vec) {
            const uint2 *row = reinterpret_cast<const uint2 *>(static_cast<const unsigned short *>(observation) + t * S);
            for (int q = lane; q < S / 4; q += kWarp) {
                const uint2 raw = __ldcs(row + q);
                term(half_bits_to_float((unsigned short)(raw.x & 0xffffu), kDtype));
                term(half_bits_to_float((unsigned short)(raw.x >> 16), kDtype));
                term(half_bits_to_float((unsigned short)(raw.y & 0xffffu), kDtype));
                term(half_bits_to_float((unsigned short)(raw.y >> 16), kDtype));
            }
        } else {
            for (int s = lane; s < S; s += kWarp) term(load_observation(observation, t * S + s, kDtype));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(kFullMask, sum, off);
            mag += __shfl_xor_sync(kFullMask, mag, off);
        }
        if (lane == 0) out[t] = make_float2(-(sum / log_states), mag / log_states);
    }
}

}  // namespace

cudaError_t launch_entropy(const void *observation, int obs_dtype, int64_t rows, int64_t S, float *out,
                           int sm_count, cudaStream_t stream) {
    // the reference divides by torch.log(torch.tensor(states)): the float32 log of the integer (chunk.py:85)
    const float log_states = logf((float)S);
    const int64_t blocks_needed = (rows + kEntropyThreads / kWarp - 1) / (kEntropyThreads / kWarp);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(blocks_needed, (int64_t)sm_count * 8));
    float2 *o = reinterpret_cast<float2 *>(out);
    if (obs_dtype == 0)
        entropy_kernel<0><<<grid, kEntropyThreads, 0, stream>>>(observation, rows, (int)S, log_states, o);
    else if (obs_dtype == 1)
        entropy_kernel<1><<<grid, kEntropyThreads, 0, stream>>>(observation, rows, (int)S, log_states, o);
    else
        entropy_kernel<2><<<grid, kEntropyThreads, 0, stream>>>(observation, rows, (int)S, log_states, o);
    return cudaGetLastError();
}

}  // namespace torbi_b200
