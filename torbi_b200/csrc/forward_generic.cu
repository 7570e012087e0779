// Generic forward kernel: one CTA per sequence, any number of states (1..65535),
// any batch.  It is the shape-agnostic member of the family (odd state counts,
// tiny batches); the batch-tiled kernel in forward_tiled.cu is the fast path.
//
// Recursion restated from /root/reference/torbi/csrc/viterbi.cpp:72-113:
//   P_0[s]   = obs[0,s] + initial[s]
//   P_t[j]   = obs[t,j] + max_s ( P_{t-1}[s] + transition[j][s] )
//   bp[t,j]  = lowest s attaining that max          (strict `>` scan, :94-101)
// Each warp owns destination rows j = warp, warp+nwarps, ...; lanes stride the
// source states in ascending order (so a lane's own strict `>` keeps its lowest
// index) and the warp reduces on (value desc, index asc) -- unlike the
// reference CUDA kernel (viterbi.cu:98-112), whose shuffle reduce breaks ties
// by lane and disagrees with its own CPU path.
#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

constexpr int kGenericThreads = 512;

template <uint32_t kFlags>
__global__ void __launch_bounds__(kGenericThreads)
forward_generic_kernel(const float *__restrict__ observation, int obs_dtype,
                       float *__restrict__ checkpoints, int64_t checkpoint_every,
                       const int32_t *__restrict__ frames_d, const int64_t *__restrict__ item_row,
                       const int64_t *__restrict__ hist_row,
                       const float *__restrict__ transition,
                       const float *__restrict__ initial, float *score0, float *score1,
                       uint16_t *__restrict__ backptr, float *__restrict__ posterior,
                       int32_t *__restrict__ final_state, int64_t ck_rows, int S,
                       int64_t score_stride, int scores_in_smem) {
    extern __shared__ float smem_scores[];
    __shared__ float red_val[kGenericThreads / kWarp];
    __shared__ int red_idx[kGenericThreads / kWarp];

    const int64_t b = blockIdx.x;
    const int frames = frames_d[b];
    const int64_t obs0 = item_row[b] * S;               // element offset of this sequence
    uint16_t *bp = backptr + hist_row[b] * S;
    float *cur = scores_in_smem ? smem_scores : score0 + b * score_stride;
    float *nxt = scores_in_smem ? smem_scores + S : score1 + b * score_stride;

    const int lane = threadIdx.x % kWarp;
    const int warp = threadIdx.x / kWarp;
    const int nwarps = blockDim.x / kWarp;

    for (int i = threadIdx.x; i < S; i += blockDim.x)
        cur[i] = convert_observation<kFlags>(load_observation(observation, obs0 + i, obs_dtype)) + initial[i];
    __syncthreads();

    const bool vec4 = (S % 4) == 0;
    for (int t = 1; t < frames; ++t) {
        const int64_t obs_t = obs0 + (int64_t)t * S;
        float *ck = (checkpoints && t % checkpoint_every == 0)
            ? checkpoints + (b * ck_rows + (t / checkpoint_every - 1)) * S : nullptr;
        uint16_t *bp_t = bp + (int64_t)t * S;
        for (int j = warp; j < S; j += nwarps) {
            const float *row = transition + (int64_t)j * S;
            float emit = 0.f;
            // (raw: converted where it is added, so the warp does not wait for the load before its scan)
            if (lane == 0) emit = load_observation(observation, obs_t + j, obs_dtype);
            float best = -INFINITY;
            int arg = 0x7fffffff;
            bool any = false;
            if (vec4) {
                const float4 *row4 = reinterpret_cast<const float4 *>(row);
                const float4 *cur4 = reinterpret_cast<const float4 *>(cur);
                for (int q = lane; q < S / 4; q += kWarp) {
                    const float4 a = __ldg(row4 + q);
                    const float4 p = cur4[q];
                    const float v0 = p.x + a.x, v1 = p.y + a.y, v2 = p.z + a.z, v3 = p.w + a.w;
                    if (!any) { best = v0; arg = 4 * q; any = true; }
                    else if (v0 > best) { best = v0; arg = 4 * q; }
                    if (v1 > best) { best = v1; arg = 4 * q + 1; }
                    if (v2 > best) { best = v2; arg = 4 * q + 2; }
                    if (v3 > best) { best = v3; arg = 4 * q + 3; }
                }
            } else {
                for (int s = lane; s < S; s += kWarp) {
                    const float v = cur[s] + __ldg(row + s);
                    if (!any) { best = v; arg = s; any = true; }
                    else if (v > best) { best = v; arg = s; }
                }
            }
            // lanes with no element carry (-inf, INT_MAX) and never win a tie
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(kFullMask, best, off);
                const int oi = __shfl_xor_sync(kFullMask, arg, off);
                if (better(ov, oi, best, arg)) { best = ov; arg = oi; }
            }
            if (lane == 0) {
                const float score = convert_observation<kFlags>(emit) + best;
                bp_t[j] = static_cast<uint16_t>(arg);
                nxt[j] = score;
                if (ck) ck[j] = score;
            }
        }
        float *tmp = cur; cur = nxt; nxt = tmp;
        __syncthreads();
    }

    // final scores out (viterbi.cpp:111-113) + first-occurrence argmax (:218)
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const float v = cur[i];
        if (posterior) posterior[b * S + i] = v;
        if (better(v, i, best, arg)) { best = v; arg = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(kFullMask, best, off);
        const int oi = __shfl_xor_sync(kFullMask, arg, off);
        if (better(ov, oi, best, arg)) { best = ov; arg = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = arg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarps; ++w)
            if (better(red_val[w], red_idx[w], best, arg)) { best = red_val[w]; arg = red_idx[w]; }
        // all scores -inf or NaN: the reference's argmax returns 0 for an all -inf row
        final_state[b] = (arg == 0x7fffffff) ? 0 : arg;
    }
}

template <uint32_t kFlags>
static cudaError_t launch_generic_t(const ForwardArgs &a, cudaStream_t stream) {
    const size_t smem_needed = sizeof(float) * 2 * a.S;
    const int in_smem = smem_needed <= 200 * 1024;
    const size_t smem = in_smem ? smem_needed : 0;
    auto kernel = forward_generic_kernel<kFlags>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<(unsigned)a.B, kGenericThreads, smem, stream>>>(
        a.observation, a.obs_dtype, a.checkpoints, a.checkpoint_every, a.ws.frames, a.ws.item_row, a.ws.hist_row, a.transition, a.initial, a.ws.score[0], a.ws.score[1],
        a.ws.backptr, a.posterior, a.ws.final_state, a.checkpoints ? (a.T - 1) / a.checkpoint_every : 0, (int)a.S, padded_sources(a.S), in_smem);
    return cudaGetLastError();
}

cudaError_t launch_forward_generic(const ForwardArgs &a, cudaStream_t stream) {
    switch (a.flags & 3u) {
        case 0: return launch_generic_t<0>(a, stream);
        case 1: return launch_generic_t<1>(a, stream);
        case 2: return launch_generic_t<2>(a, stream);
        default: return launch_generic_t<3>(a, stream);
    }
}

}  // namespace torbi_b200
