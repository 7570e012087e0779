// Ragged backtrace: restates viterbi_backtrace_trellis_cpu
// (/root/reference/torbi/csrc/viterbi.cpp:140-160) together with the
// `argmax(1).unsqueeze(1).repeat(1, T)` pre-fill of viterbi.cpp:218-221:
//   indices[b, t] = final_state[b]                      for t >= frames_b - 1
//   idx = final_state[b]; for t = frames_b-1 .. 1: idx = bp[b,t,idx]; indices[b,t-1] = idx
// One CTA per sequence: thread 0 chases the dependent uint16 loads while the
// other threads write the tail.  Per-item batch_frames, 64-bit offsets.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

constexpr int kBacktraceThreads = 128;

__global__ void __launch_bounds__(kBacktraceThreads)
backtrace_kernel(const uint16_t *__restrict__ backptr, const int32_t *__restrict__ frames_d,
                 const int64_t *__restrict__ item_row, const int64_t *__restrict__ hist_row,
                 const int32_t *__restrict__ final_state, int32_t *__restrict__ indices, int64_t T, int64_t S) {
    // T > 0: padded layout, indices (B,T), tail t in [frames_b - 1, T) filled; T == 0: packed layout,
    // indices (rows,), item b at item_row[b] .. + frames_b (the tail is the last frame alone)
    const int64_t b = blockIdx.x;
    const int frames = frames_d[b];
    const int last = final_state[b];
    const int64_t row0 = item_row[b];
    int32_t *out = indices + row0;
    const int64_t out_len = T > 0 ? T : frames;
    if (threadIdx.x == 0) {
        const uint16_t *bp = backptr + hist_row[b] * S;
        int idx = last;
        for (int64_t t = frames - 1; t >= 1; --t) {
            idx = bp[t * S + idx];
            out[t - 1] = idx;
        }
    } else {
        for (int64_t t = frames - 1 + (threadIdx.x - 1); t < out_len; t += kBacktraceThreads - 1)
            out[t] = last;
    }
}

// ---------------------------------------------------------------------------
// Blocked backtrace for long sequences with stored backpointers (config 5: 4096 states x 10^6
// frames).  backtrace_kernel above is one thread chasing frames_b dependent 2-byte loads, each an
// L2 / HBM miss (~0.7 us): 0.7 s for 10^6 frames.  The chain idx_{t-1} = bp[t][idx_t] is a
// composition of maps on {0..S-1}, so it is cut into blocks of K frames:
//   1. compose (grid: blocks x sequences, all SMs): f_m(s) = state at time lo_m - 1 when the path is
//      in state s at time hi_m, for ALL s at once -- every thread carries a few states through the
//      block's K rows (the rows are prefetched into L2 ahead; gathers within a row hit L1 / L2);
//   2. link (one thread per sequence): the state at every block boundary, entry_{m+1} = f_m(entry_m),
//      a chain of only frames / K lookups;
//   3. chase (grid: blocks x sequences): every block follows its own K backpointers from its known
//      entry state -- the same dependent loads as before, but all blocks at once.
// Same loads, same table: the indices are those of the plain chase, bit for bit.  The maps live in
// the forward-score buffers of the workspace, which are dead once the forward kernel has finished.
// ---------------------------------------------------------------------------
namespace {

constexpr int kComposeThreads = 1024;
constexpr int kComposeStatesPerThread = 64;         // S <= 65535
constexpr int kPrefetchRows = 12;

__device__ __forceinline__ void prefetch_l2_line(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

struct BlockedGeometry {
    int64_t K;            // frames per block
    int64_t max_blocks;   // blocks of the longest sequence
};

// block m of a sequence with `frames` frames covers t in [lo, hi], hi = frames - 1 - m K, lo = max(1, hi - K + 1)
template <int kPerThread>       // states carried per thread: S <= kPerThread * kComposeThreads
__global__ void __launch_bounds__(kComposeThreads)
backtrace_compose_kernel(const uint16_t *__restrict__ backptr, const int32_t *__restrict__ frames_d,
                         const int64_t *__restrict__ hist_row, uint16_t *__restrict__ maps, int S, int64_t K,
                         int64_t max_blocks) {
    const int64_t b = blockIdx.y, m = blockIdx.x;
    const int64_t hi = (int64_t)frames_d[b] - 1 - m * K;
    if (hi < 1) return;
    const int64_t lo = max((int64_t)1, hi - K + 1);
    const uint16_t *bp = backptr + hist_row[b] * S;
    int cur[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) cur[i] = min(threadIdx.x + i * kComposeThreads, S - 1);
    const int64_t row_bytes = (int64_t)S * 2;
    for (int64_t t = hi; t >= lo; --t) {
        // rows are contiguous: pull the one kPrefetchRows ahead into L2 (one 128-byte line per thread)
        const int64_t tp = t - kPrefetchRows;
        if (tp >= lo)
            for (int64_t off = (int64_t)threadIdx.x * 128; off < row_bytes; off += (int64_t)kComposeThreads * 128)
                prefetch_l2_line(reinterpret_cast<const char *>(bp + tp * S) + off);
        const uint16_t *row = bp + t * S;
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) cur[i] = row[cur[i]];        // (slots past S - 1 repeat state S - 1)
    }
    uint16_t *out = maps + (b * max_blocks + m) * S;
#pragma unroll
    for (int i = 0; i < kPerThread; ++i)
        if (threadIdx.x + i * kComposeThreads < S) out[threadIdx.x + i * kComposeThreads] = (uint16_t)cur[i];
}

__global__ void backtrace_link_kernel(const uint16_t *__restrict__ maps, const int32_t *__restrict__ frames_d,
                                      const int32_t *__restrict__ final_state, int32_t *__restrict__ entry,
                                      int64_t B, int S, int64_t K, int64_t max_blocks) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    int idx = final_state[b];
    const int64_t frames = frames_d[b];
    for (int64_t m = 0; m < max_blocks; ++m) {
        entry[b * max_blocks + m] = idx;                      // state at time hi_m
        if (frames - 1 - m * K < 1) break;
        idx = maps[(b * max_blocks + m) * S + idx];
    }
}

__global__ void __launch_bounds__(kBacktraceThreads)
backtrace_chase_kernel(const uint16_t *__restrict__ backptr, const int32_t *__restrict__ frames_d,
                       const int64_t *__restrict__ item_row, const int64_t *__restrict__ hist_row,
                       const int32_t *__restrict__ final_state, const int32_t *__restrict__ entry, int32_t *__restrict__ indices, int64_t T, int64_t S,
                       int64_t K, int64_t max_blocks) {
    const int64_t b = blockIdx.y, m = blockIdx.x;
    const int64_t frames = frames_d[b];
    const int64_t row0 = item_row[b];
    int32_t *out = indices + row0;
    if (m == 0 && threadIdx.x > 0) {
        const int last = final_state[b];                      // viterbi.cpp:219-221
        const int64_t out_len = T > 0 ? T : frames;           // (packed layout: the last frame alone)
        for (int64_t t = frames - 1 + (threadIdx.x - 1); t < out_len; t += kBacktraceThreads - 1) out[t] = last;
        return;
    }
    const int64_t hi = frames - 1 - m * K;
    if (hi < 1 || threadIdx.x != 0) return;
    const int64_t lo = max((int64_t)1, hi - K + 1);
    const uint16_t *bp = backptr + hist_row[b] * S;
    int idx = entry[b * max_blocks + m];
    for (int64_t t = hi; t >= lo; --t) {
        idx = bp[t * S + idx];
        out[t - 1] = idx;
    }
}

}  // namespace

cudaError_t launch_backtrace_blocked(const Workspace &ws, int32_t *indices, int64_t B, int64_t T, int64_t S,
                                     int64_t max_frames, int *launches, cudaStream_t stream) {
    // maps + entry states go where the rotating forward scores were (score[0] .. transition_pad)
    char *region = reinterpret_cast<char *>(ws.score[0]);
    const size_t capacity = reinterpret_cast<char *>(ws.transition_pad) - region - 256;
    const size_t per_block = (size_t)B * ((size_t)S * sizeof(uint16_t) + sizeof(int32_t));   // one map + one entry per sequence
    int64_t max_blocks = (int64_t)std::min<size_t>(capacity / per_block, 65535);
    max_blocks = std::min<int64_t>(max_blocks, (max_frames - 1 + 255) / 256);               // blocks of >= 256 frames
    if (max_blocks < 2 || S > (int64_t)kComposeThreads * kComposeStatesPerThread) {
        if (launches) *launches = 1;
        return launch_backtrace(ws, indices, B, T, S, stream);
    }
    const int64_t K = (max_frames - 1 + max_blocks - 1) / max_blocks;
    max_blocks = (max_frames - 1 + K - 1) / K;
    uint16_t *maps = reinterpret_cast<uint16_t *>(region);
    int32_t *entry = reinterpret_cast<int32_t *>(region + align_up((size_t)B * max_blocks * S * sizeof(uint16_t), 256));
    const dim3 grid((unsigned)max_blocks, (unsigned)B);
    if (S <= 2 * kComposeThreads)
        backtrace_compose_kernel<2><<<grid, kComposeThreads, 0, stream>>>(ws.backptr, ws.frames, ws.hist_row, maps, (int)S, K, max_blocks);
    else if (S <= 4 * kComposeThreads)
        backtrace_compose_kernel<4><<<grid, kComposeThreads, 0, stream>>>(ws.backptr, ws.frames, ws.hist_row, maps, (int)S, K, max_blocks);
    else if (S <= 16 * kComposeThreads)
        backtrace_compose_kernel<16><<<grid, kComposeThreads, 0, stream>>>(ws.backptr, ws.frames, ws.hist_row, maps, (int)S, K, max_blocks);
    else
        backtrace_compose_kernel<kComposeStatesPerThread><<<grid, kComposeThreads, 0, stream>>>(ws.backptr, ws.frames, ws.hist_row, maps, (int)S, K, max_blocks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    backtrace_link_kernel<<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(maps, ws.frames, ws.final_state, entry, B,
                                                                          (int)S, K, max_blocks);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    backtrace_chase_kernel<<<dim3((unsigned)max_blocks, (unsigned)B), kBacktraceThreads, 0, stream>>>(
        ws.backptr, ws.frames, ws.item_row, ws.hist_row, ws.final_state, entry, indices, T, S, K, max_blocks);
    if (launches) *launches = 3;
    return cudaGetLastError();
}

// torbi_b200_decode_checkpointed after a score-history kernel: the history already holds every P_t
namespace {
__global__ void gather_checkpoints_kernel(const float *__restrict__ history, const int32_t *__restrict__ frames_d,
                                          const int64_t *__restrict__ hist_row, float *__restrict__ checkpoints,
                                          int64_t B, int64_t S, int64_t every, int64_t rows) {
    const int64_t total = B * rows * S;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i % S, k = (i / S) % rows, b = i / (S * rows);
        const int64_t t = (k + 1) * every;
        if (t < frames_d[b]) checkpoints[i] = history[(hist_row[b] + t) * S + s];
    }
}
}  // namespace

cudaError_t launch_gather_checkpoints(const ForwardArgs &a, cudaStream_t stream) {
    const int64_t rows = (a.T - 1) / a.checkpoint_every;
    if (rows < 1) return cudaSuccess;
    gather_checkpoints_kernel<<<a.sm_count * 4, 256, 0, stream>>>(
        reinterpret_cast<const float *>(a.ws.backptr), a.ws.frames, a.ws.hist_row, a.checkpoints, a.B, a.S, a.checkpoint_every, rows);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Backtrace by recomputation (score-history mode of the tiled forward kernel).
// The forward pass kept every P_t (fp32) instead of S backpointers per step; only the
// one backpointer per step that lies on the decoded path is ever needed:
//   idx_{t-1} = lowest s maximising P_{t-1}[s] + A[idx_t][s]
// -- the same single fp32 add and the same first-occurrence rule as the forward scan
// (viterbi.cpp:94-101), so the result is bit-identical to chasing stored backpointers.
// One warp per sequence; per step one row of A (L2-resident) and one row of the history
// (prefetched into L2 kAhead steps early, then into registers one step early).
// ---------------------------------------------------------------------------
namespace {

constexpr int kRecomputeThreads = 128;
constexpr int kAhead = 8;

// monotonic fp32 -> uint32 map (no NaNs), so the warp max is one REDUX instruction
__device__ __forceinline__ unsigned ordered_bits(float v) {
    const unsigned b = __float_as_uint(v);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ float warp_max(float m) {
    return from_ordered_bits(__reduce_max_sync(kFullMask, ordered_bits(m)));
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_row(const float *row, int S, int lane) {
    const char *base = reinterpret_cast<const char *>(row);
    for (int off = lane * 128; off < S * 4; off += kWarp * 128) prefetch_l2(base + off);
}

// NK > 0: S % 4 == 0 and S <= 128 * NK -- a lane keeps its NK float4 groups of the row in registers.
// NK == 0: any S, two passes over the row (max, then first index attaining it).
template <int NK>
__global__ void __launch_bounds__(kRecomputeThreads)
backtrace_recompute_kernel(const float *__restrict__ history, const float *__restrict__ transition,
                           int64_t lda, const int32_t *__restrict__ frames_d,
                           const int64_t *__restrict__ item_row, const int64_t *__restrict__ hist_row,
                           const int32_t *__restrict__ final_state, int32_t *__restrict__ indices, int64_t B,
                           int64_t T, int S) {
    const int lane = threadIdx.x % kWarp;
    const int64_t b = (int64_t)blockIdx.x * (kRecomputeThreads / kWarp) + threadIdx.x / kWarp;
    if (b >= B) return;
    const int frames = frames_d[b];
    const int last = final_state[b];
    const int64_t row0 = item_row[b];
    int32_t *out = indices + row0;
    const int64_t out_len = T > 0 ? T : frames;       // T == 0: packed layout, the tail is the last frame alone
    for (int64_t t = frames - 1 + lane; t < out_len; t += kWarp) out[t] = last;      // viterbi.cpp:219-221
    const float *hist = history + hist_row[b] * S;
    int idx = last;

    if constexpr (NK > 0) {
        bool valid[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) valid[k] = 4 * (lane + kWarp * k) < S;
        float4 next[NK];                                  // history row for the coming step
        auto load_row = [&](int64_t t) {
            const float4 *row = reinterpret_cast<const float4 *>(hist + t * S) + lane;
#pragma unroll
            for (int k = 0; k < NK; ++k)
                if (valid[k]) next[k] = __ldcs(row + kWarp * k);
        };
        for (int t = frames - 2; t >= 0 && t > frames - 2 - kAhead; --t) prefetch_row(hist + (int64_t)t * S, S, lane);
        if (frames >= 2) load_row(frames - 2);
        for (int t = frames - 1; t >= 1; --t) {
            const float4 *arow = reinterpret_cast<const float4 *>(transition + (int64_t)idx * lda) + lane;
            float4 v[NK];
#pragma unroll
            for (int k = 0; k < NK; ++k)
                if (valid[k]) v[k] = __ldg(arow + kWarp * k);         // the dependent load of the step
            float4 pc[NK];
#pragma unroll
            for (int k = 0; k < NK; ++k) pc[k] = next[k];
            if (t >= 2) load_row(t - 2);                              // independent: overlaps
            if (t - 2 - kAhead >= 0) prefetch_row(hist + (int64_t)(t - 2 - kAhead) * S, S, lane);
            float m = -INFINITY;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                if (valid[k]) {
                    v[k].x += pc[k].x; v[k].y += pc[k].y; v[k].z += pc[k].z; v[k].w += pc[k].w;
                    m = fmaxf(fmaxf(fmaxf(m, v[k].x), fmaxf(v[k].y, v[k].z)), v[k].w);
                }
            }
            m = warp_max(m);
            unsigned first = 0x7fffffffu;
#pragma unroll
            for (int k = NK - 1; k >= 0; --k) {
                if (valid[k]) {
                    const unsigned s0 = 4 * (lane + kWarp * k);
                    if (v[k].w == m) first = s0 + 3;
                    if (v[k].z == m) first = s0 + 2;
                    if (v[k].y == m) first = s0 + 1;
                    if (v[k].x == m) first = s0;
                }
            }
            first = __reduce_min_sync(kFullMask, first);
            idx = first < (unsigned)S ? (int)first : 0;               // (NaN input only: outside the contract)
            if (lane == 0) out[t - 1] = idx;
        }
    } else {
        for (int t = frames - 1; t >= 1; --t) {
            const float *prow = hist + (int64_t)(t - 1) * S;
            const float *arow = transition + (int64_t)idx * lda;
            if (t - 1 - kAhead >= 0) prefetch_row(hist + (int64_t)(t - 1 - kAhead) * S, S, lane);
            float m = -INFINITY;
            for (int s = lane; s < S; s += kWarp) m = fmaxf(m, __ldg(prow + s) + __ldg(arow + s));
            m = warp_max(m);
            unsigned first = 0x7fffffffu;
            for (int s = lane; s < S; s += kWarp)
                if (first == 0x7fffffffu && __ldg(prow + s) + __ldg(arow + s) == m) first = s;
            first = __reduce_min_sync(kFullMask, first);
            idx = first < (unsigned)S ? (int)first : 0;
            if (lane == 0) out[t - 1] = idx;
        }
    }
}

}  // namespace

// `transition` is the padded copy the tiled kernel made: (padded_dests, lda) with 16-byte aligned rows.
cudaError_t launch_backtrace_recompute(const Workspace &ws, int32_t *indices, int64_t B, int64_t T,
                                       int64_t S, cudaStream_t stream) {
    const float *history = reinterpret_cast<const float *>(ws.backptr);
    const int64_t lda = padded_sources(S);
    const unsigned grid = (unsigned)((B + kRecomputeThreads / kWarp - 1) / (kRecomputeThreads / kWarp));
    if (S % 4 == 0 && S <= 128 * 4)
        backtrace_recompute_kernel<4><<<grid, kRecomputeThreads, 0, stream>>>(
            history, ws.transition_pad, lda, ws.frames, ws.item_row, ws.hist_row, ws.final_state, indices, B, T, (int)S);
    else if (S % 4 == 0 && S <= 128 * 12)
        backtrace_recompute_kernel<12><<<grid, kRecomputeThreads, 0, stream>>>(
            history, ws.transition_pad, lda, ws.frames, ws.item_row, ws.hist_row, ws.final_state, indices, B, T, (int)S);
    else
        backtrace_recompute_kernel<0><<<grid, kRecomputeThreads, 0, stream>>>(
            history, ws.transition_pad, lda, ws.frames, ws.item_row, ws.hist_row, ws.final_state, indices, B, T, (int)S);
    return cudaGetLastError();
}

cudaError_t launch_backtrace(const Workspace &ws, int32_t *indices, int64_t B, int64_t T,
                             int64_t S, cudaStream_t stream) {
    backtrace_kernel<<<(unsigned)B, kBacktraceThreads, 0, stream>>>(
        ws.backptr, ws.frames, ws.item_row, ws.hist_row, ws.final_state, indices, T, S);
    return cudaGetLastError();
}

}  // namespace torbi_b200
