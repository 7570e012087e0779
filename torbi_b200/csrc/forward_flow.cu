// "Flow" forward kernel: one sequence at a time on all SMs (config 1: 1440 states, batch 1; config 5:
// 4096 states, one sequence of 10^6 frames) without any grid barrier.
//
// forward_wide.cu spends a timestep on: every CTA re-reading the score vector from L2, the scan, and a
// counter barrier (atomicAdd + spin) across ~148 CTAs -- 3.2 us per step at S = 1440 where the math is
// 0.1 us.  Here the timestep is a DATAFLOW exchange, the scheme NCCL's LL protocol uses between GPUs:
//   * the score vector lives in L2 as 8-byte words {fp32 score, 32-bit step tag}.  A producer publishes
//     P_t[j] with ONE 64-bit store; a consumer polls the words it needs until their tag says "step t".
//     An aligned 64-bit access is single-copy atomic, so score and tag always belong together: no fence,
//     no counter, no atomics -- a step costs one L2 store + one L2 load of latency (~2 x 250 cycles)
//     plus the math;
//   * CTA c owns R = ceil(S / #CTAs) destination states.  Thread i owns the SOURCE states 4 (i + 512 k)
//     .. + 3: it polls exactly those scores -- they go straight from L2 into its registers, never through
//     shared memory -- and keeps its columns of the first RR transition rows of the CTA IN REGISTERS
//     for the whole launch (S = 1440: all 10 rows x 4 floats).  Further rows sit in shared memory
//     (S = 4096: 10 rows in registers + 13 in shared memory of the CTA's 28); only the rest streams
//     from L2 each step;
//   * max over the sources: per thread over its columns, per warp with one REDUX on order-preserving
//     integer keys, across the 16 warps through shared memory; the new score is published at once, and
//     the argmax (lowest source index attaining the maximum, viterbi.cpp:94-101) is resolved AFTER the
//     publish, while the scores are on their way to the other SMs: only warps whose partial maximum
//     equals the final one re-evaluate their candidates (the same single fp32 add, so equality is exact).
// Sequences are taken one after the other (longest first); tags keep counting across sequences and two
// buffer sets alternate, so no grid barrier separates them either: every CTA polls the final scores of a
// sequence (and meets at a CTA barrier) before it publishes anything for the next one, which is what
// makes reusing a set two sequences later safe.
// Stores uint16 backpointers like forward_wide (backtrace.cu chases them, blocked for long sequences).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "sync.cuh"

namespace torbi_b200 {

namespace {

constexpr int kFlowThreads = 512;
constexpr int kFlowWarps = kFlowThreads / kWarp;
constexpr int kFlowMaxRows = 64;                    // destination states per CTA the scratch arrays hold

struct FlowParams {
    const void *observation;    // (rows,S), item b at rows item_row[b] ..; 2-byte elements when obs_dtype != 0
    int obs_dtype;
    uint32_t flags;             // conversion flags (runtime: the emission is off the hot loop)
    float *checkpoints;         // (B, ck_rows, S) or null
    int64_t checkpoint_every, ck_rows;
    const float *transition;    // padded copy (SJ, LDP), -inf outside (S,S)
    const float *initial;       // (S,)
    const int32_t *frames;      // (B,)
    const int64_t *item_row;    // (B,) first row of each item in observation
    const int64_t *hist_row;    // (B,) first row of each item in the (compact) backpointers
    const int32_t *order;       // (Bpad,) items longest first
    u64 *ll;                    // [2 sets][2 parities][LDP] tagged scores {fp32 bits, tag}
    uint16_t *backptr;          // (rows,S)
    float *posterior;           // (B,S) or null
    int32_t *final_state;       // (B,)
    int B, S, LDP, NC, R;
    int smem_rows;              // rows kept in shared memory after the RR register rows
};

__device__ __forceinline__ void st_ll(u64 *p, float v, unsigned tag) {
    const u64 w = ((u64)tag << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ld_ll2(const u64 *p, u64 &a, u64 &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ u64 ld_ll(const u64 *p) {
    u64 a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ unsigned tag_of(u64 w) { return (unsigned)(w >> 32); }
__device__ __forceinline__ float value_of(u64 w) { return __uint_as_float((unsigned)w); }

// -DTORBI_FLOW_WATCHDOG (debug builds only): a poll that spins for ~seconds reports what it saw and traps
#ifdef TORBI_FLOW_WATCHDOG
#define FLOW_WATCHDOG_DECL long long spins = 0;
#define FLOW_WATCHDOG(word, q_)                                                                                \
    if (++spins > (1ll << 24)) {                                                                               \
        printf("flow watchdog: cta %d tid %d group %d expects tag %u, sees tag %u value %f\n", blockIdx.x,    \
               threadIdx.x, (q_), tag, tag_of(word), value_of(word));                                          \
        __trap();                                                                                              \
    }
#else
#define FLOW_WATCHDOG_DECL
#define FLOW_WATCHDOG(word, q_)
#endif

// the four scores of source group `q` (sources 4q .. 4q+3) once all carry `tag`; `valid` = how many of
// the four are real states (the others read as 0: their transition entries are -inf)
__device__ __forceinline__ float4 poll_group(const u64 *buf, int q, int valid, unsigned tag) {
    const u64 *p = buf + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid == 4) {
        u64 w0, w1, w2, w3;
        FLOW_WATCHDOG_DECL
        do {
            ld_ll2(p, w0, w1);
            ld_ll2(p + 2, w2, w3);
            FLOW_WATCHDOG(tag_of(w0) != tag ? w0 : tag_of(w1) != tag ? w1 : tag_of(w2) != tag ? w2 : w3, q)
        } while (tag_of(w0) != tag || tag_of(w1) != tag || tag_of(w2) != tag || tag_of(w3) != tag);
        v = make_float4(value_of(w0), value_of(w1), value_of(w2), value_of(w3));
    } else {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        for (int x = 0; x < valid; ++x) {
            u64 w;
            FLOW_WATCHDOG_DECL
            do {
                w = ld_ll(p + x);
                FLOW_WATCHDOG(w, q)
            } while (tag_of(w) != tag);
            e[x] = value_of(w);
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
    }
    return v;
}

// monotonic fp32 -> uint32 map (no NaNs), so a warp maximum is one REDUX
__device__ __forceinline__ unsigned ordered_bits(float v) {
    const unsigned b = __float_as_uint(v);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ float max4(float4 p, float4 a) {
    return fmaxf(fmaxf(p.x + a.x, p.y + a.y), fmaxf(p.z + a.z, p.w + a.w));
}

template <int KQ, int RR>
__global__ void __launch_bounds__(kFlowThreads, 1) forward_flow_kernel(const FlowParams p) {
    extern __shared__ __align__(16) float As[];          // [smem_rows][LDP] transition rows RR .. RR + smem_rows - 1
    __shared__ unsigned part[kFlowMaxRows][kFlowWarps];  // per-warp maxima (ordered bits)
    __shared__ unsigned final_bits[kFlowMaxRows];
    __shared__ int arg[kFlowMaxRows];
    __shared__ float red_val[kFlowWarps];
    __shared__ int red_idx[kFlowWarps];

    const int S = p.S, LDP = p.LDP, R = p.R;
    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int j_begin = blockIdx.x * R;
    const int rows = max(0, min(R, S - j_begin));        // destination states of this CTA

    // the source groups this thread owns, and its columns of the register rows
    int q[KQ], valid[KQ];
    float4 a_reg[RR][KQ];
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
        q[k] = tid + k * kFlowThreads;
        valid[k] = max(0, min(4, S - 4 * q[k]));
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            a_reg[r][k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (r < rows && 4 * q[k] < LDP)
                a_reg[r][k] = __ldg(reinterpret_cast<const float4 *>(p.transition + (int64_t)(j_begin + r) * LDP) + q[k]);
        }
    }
    const int smem_rows = max(0, min(p.smem_rows, rows - RR));
    for (int idx = tid; idx < smem_rows * (LDP / 4); idx += kFlowThreads) {
        const int r = idx / (LDP / 4), c = idx % (LDP / 4);
        reinterpret_cast<float4 *>(As + r * LDP)[c] =
            __ldg(reinterpret_cast<const float4 *>(p.transition + (int64_t)(j_begin + RR + r) * LDP) + c);
    }
    __syncthreads();

    unsigned base = 0;                                   // tag of step 0 of the current sequence
    for (int n = 0; n < p.B; ++n) {
        const int item = p.order[n];
        const int frames = p.frames[item];
        const int64_t row0 = p.item_row[item];
        const int64_t hrow0 = p.hist_row[item];
        u64 *set = p.ll + (size_t)(n & 1) * 2 * LDP;

        // ---- t = 0: P_0[j] = obs[0][j] + initial[j]   (viterbi.cpp:72-76), own destination states ----
        if (tid < rows) {
            const int j = j_begin + tid;
            const float v = convert_observation_rt(load_observation(p.observation, row0 * S + j, p.obs_dtype), p.flags) +
                            p.initial[j];
            st_ll(set + j, v, base);
        }

        // the warp that finalises row r (r % 16 == warp) fetches its emission one step ahead, in lane 0
        constexpr int kRounds = (kFlowMaxRows + kFlowWarps - 1) / kFlowWarps;
        float emit[kRounds];
        auto fetch_emissions = [&](int t) {
#pragma unroll
            for (int i = 0; i < kRounds; ++i) {
                const int r = warp + i * kFlowWarps;
                if (lane == 0 && r < rows && t < frames)
                    emit[i] = convert_observation_rt(
                        load_observation(p.observation, (row0 + t) * S + j_begin + r, p.obs_dtype), p.flags);
            }
        };
        fetch_emissions(1);

        for (int t = 1; t < frames; ++t) {
            const u64 *cur = set + ((t - 1) & 1) * LDP;
            u64 *nxt = set + (t & 1) * LDP;
            const unsigned tag = base + (unsigned)(t - 1);

            // ---- the scores of this thread's sources, straight from L2 into registers ----
            float4 P[KQ];
#pragma unroll
            for (int k = 0; k < KQ; ++k) {
                P[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[k] > 0) P[k] = poll_group(cur, q[k], valid[k], tag);
            }

            // ---- per-warp maxima of every destination row ----
            auto row_max = [&](int r, const float4 *a) {
                float m = -INFINITY;
#pragma unroll
                for (int k = 0; k < KQ; ++k)
                    if (valid[k] > 0) m = fmaxf(m, max4(P[k], a[k]));
                const unsigned w = __reduce_max_sync(kFullMask, ordered_bits(m));
                if (lane == 0) part[r][warp] = w;
            };
#pragma unroll
            for (int r = 0; r < RR; ++r)
                if (r < rows) row_max(r, a_reg[r]);
            for (int r = RR; r < rows; ++r) {
                float4 a[KQ];
                if (r - RR < smem_rows) {
                    const float4 *arow = reinterpret_cast<const float4 *>(As + (r - RR) * LDP);
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        a[k] = valid[k] > 0 ? arow[q[k]] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                } else {
                    const float4 *arow = reinterpret_cast<const float4 *>(p.transition + (int64_t)(j_begin + r) * LDP);
#pragma unroll
                    for (int k = 0; k < KQ; ++k)
                        a[k] = valid[k] > 0 ? arow[q[k]] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                }
                row_max(r, a);
            }
            __syncthreads();

            // ---- across the warps; publish P_t[j] = obs[t][j] + max  (viterbi.cpp:102) ----
#pragma unroll
            for (int i = 0; i < kRounds; ++i) {
                const int r = warp + i * kFlowWarps;
                if (r < rows) {
                    const unsigned w = __reduce_max_sync(kFullMask, lane < kFlowWarps ? part[r][lane] : 0u);
                    if (lane == 0) {
                        const int j = j_begin + r;
                        const float v = emit[i] + from_ordered_bits(w);
                        st_ll(nxt + j, v, base + (unsigned)t);
                        final_bits[r] = w;
                        arg[r] = 0x7fffffff;
                        if (p.checkpoints && t % p.checkpoint_every == 0)
                            p.checkpoints[((int64_t)item * p.ck_rows + (t / p.checkpoint_every - 1)) * S + j] = v;
                    }
                }
            }
            fetch_emissions(t + 1);
            __syncthreads();

            // ---- argmax, off the critical path: lowest source index attaining the maximum ----
            for (int r = 0; r < rows; ++r) {
                const unsigned w = final_bits[r];
                if (part[r][warp] != w) continue;                         // warp-uniform
                const float best = from_ordered_bits(w);
                int first = 0x7fffffff;
#pragma unroll
                for (int k = KQ - 1; k >= 0; --k) {
                    if (valid[k] > 0) {
                        float4 a;
                        if (r < RR) {
                            a = a_reg[0][k];
#pragma unroll
                            for (int x = 1; x < RR; ++x)
                                if (x == r) a = a_reg[x][k];
                        } else if (r - RR < smem_rows) {
                            a = reinterpret_cast<const float4 *>(As + (r - RR) * LDP)[q[k]];
                        } else {
                            a = reinterpret_cast<const float4 *>(p.transition + (int64_t)(j_begin + r) * LDP)[q[k]];
                        }
                        const int s0 = 4 * q[k];
                        if (P[k].w + a.w == best) first = s0 + 3;
                        if (P[k].z + a.z == best) first = s0 + 2;
                        if (P[k].y + a.y == best) first = s0 + 1;
                        if (P[k].x + a.x == best) first = s0;
                    }
                }
                first = (int)__reduce_min_sync(kFullMask, (unsigned)first);
                if (lane == 0) atomicMin(&arg[r], first);
            }
            __syncthreads();
            if (tid < rows)
                p.backptr[(hrow0 + t) * S + j_begin + tid] = (uint16_t)min(arg[tid], S - 1);
        }

        // ---- every CTA waits for the final scores (this is what orders the reuse of the buffer set);
        //      one of them writes them out and takes the first-occurrence argmax (viterbi.cpp:111-113, :218)
        {
            const u64 *fin = set + ((frames - 1) & 1) * LDP;
            const unsigned tag = base + (unsigned)(frames - 1);
            const bool writer = (int)blockIdx.x == n % p.NC;
            float best = -INFINITY;
            int best_i = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < KQ; ++k) {
                if (valid[k] > 0) {
                    const float4 v = poll_group(fin, q[k], valid[k], tag);
                    if (writer) {
                        const float e[4] = {v.x, v.y, v.z, v.w};
                        for (int x = 0; x < valid[k]; ++x) {
                            const int s = 4 * q[k] + x;
                            if (p.posterior) p.posterior[(int64_t)item * S + s] = e[x];
                            if (better(e[x], s, best, best_i)) { best = e[x]; best_i = s; }
                        }
                    }
                }
            }
            // EVERY thread of the CTA has its final scores before any thread publishes for the next sequence:
            // the next sequence's first publish is what lets the other CTAs run ahead and reuse this set
            // (with a one-frame sequence in between there is no other barrier on that path)
            __syncthreads();
            if (writer) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(kFullMask, best, off);
                    const int oi = __shfl_xor_sync(kFullMask, best_i, off);
                    if (better(ov, oi, best, best_i)) { best = ov; best_i = oi; }
                }
                if (lane == 0) { red_val[warp] = best; red_idx[warp] = best_i; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < kFlowWarps; ++w)
                        if (better(red_val[w], red_idx[w], best, best_i)) { best = red_val[w]; best_i = red_idx[w]; }
                    p.final_state[item] = (best_i == 0x7fffffff) ? 0 : best_i;
                }
                __syncthreads();
            }
        }
        base += (unsigned)frames;
    }
}

struct FlowGeometry {
    int NC, R, KQ, RR, smem_rows;
    size_t smem;
    bool ok;
};

// register rows per instantiation: KQ x RR x 4 floats of the transition block stay in registers
constexpr int flow_reg_rows(int kq) { return kq == 1 ? 20 : kq == 2 ? 10 : kq == 4 ? 4 : 1; }

FlowGeometry flow_geometry(int64_t S, int sm_count) {
    FlowGeometry g{};
    const int64_t ldp = padded_sources(S);
    g.R = (int)((S + sm_count - 1) / sm_count);
    g.NC = (int)((S + g.R - 1) / g.R);
    const int groups = (int)(((S + 3) / 4 + kFlowThreads - 1) / kFlowThreads);   // float4 source groups per thread
    g.KQ = groups <= 1 ? 1 : groups <= 2 ? 2 : groups <= 4 ? 4 : 8;
    g.RR = flow_reg_rows(g.KQ);
    if (g.KQ == 1 && g.R <= 10) g.RR = 10;
    if (g.KQ == 1 && g.R <= 4) g.RR = 4;
    const size_t budget = 220 * 1024;                    // dynamic shared memory for transition rows
    const int64_t want = std::max<int64_t>(0, g.R - g.RR);
    g.smem_rows = (int)std::min<int64_t>(want, (int64_t)(budget / (sizeof(float) * ldp)));
    g.smem = sizeof(float) * ldp * g.smem_rows;
    g.ok = groups <= 8 && g.R <= kFlowMaxRows;
    return g;
}

}  // namespace

bool forward_flow_supported(const ForwardArgs &a) {
    return flow_geometry(a.S, a.sm_count).ok && a.ws.transition_pad != nullptr;
}

// Estimated seconds: sequences run one after the other; a step is an L2 round trip (~0.45 us) plus the
// scan of the CTA's rows, with the rows that neither registers nor shared memory hold streamed from L2.
double flow_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count) {
    const FlowGeometry g = flow_geometry(S, sm_count);
    const double streamed = std::max(0, g.R - g.RR - g.smem_rows);
    const double step_us = 0.45 + 0.012 * g.R * g.KQ + 0.12 * streamed * g.KQ;
    double total = 0.0;
    for (int64_t b = 0; b < B; ++b) total += frames_sorted[b] * step_us + 2.0;
    return total * 1e-6;
}

cudaError_t launch_forward_flow(const ForwardArgs &a, cudaStream_t stream) {
    const FlowGeometry g = flow_geometry(a.S, a.sm_count);
    FlowParams p{};
    p.observation = a.observation;
    p.obs_dtype = a.obs_dtype;
    p.flags = a.flags;
    p.checkpoints = a.checkpoints;
    p.checkpoint_every = a.checkpoint_every > 0 ? a.checkpoint_every : 1;
    p.ck_rows = a.checkpoints ? (a.T - 1) / a.checkpoint_every : 0;
    p.transition = a.ws.transition_pad;
    p.initial = a.initial;
    p.frames = a.ws.frames;
    p.item_row = a.ws.item_row;
    p.hist_row = a.ws.hist_row;
    p.order = a.ws.order;
    p.ll = reinterpret_cast<u64 *>(a.ws.score[0]);
    p.backptr = a.ws.backptr;
    p.posterior = a.posterior;
    p.final_state = a.ws.final_state;
    p.B = (int)a.B;
    p.S = (int)a.S;
    p.LDP = (int)padded_sources(a.S);
    p.NC = g.NC;
    p.R = g.R;
    p.smem_rows = g.smem_rows;
    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    // tags start out invalid (0xffffffff is never a step of this call): 2 sets x 2 parities x LDP words,
    // at the head of the forward-score region (128 x LDP floats per buffer: more than enough)
    e = cudaMemsetAsync(p.ll, 0xff, sizeof(u64) * 4 * p.LDP, stream);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&p};
    auto launch = [&](auto kernel) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (err != cudaSuccess) return err;
        // cooperative: every CTA polls every other CTA's stores, so all must be co-resident
        return cudaLaunchCooperativeKernel((void *)kernel, dim3(g.NC), dim3(kFlowThreads), args, g.smem, stream);
    };
    if (g.KQ == 1 && g.RR == 4) return launch(forward_flow_kernel<1, 4>);
    if (g.KQ == 1 && g.RR == 10) return launch(forward_flow_kernel<1, 10>);
    if (g.KQ == 1) return launch(forward_flow_kernel<1, 20>);
    if (g.KQ == 2) return launch(forward_flow_kernel<2, 10>);
    if (g.KQ == 4) return launch(forward_flow_kernel<4, 4>);
    return launch(forward_flow_kernel<8, 1>);
}

}  // namespace torbi_b200
