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
//   * max over the sources: every thread scans its columns in ascending order (strict >, so the lowest
//     source wins inside a thread, viterbi.cpp:94-101); a warp needs two REDUX instructions per row -- the
//     maximum on order-preserving integer keys, then the minimum source index among the lanes that attain
//     it; the 16 warps meet through (double-buffered) shared memory and the ONE CTA barrier of the step,
//     after which the warp that owns the row repeats the two REDUX over the per-warp results, adds the
//     emission, publishes the new score and stores the backpointer;
//   * the exchange buffer can be replicated (TORBI_B200_FLOW_REPLICAS = K: a producer stores each new score K
//     times, CTA c polls copy c % K) to spread the polling over more L2 slices.  Measured on the B200: the
//     poll costs ~700 cycles per step with K = 1 and nothing is gained up to K = 64
//     (profiles/r02_flow_kernel.txt), so one copy is the default;
// Sequences are taken one after the other (longest first); tags keep counting across sequences and two
// buffer sets alternate, so no grid barrier separates them either: every CTA polls the final scores of a
// sequence (and meets at a CTA barrier) before it publishes anything for the next one, which is what
// makes reusing a set two sequences later safe.
// Stores uint16 backpointers like forward_wide (backtrace.cu chases them, blocked for long sequences).
#include <algorithm>
#include <cstdlib>

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
    u64 *ll;                    // [replicas][2 sets][2 parities][LDP] tagged scores {fp32 bits, tag}
    int replicas;               // copies of the exchange buffer: CTA c polls copy c % replicas (see below)
    uint16_t *backptr;          // (rows,S)
    float *posterior;           // (B,S) or null
    int32_t *final_state;       // (B,)
    int B, S, LDP, NC, R;
    int smem_rows;              // rows kept in shared memory after the RR register rows
    int gate;                   // 1: a few threads poll one word per producer CTA before everyone loads its scores
    int direct;                 // 1: per-THREAD candidates go to shared memory and the row's warp reduces all of them
                                //    (no in-warp butterfly per row; needs 2 x rows x 512 x 8 bytes of shared memory)
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

// per-phase cycle counters (tools/profile_phases.sh builds with TORBI_TILED_PROFILE; never in the shipped library)
#ifdef TORBI_TILED_PROFILE
#define FPROF_DECL long long fprof_t0 = clock64(), fprof_acc[6] = {0, 0, 0, 0, 0, 0};
#define FPROF_MARK(i) { const long long now_ = clock64(); fprof_acc[i] += now_ - fprof_t0; fprof_t0 = now_; }
#define FPROF_PRINT(steps)                                                                                     \
    if (tid == 0 && blockIdx.x % 37 == 0)                                                                      \
        printf("cta %d steps %d cycles/step: gate %lld, score loads %lld, row maxima + argmax %lld, CTA barrier + final "     \
               "reduce + publish %lld, streamed-row requests %lld\n", blockIdx.x, (steps), fprof_acc[5] / (steps),            \
               fprof_acc[0] / (steps), fprof_acc[1] / (steps), fprof_acc[2] / (steps), fprof_acc[4] / (steps));
#else
#define FPROF_DECL
#define FPROF_MARK(i)
#define FPROF_PRINT(steps)
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
__device__ __forceinline__ u64 add2(u64 a, u64 b) {     // two fp32 adds (FADD2), IEEE rn
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

// Warp-wide (value desc, source index asc) reduction of NR rows at once without REDUX (measured on the
// B200: ~37 cycles of an SM-wide serial resource per REDUX, which made 2 REDUX per row per warp the whole
// step).  Butterfly with row halving: at distance 16 a lane hands the partner the half of the rows the
// partner keeps and reduces the half it keeps, then likewise at 8, 4, ... -- NR + log2(32 / NR) shuffle
// pairs in all instead of 5 NR.  On return v[0] / ix[0] of lane l hold row l / (32 / NR).
constexpr int kChunk = 8;
constexpr int kEmissionAhead = 8;
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void take_better(float &v, int &ix, float ov, int oi) {
    if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
}
template <int N, int OFF>
__device__ __forceinline__ void reduce_rows(float (&v)[kChunk], int (&ix)[kChunk], int lane) {
    if constexpr (OFF >= 1) {
        if constexpr (N > 1) {
            constexpr int H = N / 2;
            const bool upper = lane & OFF;
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const float send_v = upper ? v[i] : v[i + H];
                const int send_i = upper ? ix[i] : ix[i + H];
                float keep_v = upper ? v[i + H] : v[i];
                int keep_i = upper ? ix[i + H] : ix[i];
                take_better(keep_v, keep_i, __shfl_xor_sync(kFullMask, send_v, OFF), __shfl_xor_sync(kFullMask, send_i, OFF));
                v[i] = keep_v;
                ix[i] = keep_i;
            }
            reduce_rows<H, OFF / 2>(v, ix, lane);
        } else {
            take_better(v[0], ix[0], __shfl_xor_sync(kFullMask, v[0], OFF), __shfl_xor_sync(kFullMask, ix[0], OFF));
            reduce_rows<1, OFF / 2>(v, ix, lane);
        }
    }
}

// kAllReg: every row of the CTA's transition slice sits in registers (R <= RR): the shared-memory and streamed
// row paths -- and the per-thread arrays they index at run time, which ptxas keeps in local memory -- are
// compiled out
template <int KQ, int RR, bool kAllReg>
__global__ void __launch_bounds__(kFlowThreads, 1) forward_flow_kernel(const FlowParams p) {
    // dynamic shared memory: per-warp partial maxima of every row, double buffered (ordered score bits and
    // lowest source index), then the shared-memory rows of the transition slice
    extern __shared__ __align__(16) unsigned flow_smem[];
    const int S = p.S, LDP = p.LDP, R = p.R;
    const int RP = (R + kChunk - 1) / kChunk * kChunk;             // rows padded to whole chunks
    float *pval = reinterpret_cast<float *>(flow_smem);            // [2][RP][kFlowWarps]
    int *pidx = reinterpret_cast<int *>(pval + 2 * RP * kFlowWarps);    // [2][RP][kFlowWarps]
    float *As = reinterpret_cast<float *>(pidx + 2 * RP * kFlowWarps);  // [smem_rows][LDP] rows RR .. RR + smem_rows - 1
    // direct mode (no shared-memory transition rows): [2][R][kFlowThreads] candidate keys, one per thread and row:
    // (ordered score bits << 32) | ~source -- a 64-bit integer maximum is (score desc, source asc)
    u64 *cand = reinterpret_cast<u64 *>(flow_smem);
    __shared__ float red_val[kFlowWarps];
    __shared__ int red_idx[kFlowWarps];

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int j_begin = blockIdx.x * R;
    const int rows = max(0, min(R, S - j_begin));        // destination states of this CTA
    // warps whose lanes own no source state at all (S = 1440: warps 12 .. 15) skip the scan
    const int active_warps = min(kFlowWarps, ((S + 3) / 4 + kWarp - 1) / kWarp);
    const bool warp_active = warp < active_warps;

    // the source groups this thread owns, and its columns of the register rows
    int q[KQ], valid[KQ];
    ulonglong2 a_reg[RR][KQ];
    const ulonglong2 minus_inf = make_ulonglong2(0xff800000ff800000ull, 0xff800000ff800000ull);
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
        q[k] = tid + k * kFlowThreads;
        valid[k] = max(0, min(4, S - 4 * q[k]));
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            a_reg[r][k] = minus_inf;
            if (r < rows && 4 * q[k] < LDP)
                a_reg[r][k] = __ldg(reinterpret_cast<const ulonglong2 *>(p.transition + (int64_t)(j_begin + r) * LDP) + q[k]);
        }
    }
    const int smem_rows = kAllReg ? 0 : max(0, min(p.smem_rows, rows - RR));
    for (int idx = tid; idx < smem_rows * (LDP / 4); idx += kFlowThreads) {
        const int r = idx / (LDP / 4), c = idx % (LDP / 4);
        reinterpret_cast<float4 *>(As + r * LDP)[c] =
            __ldg(reinterpret_cast<const float4 *>(p.transition + (int64_t)(j_begin + RR + r) * LDP) + c);
    }
    const int first_streamed = RR + smem_rows;
    const int ns = kAllReg ? 0 : max(0, rows - first_streamed);        // rows streamed from L2 every step
    const int stream_stride = ns > 1 ? max(1, smem_rows / (ns - 1)) : smem_rows + 1;
    ulonglong2 sbuf0[KQ], sbuf1[KQ];
    auto load_streamed = [&](int i, ulonglong2 *dst) {
        const ulonglong2 *arow = reinterpret_cast<const ulonglong2 *>(p.transition + (int64_t)(j_begin + first_streamed + i) * LDP);
#pragma unroll
        for (int k = 0; k < KQ; ++k) dst[k] = valid[k] > 0 ? __ldg(arow + q[k]) : minus_inf;
    };
    FPROF_DECL
    __syncthreads();

    unsigned base = 0;                                   // tag of step 0 of the current sequence
    unsigned steps_done = 0;                             // parity of the partial-maxima buffers
    for (int n = 0; n < p.B; ++n) {
        const int item = p.order[n];
        const int frames = p.frames[item];
        const int64_t row0 = p.item_row[item];
        const int64_t hrow0 = p.hist_row[item];
        // buffer set of this sequence in replica 0 / in the replica this CTA reads
        const size_t replica_words = (size_t)4 * LDP;
        u64 *set0 = p.ll + (size_t)(n & 1) * 2 * LDP;
        const u64 *set = set0 + (size_t)(blockIdx.x % p.replicas) * replica_words;

        // ---- t = 0: P_0[j] = obs[0][j] + initial[j]   (viterbi.cpp:72-76), own destination states ----
        for (int idx = tid; idx < rows * p.replicas; idx += kFlowThreads) {
            const int j = j_begin + idx % rows, k = idx / rows;
            const float v = convert_observation_rt(load_observation(p.observation, row0 * S + j, p.obs_dtype), p.flags) +
                            p.initial[j];
            st_ll(set0 + k * replica_words + j, v, base);
        }

        // the warp that finalises row r (r % 16 == warp) fetches the row's emission one step ahead in lane 0,
        // raw: it is converted where it is used, so nothing waits on the load here
        constexpr int kRounds = (kFlowMaxRows + kFlowWarps - 1) / kFlowWarps;
        float emit[kRounds];
        auto fetch_emissions = [&](int t) {
#pragma unroll
            for (int i = 0; i < kRounds; ++i) {
                const int r = warp + i * kFlowWarps;
                if (lane == 0 && r < rows && t < frames) {
                    emit[i] = load_observation(p.observation, (row0 + t) * S + j_begin + r, p.obs_dtype);
                    // ... and pulls the line of a later frame into L2, so that load never waits on HBM
                    if (t + kEmissionAhead < frames)
                        prefetch_l2(static_cast<const char *>(p.observation) +
                                    ((row0 + t + kEmissionAhead) * S + j_begin + r) * (p.obs_dtype == 0 ? 4 : 2));
                }
            }
        };
        fetch_emissions(1);

        for (int t = 1; t < frames; ++t, ++steps_done) {
            const u64 *cur = set + ((t - 1) & 1) * LDP;
            u64 *nxt = set0 + (t & 1) * LDP;                   // (replica 0; the others `replica_words` apart)
            const unsigned tag = base + (unsigned)(t - 1);
            float *my_val = pval + (steps_done & 1u) * RP * kFlowWarps;
            int *my_idx = pidx + (steps_done & 1u) * RP * kFlowWarps;

            // ---- wait for the step: thread c < #CTAs polls ONE word, the first score CTA c publishes.  If
            //      every thread polled its own 32 bytes instead, the spinning alone would draw more than half
            //      of the L2 bandwidth (148 SMs x 8 S bytes per ~500-cycle round) and every load on the chip --
            //      the publishes, the emissions, the streamed transition rows -- would queue behind it ----
            if constexpr (!kAllReg) {
                if (warp_active) {
                    // rows of the transition slice that neither registers nor shared memory hold stream from L2:
                    // two rows in flight, requested before the wait (they do not depend on the scores)
                    if (ns > 0) load_streamed(0, sbuf0);
                    if (ns > 1) load_streamed(1, sbuf1);
                }
            }
            FPROF_MARK(4)
            if (p.gate && tid < p.NC) {
                const u64 *word = cur + (size_t)tid * R;
                FLOW_WATCHDOG_DECL
                for (;;) {
                    const u64 w = ld_ll(word);
                    if (tag_of(w) == tag) break;
                    FLOW_WATCHDOG(w, -tid)
                }
            }
            if (p.gate) __syncthreads();
            FPROF_MARK(5)
            if (warp_active) {

                // ---- the scores of this thread's sources, straight from L2 into registers ----
                ulonglong2 P[KQ];
#pragma unroll
                for (int k = 0; k < KQ; ++k) {
                    P[k] = make_ulonglong2(0ull, 0ull);
                    if (valid[k] > 0) {
                        const float4 v = poll_group(cur, q[k], valid[k], tag);
                        P[k].x = ((u64)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
                        P[k].y = ((u64)__float_as_uint(v.w) << 32) | __float_as_uint(v.z);
                    }
                }
                FPROF_MARK(0)

                // ---- this thread's best candidate of one destination row and the lowest source attaining
                //      it (ascending scan with strict >, viterbi.cpp:94-101) ----
                auto row_best = [&](const ulonglong2 *a, float &m, int &first) {
                    m = -INFINITY;
                    first = 0x7fffffff;
#pragma unroll
                    for (int k = 0; k < KQ; ++k) {
                        if (valid[k] > 0) {
                            const u64 v01 = add2(P[k].x, a[k].x), v23 = add2(P[k].y, a[k].y);
                            const float v0 = lo(v01), v1 = hi(v01), v2 = lo(v23), v3 = hi(v23);
                            const float mk = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                            if (mk > m || first == 0x7fffffff) {
                                m = mk;
                                first = 4 * q[k] + (v0 == mk ? 0 : v1 == mk ? 1 : v2 == mk ? 2 : 3);
                            }
                        }
                    }
                };
                // a chunk of kChunk rows: reduced across the warp together, one lane per row stores the result
                auto store_chunk = [&](int r0, int count, float (&cv)[kChunk], int (&ci)[kChunk]) {
                    if (p.direct) {
                        u64 *mine = cand + ((size_t)(steps_done & 1u) * R + r0) * kFlowThreads + tid;
#pragma unroll
                        for (int i = 0; i < kChunk; ++i)
                            if (i < count) mine[(size_t)i * kFlowThreads] = ((u64)ordered_bits(cv[i]) << 32) | (unsigned)~ci[i];
                        return;
                    }
                    reduce_rows<kChunk, 16>(cv, ci, lane);
                    if (lane % (kWarp / kChunk) == 0 && lane / (kWarp / kChunk) < count) {
                        const int r = r0 + lane / (kWarp / kChunk);
                        my_val[r * kFlowWarps + warp] = cv[0];
                        my_idx[r * kFlowWarps + warp] = ci[0];
                    }
                };
                // rows streamed from L2 fill their own chunk(s) as they arrive: one is consumed, and the row
                // after next requested, every `stream_stride` resident rows
                float sv[kChunk];
                int si[kChunk];
#pragma unroll
                for (int i = 0; i < kChunk; ++i) { sv[i] = -INFINITY; si[i] = 0x7fffffff; }
                int retired = 0;
                auto retire = [&]() {
                    if constexpr (kAllReg) return;
                    if (retired >= ns) return;
                    float m;
                    int first;
                    if (retired & 1) {
                        row_best(sbuf1, m, first);
                        if (retired + 2 < ns) load_streamed(retired + 2, sbuf1);
                    } else {
                        row_best(sbuf0, m, first);
                        if (retired + 2 < ns) load_streamed(retired + 2, sbuf0);
                    }
#pragma unroll
                    for (int i = 0; i < kChunk; ++i)
                        if (i == retired % kChunk) { sv[i] = m; si[i] = first; }
                    ++retired;
                    if (retired % kChunk == 0 || retired == ns) {
                        store_chunk(first_streamed + (retired - 1) / kChunk * kChunk, (retired - 1) % kChunk + 1, sv, si);
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) { sv[i] = -INFINITY; si[i] = 0x7fffffff; }
                    }
                };
                retire();
                // register rows, chunk by chunk
#pragma unroll
                for (int r0 = 0; r0 < RR; r0 += kChunk) {
                    if (r0 < rows) {
                        float cv[kChunk];
                        int ci[kChunk];
#pragma unroll
                        for (int i = 0; i < kChunk; ++i) {
                            cv[i] = -INFINITY;
                            ci[i] = 0x7fffffff;
                            if (r0 + i < RR && r0 + i < rows) row_best(a_reg[r0 + i < RR ? r0 + i : 0], cv[i], ci[i]);
                        }
                        store_chunk(r0, min(kChunk, min(RR, rows) - r0), cv, ci);
                    }
                }
                // shared-memory rows
                for (int rs0 = 0; !kAllReg && rs0 < smem_rows; rs0 += kChunk) {
                    float cv[kChunk];
                    int ci[kChunk];
#pragma unroll
                    for (int i = 0; i < kChunk; ++i) {
                        cv[i] = -INFINITY;
                        ci[i] = 0x7fffffff;
                        if (rs0 + i < smem_rows) {
                            ulonglong2 a[KQ];
                            const ulonglong2 *arow = reinterpret_cast<const ulonglong2 *>(As + (rs0 + i) * LDP);
#pragma unroll
                            for (int k = 0; k < KQ; ++k) a[k] = valid[k] > 0 ? arow[q[k]] : minus_inf;
                            row_best(a, cv[i], ci[i]);
                            if ((rs0 + i + 1) % stream_stride == 0) retire();
                        }
                    }
                    store_chunk(RR + rs0, min(kChunk, smem_rows - rs0), cv, ci);
                }
                while (!kAllReg && retired < ns) retire();
                FPROF_MARK(1)
            }
            __syncthreads();        // the only CTA barrier of a step

            // ---- across the warps; publish P_t[j] = obs[t][j] + max  (viterbi.cpp:102) and the backpointer ----
#pragma unroll
            for (int i = 0; i < kRounds; ++i) {
                const int r = warp + i * kFlowWarps;
                if (r < rows) {
                    float best;
                    int source;
                    if (p.direct) {
                        // every thread's candidate of this row: 32 x active_warps keys, a few per lane, then a butterfly
                        const u64 *keys = cand + ((size_t)(steps_done & 1u) * R + r) * kFlowThreads + lane;
                        u64 key = 0;
#pragma unroll
                        for (int w = 0; w < kFlowWarps; ++w)          // (all loads issued before the first compare)
                            if (w < active_warps) key = max(key, keys[w * kWarp]);
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) key = max(key, __shfl_xor_sync(kFullMask, key, off));
                        best = from_ordered_bits((unsigned)(key >> 32));
                        source = (int)~(unsigned)key;
                    } else {
                        best = lane < active_warps ? my_val[r * kFlowWarps + lane] : -INFINITY;
                        source = lane < active_warps ? my_idx[r * kFlowWarps + lane] : 0x7fffffff;
#pragma unroll
                        for (int off = kFlowWarps / 2; off > 0; off >>= 1)
                            take_better(best, source, __shfl_xor_sync(kFullMask, best, off), __shfl_xor_sync(kFullMask, source, off));
                    }
                    const int j = j_begin + r;
                    const float v = convert_observation_rt(__shfl_sync(kFullMask, emit[i], 0), p.flags) + best;
                    for (int k = lane; k < p.replicas; k += kWarp) st_ll(nxt + k * replica_words + j, v, base + (unsigned)t);
                    if (lane == 0) {
                        p.backptr[(hrow0 + t) * S + j] = (uint16_t)min(source, S - 1);
                        if (p.checkpoints && t % p.checkpoint_every == 0)
                            p.checkpoints[((int64_t)item * p.ck_rows + (t / p.checkpoint_every - 1)) * S + j] = v;
                    }
                }
            }
            fetch_emissions(t + 1);
            FPROF_MARK(2)
        }
        FPROF_PRINT(max(frames - 1, 1))

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

// ---------------------------------------------------------------------------
// The same recursion for the shapes that matter most to it (config 1: 1440 states): one float4 source group per
// thread (S <= 2048) and every transition row of the CTA in registers.  A dedicated, branch-free step: the general
// kernel above spends ~680 instructions per warp and step of which 40 are the max-plus math -- the rest is index
// arithmetic, predication and divergence bookkeeping of its register / shared-memory / streamed row paths (ncu
// source view of profiles/r02_forward_flow_ncu_full.json: IMAD 1408, ISETP 1143, SEL 792, BRA 691, BSSY + BSYNC 816
// against FADD2 240 per CTA and step).  Here a thread computes its RR candidates unconditionally (rows the CTA does
// not own hold -inf), stores one 64-bit key per row -- (order-preserving score bits << 32) | ~source, so an integer
// maximum is (score desc, source asc) -- and after the ONE CTA barrier the warp that owns a row reduces its 512 keys.
// ---------------------------------------------------------------------------
template <int RR>
__global__ void __launch_bounds__(kFlowThreads, 1) forward_flow_direct_kernel(const FlowParams p) {
    extern __shared__ __align__(16) unsigned flow_smem[];
    u64 *cand = reinterpret_cast<u64 *>(flow_smem);               // [2][RR][kFlowThreads] candidate keys
    __shared__ float red_val[kFlowWarps];
    __shared__ int red_idx[kFlowWarps];
    constexpr int kRounds = (RR + kFlowWarps - 1) / kFlowWarps;   // rows a warp finalises

    const int S = p.S, LDP = p.LDP, R = p.R;
    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int j_begin = blockIdx.x * R;
    const int rows = max(0, min(R, S - j_begin));
    const int valid = max(0, min(4, S - 4 * tid));                // real states among this thread's four sources
    const bool warp_active = 4 * (warp * kWarp) < S;              // (warps whose lanes own no source skip the scan)

    ulonglong2 a_reg[RR];
    const ulonglong2 minus_inf = make_ulonglong2(0xff800000ff800000ull, 0xff800000ff800000ull);
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        a_reg[r] = minus_inf;
        if (r < rows && 4 * tid < LDP)
            a_reg[r] = __ldg(reinterpret_cast<const ulonglong2 *>(p.transition + (int64_t)(j_begin + r) * LDP) + tid);
    }
    // keys of threads that never store (idle warps) stay 0: below every real key
    for (int i = tid; i < 2 * RR * kFlowThreads; i += kFlowThreads) cand[i] = 0ull;
    FPROF_DECL
    __syncthreads();

    unsigned base = 0, steps_done = 0;
    for (int n = 0; n < p.B; ++n) {
        const int item = p.order[n];
        const int frames = p.frames[item];
        const int64_t row0 = p.item_row[item];
        const int64_t hrow0 = p.hist_row[item];
        u64 *set = p.ll + (size_t)(n & 1) * 2 * LDP;

        // ---- t = 0: P_0[j] = obs[0][j] + initial[j]   (viterbi.cpp:72-76), own destination states ----
        if (tid < rows) {
            const int j = j_begin + tid;
            st_ll(set + j, convert_observation_rt(load_observation(p.observation, row0 * S + j, p.obs_dtype), p.flags) +
                               p.initial[j], base);
        }
        float emit[kRounds];
        auto fetch_emissions = [&](int t) {
#pragma unroll
            for (int i = 0; i < kRounds; ++i) {
                const int r = warp + i * kFlowWarps;
                if (lane == 0 && r < rows && t < frames) {
                    emit[i] = load_observation(p.observation, (row0 + t) * S + j_begin + r, p.obs_dtype);
                    if (t + kEmissionAhead < frames)
                        prefetch_l2(static_cast<const char *>(p.observation) +
                                    ((row0 + t + kEmissionAhead) * S + j_begin + r) * (p.obs_dtype == 0 ? 4 : 2));
                }
            }
        };
        fetch_emissions(1);

        for (int t = 1; t < frames; ++t, ++steps_done) {
            const u64 *cur = set + ((t - 1) & 1) * LDP;
            u64 *nxt = set + (t & 1) * LDP;
            const unsigned tag = base + (unsigned)(t - 1);
            u64 *keys = cand + (size_t)(steps_done & 1u) * RR * kFlowThreads;
            FPROF_MARK(4)
            // gate: thread c < #CTAs waits for the first score CTA c publishes, then everyone loads once
            if (tid < p.NC) {
                const u64 *word = cur + (size_t)tid * R;
                FLOW_WATCHDOG_DECL
                for (;;) {
                    const u64 w = ld_ll(word);
                    if (tag_of(w) == tag) break;
                    FLOW_WATCHDOG(w, -tid)
                }
            }
            __syncthreads();
            FPROF_MARK(5)
            if (warp_active) {
                ulonglong2 P = make_ulonglong2(0ull, 0ull);
                if (valid > 0) {
                    const float4 v = poll_group(cur, tid, valid, tag);
                    P.x = ((u64)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
                    P.y = ((u64)__float_as_uint(v.w) << 32) | __float_as_uint(v.z);
                }
                FPROF_MARK(0)
                const unsigned not_source = ~(unsigned)(4 * tid);
#pragma unroll
                for (int r = 0; r < RR; ++r) {
                    const u64 v01 = add2(P.x, a_reg[r].x), v23 = add2(P.y, a_reg[r].y);
                    const float v0 = lo(v01), v1 = hi(v01), v2 = lo(v23), v3 = hi(v23);
                    const float mk = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                    // ~(4 tid + i) = ~(4 tid) - i; i = first of the four attaining the maximum (ascending, viterbi.cpp:94-101)
                    const unsigned low = not_source - (v0 == mk ? 0u : v1 == mk ? 1u : v2 == mk ? 2u : 3u);
                    keys[(size_t)r * kFlowThreads + tid] = ((u64)ordered_bits(mk) << 32) | low;
                }
                FPROF_MARK(1)
            }
            __syncthreads();        // the only other CTA barrier of a step

#pragma unroll
            for (int i = 0; i < kRounds; ++i) {
                const int r = warp + i * kFlowWarps;
                if (r < rows) {
                    const u64 *col = keys + (size_t)r * kFlowThreads + lane;
                    u64 k[kFlowWarps];
#pragma unroll
                    for (int w = 0; w < kFlowWarps; ++w) k[w] = col[w * kWarp];
                    u64 key = k[0];
#pragma unroll
                    for (int w = 1; w < kFlowWarps; ++w) key = max(key, k[w]);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) key = max(key, __shfl_xor_sync(kFullMask, key, off));
                    const float best = from_ordered_bits((unsigned)(key >> 32));
                    const int source = (int)~(unsigned)key;
                    const int j = j_begin + r;
                    const float v = convert_observation_rt(__shfl_sync(kFullMask, emit[i], 0), p.flags) + best;
                    if (lane == 0) {
                        st_ll(nxt + j, v, base + (unsigned)t);
                        p.backptr[(hrow0 + t) * S + j] = (uint16_t)min(source, S - 1);
                        if (p.checkpoints && t % p.checkpoint_every == 0)
                            p.checkpoints[((int64_t)item * p.ck_rows + (t / p.checkpoint_every - 1)) * S + j] = v;
                    }
                }
            }
            fetch_emissions(t + 1);
            FPROF_MARK(2)
        }
        FPROF_PRINT(max(frames - 1, 1))

        // ---- every CTA waits for the final scores, one writes them out (as in the general kernel) ----
        {
            const u64 *fin = set + ((frames - 1) & 1) * LDP;
            const unsigned tag = base + (unsigned)(frames - 1);
            const bool writer = (int)blockIdx.x == n % p.NC;
            float best = -INFINITY;
            int best_i = 0x7fffffff;
            if (valid > 0) {
                const float4 v = poll_group(fin, tid, valid, tag);
                if (writer) {
                    const float e[4] = {v.x, v.y, v.z, v.w};
                    for (int x = 0; x < valid; ++x) {
                        const int s = 4 * tid + x;
                        if (p.posterior) p.posterior[(int64_t)item * S + s] = e[x];
                        if (better(e[x], s, best, best_i)) { best = e[x]; best_i = s; }
                    }
                }
            }
            __syncthreads();        // (orders the reuse of the buffer set, see the general kernel)
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
    int NC, R, KQ, RR, smem_rows, direct;
    size_t smem;
    bool ok;
};

// register rows per instantiation: KQ x RR x 4 floats of the transition block stay in registers
constexpr int flow_reg_rows(int kq) { return kq == 1 ? 20 : kq == 2 ? 8 : kq == 4 ? 3 : 1; }

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
    // dynamic shared memory: the double-buffered per-warp partials (2 x 2 x R x 16 words), then transition rows
    const size_t partials = sizeof(unsigned) * 4 * ((g.R + kChunk - 1) / kChunk * kChunk) * kFlowWarps;
    const size_t budget = 226 * 1024 - partials;
    const int64_t want = std::max<int64_t>(0, g.R - g.RR);
    g.smem_rows = (int)std::min<int64_t>(want, (int64_t)(budget / (sizeof(float) * ldp)));
    g.smem = partials + sizeof(float) * ldp * g.smem_rows;
    // direct candidates instead of per-warp partials where every row sits in registers and the keys fit
    const size_t direct_bytes = sizeof(u64) * 2 * g.R * kFlowThreads;
    g.direct = g.KQ == 1 && g.R <= g.RR && direct_bytes <= 160 * 1024;
    if (g.direct) g.smem = std::max(g.smem, direct_bytes);
    g.ok = groups <= 8 && g.R <= kFlowMaxRows;
    return g;
}

}  // namespace

bool forward_flow_supported(const ForwardArgs &a) {
    return flow_geometry(a.S, a.sm_count).ok && a.ws.transition_pad != nullptr;
}

// Estimated seconds, from B200 measurements (profiles/r02_flow_kernel.txt; one sequence, us per step):
// S = 640: 2.02, 1440: 2.53, 2048: 2.82 -- against 2.6 / 3.2 / 3.9 for the wide kernel -- but 13 at 4096 (rows
// beyond the registers) and 4.4 at 256 (two busy warps per CTA).  Sequences run one after the other.
double flow_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count) {
    const FlowGeometry g = flow_geometry(S, sm_count);
    double step_us;
    if (S < 512) step_us = 4.4;
    else if (g.direct && S <= 1536) step_us = 1.55 + 0.00068 * (double)S;
    else if (g.direct) step_us = 2.85;
    else if (g.KQ == 1) step_us = 1.0 + 0.0017 * (double)S;
    else step_us = 2.0 + 0.0027 * (double)S * g.KQ * g.KQ / 4.0 + 0.6 * std::max(0, g.R - g.RR - g.smem_rows);
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
    p.direct = g.direct;
    if (const char *env = std::getenv("TORBI_B200_FLOW_DIRECT")) p.direct = g.direct && std::atoi(env) != 0;
    p.gate = 1;
    if (const char *env = std::getenv("TORBI_B200_FLOW_GATE")) p.gate = std::atoi(env) != 0;
    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    // tags start out invalid (0xffffffff is never a step of this call): 2 sets x 2 parities x LDP words,
    // at the head of the forward-score region (128 x LDP floats per buffer: more than enough)
    // K replicas of (2 sets x 2 parities x LDP words) at the head of the forward-score region (4 buffers of
    // pad_batch(B) x LDP floats, contiguous: room for 64 replicas at the smallest batch)
    int replicas = 1;
    if (const char *env = std::getenv("TORBI_B200_FLOW_REPLICAS")) replicas = std::atoi(env);
    const size_t room = (size_t)(reinterpret_cast<char *>(a.ws.block_max) - reinterpret_cast<char *>(a.ws.score[0]));
    replicas = std::max(1, std::min<int>(replicas, (int)std::min<size_t>(g.NC, room / (sizeof(u64) * 4 * p.LDP))));
    p.replicas = replicas;
    e = cudaMemsetAsync(p.ll, 0xff, sizeof(u64) * 4 * p.LDP * replicas, stream);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&p};
    auto launch = [&](auto kernel) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (err != cudaSuccess) return err;
        // cooperative: every CTA polls every other CTA's stores, so all must be co-resident
        return cudaLaunchCooperativeKernel((void *)kernel, dim3(g.NC), dim3(kFlowThreads), args, g.smem, stream);
    };
    const bool all_reg = g.R <= g.RR;
    // The dedicated kernel halves the instruction count (scan 1270 -> 620 cycles, key reduction + publish 2200 -> 1330)
    // yet the step only gets shorter where a CTA owns more than 10 rows: S = 2048 3.24 -> 2.82 us, but S = 1440
    // 2.52 -> 2.71 and S = 640 2.02 -> 2.15 us (the wait moves into the score loads: the gate opens on a producer's
    // FIRST row and the others now trail it), profiles/r02_flow_kernel.txt.  So it serves RR = 20 only;
    // TORBI_B200_FLOW_DIRECT_KERNEL=1 / 0 forces it on (where possible) / off.
    const bool can_dedicate = p.direct && all_reg && p.replicas == 1 && p.gate && g.KQ == 1;
    bool dedicated = can_dedicate && g.RR == 20;
    if (const char *env = std::getenv("TORBI_B200_FLOW_DIRECT_KERNEL")) dedicated = can_dedicate && std::atoi(env) != 0;
    if (dedicated) {
        const size_t bytes = sizeof(u64) * 2 * g.RR * kFlowThreads;
        auto launch_direct = [&](auto kernel) {
            cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (err != cudaSuccess) return err;
            return cudaLaunchCooperativeKernel((void *)kernel, dim3(g.NC), dim3(kFlowThreads), args, bytes, stream);
        };
        if (g.RR == 4) return launch_direct(forward_flow_direct_kernel<4>);
        if (g.RR == 10) return launch_direct(forward_flow_direct_kernel<10>);
        return launch_direct(forward_flow_direct_kernel<20>);
    }
    if (g.KQ == 1 && g.RR == 4) return all_reg ? launch(forward_flow_kernel<1, 4, true>) : launch(forward_flow_kernel<1, 4, false>);
    if (g.KQ == 1 && g.RR == 10) return all_reg ? launch(forward_flow_kernel<1, 10, true>) : launch(forward_flow_kernel<1, 10, false>);
    if (g.KQ == 1) return all_reg ? launch(forward_flow_kernel<1, 20, true>) : launch(forward_flow_kernel<1, 20, false>);
    if (g.KQ == 2) return launch(forward_flow_kernel<2, 8, false>);
    if (g.KQ == 4) return launch(forward_flow_kernel<4, 3, false>);
    return launch(forward_flow_kernel<8, 1, false>);
}

}  // namespace torbi_b200
