// Batch-tiled forward kernel: the max-plus (tropical) product
//     V[b][j] = max_s ( P[b][s] + A[j][s] ),  argmax with lowest-s ties
// done as a register-tiled "GEMM" per timestep on the FP32/ALU pipes (FADD2 +
// 3-input FMNMX; not tensor cores -- max-plus is not a multiply-add contraction).
//
// Work decomposition (persistent, cooperative launch, one CTA per SM):
//   * sequences are processed longest-first in batches of TM = 64;
//   * a GROUP of NJT = ceil(S / TN) CTAs owns one batch at a time; CTA `jt` of the
//     group owns the TN = 80 destination states [jt*TN, jt*TN+TN) for all T steps;
//   * per step every CTA streams its A slice [TN x S] and the batch's previous
//     scores P[TM x S] from L2 through a 3-stage cp.async ring (96 source states
//     per stage) synchronised by mbarriers (full: copies landed, empty: warps done;
//     no CTA-wide barrier per stage), so A is re-read once per 64 sequences instead
//     of once per sequence as in the reference kernel (viterbi.cu:91,97);
//   * the new scores are published to global memory (L2) and the group meets at
//     a counter barrier; groups never wait for each other.
//
// Exactness (see DESIGN.md "Numerics"): every candidate is the single fp32 add
// P[s] + A[j][s]; max is order independent; the argmax is found in two levels --
// per 32-wide chunk of source states a strict `>` against the running maximum
// records the FIRST chunk attaining the final maximum, then that chunk is
// re-scanned for the lowest s with P[s] + A[j][s] == max.  This reproduces the
// reference's ascending strict-`>` scan (viterbi.cpp:94-101) bit for bit.
//
// Two builds of this file: forward_tiled.cu (backpointer mode: two-level argmax, uint16
// backpointers, as described above) and forward_tiled_hist.cu, which defines
// TORBI_TILED_HIST and compiles the SCORE-HISTORY mode: the K loop is a pure max-plus
// product (no chunk bookkeeping, no re-scan), every P_t is also stored to an fp32
// (B,T,S) history, and backtrace.cu recomputes the one argmax per step that lies on the
// decoded path (exact: the same single fp32 add P_{t-1}[s] + A[j][s]).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "sync.cuh"

#ifndef TORBI_TILED_HIST
#define TORBI_TILED_HIST 0
#endif

namespace torbi_b200 {

namespace {

constexpr bool kHist = TORBI_TILED_HIST;

// Per-phase cycle counters of the step loop (thread 0 of a few CTAs prints them at the end);
// compiled in only by tools/profile_phases.sh, never in the shipped library.
#ifdef TORBI_TILED_PROFILE
#define PROF_DECL long long prof_t0 = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_MARK(i) { const long long prof_now = clock64(); prof_acc[i] += prof_now - prof_t0; prof_t0 = prof_now; }
#define PROF_PRINT(steps)                                                                          \
    if (threadIdx.x == 0 && blockIdx.x % 37 == 0)                                                  \
        printf("cta %d steps %d cycles/step: first-stage wait %lld, K loop %lld, epilogue %lld (new scores %lld + "  \
               "CTA barrier %lld + publish %lld + arrive %lld), prefetch(+rescan, backpointers) %lld, exchange wait %lld\n", \
               blockIdx.x, (steps), prof_acc[0] / (steps), prof_acc[1] / (steps),                               \
               (prof_acc[2] + prof_acc[5] + prof_acc[6] + prof_acc[7]) / (steps), prof_acc[5] / (steps),          \
               prof_acc[6] / (steps), prof_acc[7] / (steps), prof_acc[2] / (steps), prof_acc[3] / (steps), prof_acc[4] / (steps));
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_PRINT(steps)
#endif
constexpr int TN = kTileDests;  // destination states per CTA (columns of the tile)
constexpr int KC = 32;          // source states per argmax chunk
constexpr int KS = kStageSources;  // source states per pipeline stage (3 chunks)
constexpr int LDS_ = KS + 4;    // smem row stride in floats: rows 4 banks apart
constexpr int NSTAGE = 3;
constexpr int kThreads = 256;
constexpr int N = 5;            // register tile: M sequences x 5 destination states
constexpr int LANES_B = 8, LANES_J = 4, WARPS_J = 4, WARPS_B = 2;

// The batch tile height is a template parameter: M = 4, 2, 1 rows per thread
// give TM = 64, 32, 16 sequences per batch (smaller tiles fill the machine when
// the batch is small or ragged, at a lower math/LDS ratio).
constexpr int tile_rows(int M) { return WARPS_B * LANES_B * M; }
constexpr int stage_floats(int M) { return (tile_rows(M) + TN) * LDS_; }
constexpr size_t smem_bytes(int M) {
    return sizeof(float) * (NSTAGE * stage_floats(M) + 2 * tile_rows(M) * TN)  // ring + emission/score + max tiles
           + sizeof(uint16_t) * tile_rows(M) * TN                              // winning chunk -> backpointer tile
           + sizeof(int) * 8 * tile_rows(M)                                    // per-slot item id, length, first rows (2 x int64); banded: outside max / index
           + sizeof(u64) * 2 * NSTAGE;                                         // full / empty mbarrier per ring slot
}
static_assert(smem_bytes(4) <= 227 * 1024, "shared memory budget");
constexpr int kLoaders = 240;                 // 24 float4 columns x 10 rows per pass
static_assert(KS / 4 == 24 && TN % 10 == 0, "loader mapping");

__device__ __forceinline__ u64 add2(u64 a, u64 b) {     // two fp32 adds (FADD2), IEEE rn
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    // .cg: L2 only -- the score buffers are rewritten every step by other SMs
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending));
}

__device__ __forceinline__ int ld_acquire_sys(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Group barrier in two halves so independent work can sit between them:
// arrive = "my stores of this step are published", wait = "everyone's are".
__device__ __forceinline__ void group_arrive(unsigned *counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
    }
}
__device__ __forceinline__ void group_wait(const unsigned *counter, unsigned target) {
    if (threadIdx.x == 0) {
        while (ld_acquire(counter) < target) {}
    }
    __syncthreads();
}

struct TiledParams {
    const float *observation;   // (rows,S), item b at rows item_row[b] ..; 2-byte elements when obs_dtype != 0
    int obs_dtype;              // 0 fp32, 1 fp16, 2 bf16
    const float *transition;    // padded copy (SJ, LDP), -inf outside (S,S)
    const float *initial;       // (S,)
    const int32_t *frames;      // (B,)
    const int64_t *item_row;    // (B,) first row of each item in observation
    const int64_t *hist_row;    // (B,) first row of each item in history / backptr (compact: sum(frames) rows)
    const int32_t *order;       // (Bpad,) items longest first, -1 = empty slot
    float *score0;              // 3 rotating buffers (Bpad, LDP), `score_stride` floats apart;
    int64_t score_stride;       // row = position in `order`
    uint16_t *backptr;          // (rows,S)
    float *history;             // score-history mode: (rows,S) fp32, every P_t (same memory as backptr)
    float *posterior;           // (B,S) or null
    int32_t *final_state;       // (B,)
    unsigned *sync;             // per group: barrier counter (+0) and current batch (+1); queue head at [0] of the last line
    const int32_t *frames_ready;  // frames of `observation` copied in so far, or null
    float *block_max;             // banded score-history mode: 3 rotating (Bpad, LDP/16) arrays, max of every
    int64_t block_stride;         // 16-state block of the corresponding score row; `block_stride` floats apart
    const unsigned *band;         // banded mode: per destination tile S - first / last+1 source differing from `outside`
    const float *outside;         // banded mode: the value every transition entry outside the windows has
    int S, LDP, NJT, NG, num_batches;
};

// Issue the cp.async copies of one pipeline stage.  Thread `tid` < 240 owns the
// float4 column tid % 24 of rows tid / 24 + 10*q (8 A rows + 7 P rows per thread),
// so every address is a fixed per-thread base plus a constant stride.  The 15
// copies are issued in three parts, one per argmax chunk, so they interleave
// with the math instead of piling up in the load/store queue.
template <int TM, int kPart>
__device__ __forceinline__ void load_stage_part(float *stage, const float *gA, const float *gP,
                                                int64_t row_stride10, int r0) {
    if (threadIdx.x < kLoaders) {
        float *sP = stage, *sA = stage + TM * LDS_;
#pragma unroll
        for (int q = 0; q < TN / 10; ++q)
            if (q / 5 == kPart)                                   // part 0: A rows q = 0..4, part 1: q = 5..7
                cp_async16(sA + q * 10 * LDS_, gA + q * row_stride10);
#pragma unroll
        for (int q = 0; q < (TM + 9) / 10; ++q)
            if ((q + 8) / 5 == kPart && (q * 10 + 9 < TM || r0 + q * 10 < TM))   // part 1: P q = 0..1, part 2: q = 2..6
                cp_async16(sP + q * 10 * LDS_, gP + q * row_stride10);
    }
}
// whole-stage halves used at the step boundary: the transition rows do not depend on
// the group's score exchange and are issued before the barrier wait, the score rows after
template <int TM>
__device__ __forceinline__ void load_stage_A(float *stage, const float *gA, int64_t row_stride10) {
    if (threadIdx.x < kLoaders) {
        float *sA = stage + TM * LDS_;
#pragma unroll
        for (int q = 0; q < TN / 10; ++q) cp_async16(sA + q * 10 * LDS_, gA + q * row_stride10);
    }
}
template <int TM>
__device__ __forceinline__ void load_stage_P(float *stage, const float *gP, int64_t row_stride10, int r0) {
    if (threadIdx.x < kLoaders) {
#pragma unroll
        for (int q = 0; q < (TM + 9) / 10; ++q)
            if (q * 10 + 9 < TM || r0 + q * 10 < TM) cp_async16(stage + q * 10 * LDS_, gP + q * row_stride10);
    }
}
template <int TM>
__device__ __forceinline__ void load_stage(float *stage, const float *gA, const float *gP,
                                           int64_t row_stride10, int r0) {
    load_stage_part<TM, 0>(stage, gA, gP, row_stride10, r0);
    load_stage_part<TM, 1>(stage, gA, gP, row_stride10, r0);
    load_stage_part<TM, 2>(stage, gA, gP, row_stride10, r0);
}

// Exact argmax for the whole tile: lowest s in the winning chunk with
// P[s] + A[j][s] == max.  8 lanes share one output (a float4 of the 32-wide chunk
// each) so every load is a coalesced 128-byte row; kBatch outputs are in flight
// per lane group (all shared-memory reads, then all global loads, then the
// compares -- nothing in a batch waits on an earlier output's store).
// bp_tile holds the winning chunk on entry and the source index on exit.
template <int TM, bool kFullTile, bool kBanded>
__device__ __forceinline__ void rescan_tile(const float *best_tile, uint16_t *bp_tile,
                                            const int *slot_frames, const float *P_rows,
                                            const float *A_rows, int LDP, int cols_valid, int S, int t,
                                            const float *out_v, const int *out_i) {
    constexpr int kBatch = 20;
    constexpr int kRowsPerWarp = TM / (kThreads / kWarp);
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    const int sub = lane >> 3, q4 = (lane & 7) * 4;
    const float *Aq = A_rows + (int64_t)sub * LDP + q4;            // column `sub` of each group of 4
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = warp * kRowsPerWarp + rr;
        if (t >= slot_frames[r]) continue;                          // warp-uniform
        const float *Pq = P_rows + (int64_t)r * LDP + q4;
        const float *best_row = best_tile + r * TN + sub;
        uint16_t *bp_row = bp_tile + r * TN + sub;
#pragma unroll 1
        for (int cb = 0; cb < TN / 4; cb += kBatch) {
            float best[kBatch];
            int k0[kBatch];
            float4 pv[kBatch], av[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                best[u] = best_row[(cb + u) * 4];
                k0[u] = bp_row[(cb + u) * 4] * KC;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const bool valid = kFullTile || (cb + u) * 4 + sub < cols_valid;
                if (!valid) k0[u] = 0;                              // stale tile entry: stay in bounds
                const int a_off = valid ? (cb + u) * 4 * LDP + k0[u] : 0;
                pv[u] = __ldcg(reinterpret_cast<const float4 *>(Pq + k0[u]));
                av[u] = __ldg(reinterpret_cast<const float4 *>(Aq + a_off));
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                unsigned key = 64;
                if (pv[u].w + av[u].w == best[u]) key = q4 + 3;
                if (pv[u].z + av[u].z == best[u]) key = q4 + 2;
                if (pv[u].y + av[u].y == best[u]) key = q4 + 1;
                if (pv[u].x + av[u].x == best[u]) key = q4;
                // min over the 8 lanes of the output (xor 1, 2, 4 stays inside the group)
                key = min(key, __shfl_xor_sync(kFullMask, key, 1));
                key = min(key, __shfl_xor_sync(kFullMask, key, 2));
                key = min(key, __shfl_xor_sync(kFullMask, key, 4));
                // (key == 64 only for NaN input, outside the parity contract)
                k0[u] = min(k0[u] + (int)key, S - 1);
                if (kBanded) {
                    // sources outside the tile's window: larger value wins, equal values keep the lower index
                    const float ov = out_v[r];
                    const int oi = out_i[r];
                    if (ov > best[u] || (ov == best[u] && oi < k0[u])) k0[u] = oi;
                }
            }
            __syncwarp();       // every lane has read its chunk ids before they are overwritten
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
                if ((lane & 7) == 0 && (kFullTile || (cb + u) * 4 + sub < cols_valid))
                    bp_row[(cb + u) * 4] = (uint16_t)k0[u];
            __syncwarp();
        }
    }
}

#if TORBI_TILED_HIST
#define forward_tiled_kernel forward_tiled_hist_kernel      // distinct name in profiles
#endif
template <int M, uint32_t kFlags, bool kBanded>
__global__ void __launch_bounds__(kThreads, 1) forward_tiled_kernel(const TiledParams p) {
    constexpr int TM = tile_rows(M);
    constexpr int kStageFloats = stage_floats(M);
    extern __shared__ __align__(16) float smem[];
    float *ring = smem;
    float *tile = smem + NSTAGE * kStageFloats;                      // [TM][TN] emission -> new score
    float *best_tile = tile + TM * TN;                               // [TM][TN] max over sources
    uint16_t *bp_tile = reinterpret_cast<uint16_t *>(best_tile + TM * TN);  // [TM][TN] chunk, then index
    int *slot_item = reinterpret_cast<int *>(bp_tile + TM * TN);       // [TM]
    int *slot_frames = slot_item + TM;                                 // [TM]
    float *out_v = reinterpret_cast<float *>(slot_frames + TM);        // [TM] banded: max over sources outside the window
    int *out_i = reinterpret_cast<int *>(out_v + TM);                  // [TM]          and its lowest index
    // Ring synchronisation without CTA-wide barriers: full[slot] collects one arrival per thread
    // once its async copies for the stage have landed; empty[slot] one per warp once it has
    // finished reading the stage.  Warps may drift apart by up to a stage.
    int64_t *slot_row = reinterpret_cast<int64_t *>(out_i + TM);       // [TM] first row of the slot's item
    int64_t *slot_hrow = slot_row + TM;                                // [TM] its first row in the history / backpointers
    u64 *full_bar = reinterpret_cast<u64 *>(slot_hrow + TM);           // [NSTAGE]
    u64 *empty_bar = full_bar + NSTAGE;                                // [NSTAGE]
    if (threadIdx.x == 0) {
        for (int k = 0; k < NSTAGE; ++k) {
            mbar_init(full_bar + k, kThreads);
            mbar_init(empty_bar + k, kThreads / kWarp);
        }
        mbar_fence_init();
    }
    __syncthreads();
    unsigned stage_counter = 0;      // pipeline stages consumed so far by this thread (slot = counter % NSTAGE)
    unsigned empty_parity = 0;       // bit k: parity of the next phase of empty_bar[k] this thread will wait for
    __shared__ int ready_shared;
    int ready = p.frames_ready ? 0 : 0x7fffffff;    // frames of `observation` known to be in HBM
    // the host entry copies the batch in T-chunks while this kernel runs: block
    // until frame `t` has landed (one poll per chunk, not per step)
    auto wait_for_frame = [&](int t) {
        if (t >= ready) {
            if (threadIdx.x == 0) {
                int r;
                while ((r = ld_acquire_sys(p.frames_ready)) <= t) __nanosleep(200);
                ready_shared = r;
            }
            __syncthreads();
            ready = ready_shared;
            __syncthreads();
        }
    };

    const int group = blockIdx.x / p.NJT, jt = blockIdx.x % p.NJT;
    if (group >= p.NG) return;
    const int j0 = jt * TN;
    unsigned *counter = p.sync + group * 32;                           // own 128-byte line
    unsigned barriers = 0;

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int wb = warp / WARPS_J, wj = warp % WARPS_J;
    const int lane_b = lane / LANES_J, lane_j = lane % LANES_J;
    const int r_base = wb * (LANES_B * M) + lane_b;     // rows r_base + 8*i
    const int c_base = wj * (LANES_J * N) + lane_j;     // cols c_base + 4*c
    const int S = p.S, LDP = p.LDP;
    // Pipeline stages this CTA runs per step.  Dense: all of them.  Banded: only the window
    // outside which every transition entry of this destination tile equals *p.outside (the
    // device-side analysis, analyze_band_kernel, guarantees it for any matrix).
    // The window is kept at chunk (32-state) granularity: whole stages are staged, but the chunks of
    // the first / last stage that lie outside the window are skipped by the math.
    int s_lo = 0, s_hi = LDP / KS;
    int c_lo = 0, c_hi = LDP / KC;
    float outside = 0.f;
    if (kBanded) {
        const int lo = S - (int)p.band[jt], hi = (int)p.band[p.NJT + jt];
        c_lo = lo / KC;
        c_hi = (hi + KC - 1) / KC;
        if (lo >= hi) c_lo = c_hi = 0;
        s_lo = c_lo / (KS / KC);
        s_hi = (c_hi + KS / KC - 1) / (KS / KC);
        outside = *p.outside;
    }
    const int nstages = s_hi - s_lo;
    const int stage_base = s_lo * KS;                             // first source state staged
    const int k_lo = c_lo * KC, k_hi = min(c_hi * KC, S);         // sources covered by the dense loop
    const bool obs_vec = p.obs_dtype == 0 && (S % 4) == 0 && (reinterpret_cast<uintptr_t>(p.observation) & 15) == 0;
    const bool bp_vec = (S % 8) == 0;
    const bool hist_vec = (S % 4) == 0;
    // loader mapping (see load_stage)
    const int ld_c4 = tid % (KS / 4), ld_r0 = tid / (KS / 4);
    const int ld_smem = ld_r0 * LDS_ + ld_c4 * 4;
    const float *ld_A = p.transition + (int64_t)(j0 + ld_r0) * LDP + ld_c4 * 4 + stage_base;
    const int64_t ld_stride10 = (int64_t)10 * LDP;

    unsigned *queue_head = p.sync + p.NG * 32;       // batches are handed out longest first
    for (;;) {
        // the group's first CTA takes the next batch off the queue; the group
        // barrier publishes it to the others (dynamic LPT over groups)
        if (jt == 0 && tid == 0) counter[1] = atomicAdd(queue_head, 1u);
        barriers += p.NJT;
        group_arrive(counter);
        group_wait(counter, barriers);
        const int batch = (int)__ldcg(counter + 1);
        if (batch >= p.num_batches) break;
        const int64_t row0 = (int64_t)batch * TM;
        __syncthreads();
        if (tid < TM) {
            const int item = p.order[row0 + tid];
            slot_item[tid] = item;
            slot_frames[tid] = item >= 0 ? p.frames[item] : 0;
            slot_row[tid] = item >= 0 ? p.item_row[item] : 0;
            slot_hrow[tid] = item >= 0 ? p.hist_row[item] : 0;
        }
        __syncthreads();
        const int batch_frames = slot_frames[0];        // longest first

        // ---- t = 0: P_0[s] = obs[0][s] + initial[s]   (viterbi.cpp:72-76) ----
        wait_for_frame(0);
        {
            const int c_end = (jt == p.NJT - 1) ? LDP : min(j0 + TN, LDP);
            for (int r = warp; r < TM; r += kThreads / kWarp) {
                const int item = slot_item[r];
                float *dst0 = p.score0 + (row0 + r) * LDP;
                float *dst1 = dst0 + p.score_stride;
                float *dst2 = dst1 + p.score_stride;
                const int64_t obs0 = slot_row[r] * S;
                for (int c = j0 + lane; c < c_end; c += kWarp) {
                    if (item >= 0 && c < S) {
                        const float v = convert_observation<kFlags>(load_observation(p.observation, obs0 + c, p.obs_dtype)) + p.initial[c];
                        dst0[c] = v;
                        if (kHist) p.history[slot_hrow[r] * S + c] = v;
                    } else {
                        dst0[c] = 0.f;                   // padding stays finite, in all buffers
                        dst1[c] = 0.f;
                        dst2[c] = 0.f;
                    }
                }
            }
        }
        if constexpr (kBanded && kHist) {
            // block maxima of this CTA's columns of P_0 (read back through L2 once per batch)
            __syncthreads();
            for (int idx = tid; idx < TM * (TN / 16); idx += kThreads) {
                const int r = idx / (TN / 16), blk = idx % (TN / 16);
                const int col = j0 + blk * 16;
                if (col >= S || slot_item[r] < 0) continue;
                const float *row = p.score0 + (row0 + r) * LDP + col;
                float m = -INFINITY;
                for (int e = 0; e < 16 && col + e < S; ++e) m = fmaxf(m, __ldcg(row + e));
                p.block_max[(row0 + r) * (LDP / 16) + col / 16] = m;
            }
        }
        barriers += p.NJT;
        group_arrive(counter);
        group_wait(counter, barriers);

        // Everything step `t` needs that does not depend on the other CTAs: its emission tile
        // (raw; converted when consumed) and the transition rows of pipeline stages 0 and 1.
        // Their completion is covered by the full-barrier arrival of the step's first stage
        // (cp.async.mbarrier.arrive tracks every earlier cp.async of the thread).
        auto prefetch_step = [&](int t) {
            wait_for_frame(t);
            if (obs_vec) {
                for (int idx = tid; idx < TM * (TN / 4); idx += kThreads) {
                    const int r = idx / (TN / 4), c4 = idx % (TN / 4);
                    const int item = slot_item[r];
                    if (item >= 0 && t < slot_frames[r] && j0 + c4 * 4 < S)
                        cp_async16(tile + r * TN + c4 * 4,
                                   p.observation + (slot_row[r] + t) * S + j0 + c4 * 4);
                }
            } else {
                load_emission_tile_plain<TM, TN, kThreads>(tile, p.observation, p.obs_dtype, slot_item, slot_frames,
                                                           slot_row, S, t, j0);
            }
            // (stage_counter already points at the first stage of step t: the previous step is done)
            if (nstages > 0) load_stage_A<TM>(ring + (stage_counter % NSTAGE) * kStageFloats + ld_smem, ld_A, ld_stride10);
            if (nstages > 1)
                load_stage_A<TM>(ring + ((stage_counter + 1) % NSTAGE) * kStageFloats + ld_smem, ld_A + KS, ld_stride10);
        };
        if (batch_frames > 1) prefetch_step(1);

        int buf = 0;                                    // buffer holding P_{t-1}
        PROF_DECL
        for (int t = 1; t < batch_frames; ++t) {
            // three rotating buffers: a CTA that is already publishing step t+1 never
            // writes the buffer a slower CTA of the group still re-scans for step t
            const int buf_prev = buf;
            const int buf_new = buf == 2 ? 0 : buf + 1;
            const float *P_cur = p.score0 + buf * p.score_stride;
            float *P_new = p.score0 + buf_new * p.score_stride;
            buf = buf_new;
            const float *ld_P = P_cur + (row0 + ld_r0) * LDP + ld_c4 * 4 + stage_base;

            // the emission tile and the transition rows of stages 0 / 1 were prefetched before
            // the exchange wait (prefetch_step); only the score rows are left
            {
                const unsigned slot0 = stage_counter % NSTAGE, slot1 = (stage_counter + 1) % NSTAGE;
                if (nstages > 0) {
                    load_stage_P<TM>(ring + slot0 * kStageFloats + ld_smem, ld_P, ld_stride10, ld_r0);
                    ring_full_arrive(full_bar + slot0);
                }
                if (nstages > 1) {
                    load_stage_P<TM>(ring + slot1 * kStageFloats + ld_smem, ld_P + KS, ld_stride10, ld_r0);
                    ring_full_arrive(full_bar + slot1);
                }
            }

            if constexpr (kBanded && kHist) {
                // Sources outside the window, score-history mode: only the VALUE max_s fl(P[s] + outside) is
                // needed, and fl(. + outside) is monotonic, so it equals fl(max_s P[s] + outside) -- taken from
                // the per-16-state block maxima the producers publish with the scores (window boundaries
                // are multiples of 16).  90 loads per row instead of 1440.
                constexpr int kRowsPerWarp = TM / (kThreads / kWarp);
                const int nblk = (S + 15) / 16, blk_lo = k_lo / 16, blk_hi = (k_hi + 15) / 16;
                const float *bm = p.block_max + buf_prev * p.block_stride + (row0 + warp * kRowsPerWarp) * (LDP / 16);
                float m[kRowsPerWarp];
#pragma unroll
                for (int rr = 0; rr < kRowsPerWarp; ++rr) m[rr] = -INFINITY;
                for (int b = lane; b < nblk; b += kWarp) {
                    if (b >= blk_lo && b < blk_hi) continue;
#pragma unroll
                    for (int rr = 0; rr < kRowsPerWarp; ++rr)
                        m[rr] = fmaxf(m[rr], __ldcg(bm + (int64_t)rr * (LDP / 16) + b));
                }
#pragma unroll
                for (int rr = 0; rr < kRowsPerWarp; ++rr) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) m[rr] = fmaxf(m[rr], __shfl_xor_sync(kFullMask, m[rr], off));
                    // (-inf when the window covers every source: never beats the window maximum)
                    if (lane == 0) out_v[warp * kRowsPerWarp + rr] = m[rr] + outside;
                }
                __syncthreads();
            } else if (kBanded) {
                // sources outside the window: first-occurrence max of P[s] + outside, per sequence
                // (the same single fp32 add the dense scan would do for these sources)
                // All rows of the warp advance together so their (independent) L2 loads overlap.
                // Per lane and row: running max over the lane's ascending float4 groups and the FIRST
                // group that strictly raised it; the index inside that group is resolved afterwards.
                constexpr int kRowsPerWarp = TM / (kThreads / kWarp);
                const float *Prow0 = P_cur + (row0 + warp * kRowsPerWarp) * LDP;
                float bv[kRowsPerWarp];
                int grp[kRowsPerWarp];
#pragma unroll
                for (int rr = 0; rr < kRowsPerWarp; ++rr) { bv[rr] = -INFINITY; grp[rr] = -1; }
#pragma unroll 2
                for (int q = lane * 4; q < S; q += kWarp * 4) {
                    if (q >= k_lo && q < k_hi) continue;                // window boundaries are multiples of 4
                    float4 pv[kRowsPerWarp];
#pragma unroll
                    for (int rr = 0; rr < kRowsPerWarp; ++rr)
                        pv[rr] = __ldcg(reinterpret_cast<const float4 *>(Prow0 + (int64_t)rr * LDP + q));
#pragma unroll
                    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
                        if (q + 3 >= S) {                                // ragged tail: padding never competes
                            if (q + 1 >= S) pv[rr].y = -INFINITY;
                            if (q + 2 >= S) pv[rr].z = -INFINITY;
                            pv[rr].w = -INFINITY;
                        }
                        const float m = fmaxf(fmaxf(pv[rr].x + outside, pv[rr].y + outside),
                                              fmaxf(pv[rr].z + outside, pv[rr].w + outside));
                        if (m > bv[rr] || grp[rr] < 0) { bv[rr] = m; grp[rr] = q; }
                    }
                }
#pragma unroll
                for (int rr = 0; rr < kRowsPerWarp; ++rr) {
                    const int r = warp * kRowsPerWarp + rr;
                    float v = bv[rr];
                    int bi = 0x7fffffff;
                    if (grp[rr] >= 0) {
                        const int g = grp[rr];
                        const float4 pv = __ldcg(reinterpret_cast<const float4 *>(Prow0 + (int64_t)rr * LDP + g));
                        bi = g + 3;
                        if (g + 2 < S && pv.z + outside == v) bi = g + 2;
                        if (g + 1 < S && pv.y + outside == v) bi = g + 1;
                        if (pv.x + outside == v) bi = g;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(kFullMask, v, off);
                        const int oi = __shfl_xor_sync(kFullMask, bi, off);
                        if (better(ov, oi, v, bi)) { v = ov; bi = oi; }
                    }
                    // (rows of finished / empty slots are scanned too: their result is never used)
                    if (lane == 0) { out_v[r] = v; out_i[r] = bi; }
                }
                __syncthreads();
            }

            float cur[M][N];
            int bestc[M][N];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c) { cur[i][c] = -INFINITY; bestc[i][c] = 0; }

            for (int s = 0; s < nstages; ++s, ++stage_counter) {
                const unsigned slot = stage_counter % NSTAGE;
                // every thread's copies for this stage have landed
                ring_full_wait(full_bar + slot, (stage_counter / NSTAGE) & 1);
                if (s == 0) { PROF_MARK(0) }
                const bool refill = s + 2 < nstages;
                const unsigned rslot = (stage_counter + 2) % NSTAGE;          // the previous stage's slot
                float *refill_dst = ring + rslot * kStageFloats + ld_smem;
                const float *refill_A = ld_A + (s + 2) * KS, *refill_P = ld_P + (s + 2) * KS;

                const float *Ps = ring + slot * kStageFloats + r_base * LDS_;
                const float *As = ring + slot * kStageFloats + (TM + c_base) * LDS_;
#pragma unroll
                for (int ch = 0; ch < KS / KC; ++ch) {
                    if (refill && ch == 1) {
                        // Refill the slot the previous stage used, one chunk into this stage: by now every
                        // warp has (almost surely) released it, so the wait does not couple the warps.
                        // (s == 0: the slot's last readers belong to the previous step, behind CTA barriers)
                        if (s > 0) {
                            ring_empty_wait(empty_bar + rslot, (empty_parity >> rslot) & 1u);
                            empty_parity ^= 1u << rslot;
                        }
                        load_stage<TM>(refill_dst, refill_A, refill_P, ld_stride10, ld_r0);
                        ring_full_arrive(full_bar + rslot);
                    }
                    const int chunk = (s_lo + s) * (KS / KC) + ch;
                    if (kBanded && (chunk < c_lo || chunk >= c_hi)) continue;     // outside the window (CTA-uniform)
                    float acc[M][N];
#pragma unroll
                    for (int k = 0; k < KC; k += 4) {
                        ulonglong2 pv[M], av[N];
#pragma unroll
                        for (int i = 0; i < M; ++i)
                            pv[i] = *reinterpret_cast<const ulonglong2 *>(Ps + i * LANES_B * LDS_ + ch * KC + k);
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            av[c] = *reinterpret_cast<const ulonglong2 *>(As + c * LANES_J * LDS_ + ch * KC + k);
#pragma unroll
                        for (int i = 0; i < M; ++i)
#pragma unroll
                            for (int c = 0; c < N; ++c) {
                                const u64 v0 = add2(pv[i].x, av[c].x);
                                const u64 v1 = add2(pv[i].y, av[c].y);
                                // first group of the chunk reads `cur`, writes a fresh `acc`
                                const float a0 = (k == 0) ? cur[i][c] : acc[i][c];
                                acc[i][c] = fmaxf(fmaxf(fmaxf(fmaxf(a0, lo(v0)), hi(v0)), lo(v1)), hi(v1));
                            }
                    }
#pragma unroll
                    for (int i = 0; i < M; ++i)
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            if (!kHist && acc[i][c] > cur[i][c]) bestc[i][c] = chunk;   // first chunk reaching the max
                            cur[i][c] = acc[i][c];
                        }
                }
                // this warp is done with the slot; signalled only when a refill (during the next stage,
                // for stage s + 3) is going to wait for it, so every barrier phase has its waiters
                if (s + 3 < nstages) {
                    __syncwarp();
                    if (lane == 0) ring_empty_arrive(empty_bar + slot);
                }
            }
            // The emission tile was copied with the step's first stage, whose full-barrier every thread
            // has passed.  Without any stage (banded, empty window) or with plain stores (unaligned
            // observations) a CTA barrier makes the other threads' tile writes visible instead.
            if (nstages == 0 || !obs_vec) {
                cp_async_wait<0>();
                __syncthreads();
            }
            PROF_MARK(1)

            // ---- new scores: obs[t][j] + max  (viterbi.cpp:102); keep max and winning chunk ----
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const int r = r_base + i * LANES_B;
                if (t < slot_frames[r]) {
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        const int col = c_base + c * LANES_J;
                        if (j0 + col < S) {
                            const float emit = convert_observation<kFlags>(tile[r * TN + col]);
                            tile[r * TN + col] = emit + (kBanded ? fmaxf(cur[i][c], out_v[r]) : cur[i][c]);
                            if (!kHist) {
                                best_tile[r * TN + col] = cur[i][c];
                                bp_tile[r * TN + col] = (uint16_t)bestc[i][c];
                            }
                        }
                    }
                }
            }
            PROF_MARK(5)
            __syncthreads();
            PROF_MARK(6)

            // ---- publish the new scores (finished rows are carried over), then arrive ----
            for (int idx = tid; idx < TM * (TN / 4); idx += kThreads) {       // (whole warps enter or leave together)
                const int r = idx / (TN / 4), c4 = idx % (TN / 4);
                const int col = j0 + c4 * 4;
                const bool valid = col < S && slot_item[r] >= 0;              // padding never changes
                float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                if (valid) {
                    const int64_t off = (row0 + r) * LDP + col;
                    const bool live = t < slot_frames[r];
                    // score-history mode: P_t of live rows also goes to the (B,T,S) history (streaming store)
                    float *hist = (kHist && live) ? p.history + (slot_hrow[r] + t) * S + col : nullptr;
                    if (col + 3 < S) {
                        v = live ? *reinterpret_cast<const float4 *>(tile + r * TN + c4 * 4)
                                 : __ldcg(reinterpret_cast<const float4 *>(P_cur + off));
                        *reinterpret_cast<float4 *>(P_new + off) = v;
                        if (kHist && live) {
                            if (hist_vec) __stcs(reinterpret_cast<float4 *>(hist), v);
                            else { hist[0] = v.x; hist[1] = v.y; hist[2] = v.z; hist[3] = v.w; }
                        }
                    } else {
                        float e[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (col + k < S) {
                                e[k] = live ? tile[r * TN + c4 * 4 + k] : __ldcg(P_cur + off + k);
                                P_new[off + k] = e[k];
                                if (kHist && live) hist[k] = e[k];
                            }
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
                if constexpr (kBanded && kHist) {
                    // max of every 16-state block of the new row, for the next step's outside-window maximum:
                    // the 4 lanes holding a block are adjacent (idx % 4 == c4 % 4); states >= S hold -inf
                    float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                    m = fmaxf(m, __shfl_xor_sync(kFullMask, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(kFullMask, m, 2));
                    if (valid && (c4 & 3) == 0)
                        p.block_max[buf_new * p.block_stride + (row0 + r) * (LDP / 16) + col / 16] = m;
                }
            }
            barriers += p.NJT;
            PROF_MARK(7)
            group_arrive(counter);
            PROF_MARK(2)
            if (t + 1 < batch_frames) prefetch_step(t + 1);   // `tile` is free: the publish above was its last reader

            // ---- exact argmax while the rest of the group catches up: lowest s in the
            //      winning chunk with P[s] + A[j][s] == max.  8 lanes share one output
            //      (a float4 of the chunk each), so every load is a coalesced 128-byte row.
            if constexpr (!kHist) {
            if (j0 + TN <= S)
                rescan_tile<TM, true, kBanded>(best_tile, bp_tile, slot_frames, P_cur + row0 * LDP,
                                               p.transition + (int64_t)j0 * LDP, LDP, S - j0, S, t, out_v, out_i);
            else
                rescan_tile<TM, false, kBanded>(best_tile, bp_tile, slot_frames, P_cur + row0 * LDP,
                                                p.transition + (int64_t)j0 * LDP, LDP, S - j0, S, t, out_v, out_i);
            __syncthreads();

            // ---- backpointers: coalesced, vectorised uint16 stores ----
            if (bp_vec) {
                for (int idx = tid; idx < TM * (TN / 8); idx += kThreads) {
                    const int r = idx / (TN / 8), c8 = idx % (TN / 8);
                    if (t < slot_frames[r] && j0 + c8 * 8 < S)
                        *reinterpret_cast<uint4 *>(p.backptr + (slot_hrow[r] + t) * S + j0 + c8 * 8) =
                            *reinterpret_cast<const uint4 *>(bp_tile + r * TN + c8 * 8);
                }
            } else {
                for (int idx = tid; idx < TM * TN; idx += kThreads) {
                    const int r = idx / TN, c = idx % TN;
                    if (t < slot_frames[r] && j0 + c < S)
                        p.backptr[(slot_hrow[r] + t) * S + j0 + c] = bp_tile[idx];
                }
            }
            }   // !kHist
            PROF_MARK(3)
            group_wait(counter, barriers);
            PROF_MARK(4)
        }
        PROF_PRINT(max(batch_frames - 1, 1))

        // ---- final scores out (viterbi.cpp:111-113) + first-occurrence argmax (:218) ----
        const float *P_fin = p.score0 + buf * p.score_stride;
        for (int r = jt * (kThreads / kWarp) + warp; r < TM; r += p.NJT * (kThreads / kWarp)) {
            const int item = slot_item[r];
            if (item < 0) continue;
            const float *row = P_fin + (row0 + r) * LDP;
            float best = -INFINITY;
            int arg = 0x7fffffff;
            for (int s = lane; s < S; s += kWarp) {
                const float v = __ldcg(row + s);
                if (p.posterior) p.posterior[(int64_t)item * S + s] = v;
                if (better(v, s, best, arg)) { best = v; arg = s; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(kFullMask, best, off);
                const int oi = __shfl_xor_sync(kFullMask, arg, off);
                if (better(ov, oi, best, arg)) { best = ov; arg = oi; }
            }
            if (lane == 0) p.final_state[item] = (arg == 0x7fffffff) ? 0 : arg;
        }
        // the next batch uses other score rows; no barrier needed here
    }
}

// (S,S) -> (SJ, LDP) with -inf padding: padded sources never win, padded
// destinations are computed and dropped
__global__ void pad_transition_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                      int S, int SJ, int LDP) {
    const int64_t n = (int64_t)SJ * LDP;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i / LDP), s = (int)(i % LDP);
        dst[i] = (j < S && s < S) ? src[(int64_t)j * S + s] : -INFINITY;
    }
}

// Banded mode, once per call: for every destination tile the range of sources whose
// transition entry differs (bitwise) from the corner value A[0][S-1].  band[jt] = S - first,
// band[NJT + jt] = last + 1 (both 0 for a tile that is constant), accumulated with atomicMax
// on zero-initialised words.  Whatever the matrix looks like, entries outside the recorded
// range equal `outside`, which is all the forward kernel relies on.
__global__ void analyze_band_kernel(const float *__restrict__ transition, int S, int NJT,
                                    unsigned *__restrict__ band, float *__restrict__ outside) {
    const int lane = threadIdx.x % kWarp;
    const int warps = gridDim.x * (blockDim.x / kWarp);
    const unsigned c = __float_as_uint(transition[S - 1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) *outside = __uint_as_float(c);
    for (int j = blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp; j < S; j += warps) {
        const float *row = transition + (int64_t)j * S;
        int first = S, last = -1;
        for (int s = lane; s < S; s += kWarp)
            if (__float_as_uint(row[s]) != c) { first = min(first, s); last = max(last, s); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            first = min(first, __shfl_xor_sync(kFullMask, first, off));
            last = max(last, __shfl_xor_sync(kFullMask, last, off));
        }
        if (lane == 0 && last >= 0) {
            atomicMax(band + j / TN, (unsigned)(S - first));
            atomicMax(band + NJT + j / TN, (unsigned)(last + 1));
        }
    }
}

}  // namespace

static int64_t tiled_ldp(int64_t S) { return padded_sources(S); }
static int64_t tiled_sj(int64_t S) { return padded_dests(S); }

#if !TORBI_TILED_HIST
cudaError_t launch_pad_transition(const ForwardArgs &a, cudaStream_t stream) {
    pad_transition_kernel<<<a.sm_count * 4, 256, 0, stream>>>(
        a.transition, a.ws.transition_pad, (int)a.S, (int)padded_dests(a.S), (int)padded_sources(a.S));
    return cudaGetLastError();
}

bool forward_tiled_supported(const ForwardArgs &a) {
    const int64_t njt = tiled_sj(a.S) / TN;
    return njt <= a.sm_count && a.ws.transition_pad != nullptr;
}
#endif

template <int M, uint32_t kFlags, bool kBanded>
static cudaError_t launch_tiled_t(const ForwardArgs &a, cudaStream_t stream) {
    constexpr int TM = tile_rows(M);
    TiledParams p{};
    p.observation = a.observation;
    p.obs_dtype = a.obs_dtype;
    p.transition = a.ws.transition_pad;
    p.initial = a.initial;
    p.frames = a.ws.frames;
    p.item_row = a.ws.item_row;
    p.hist_row = a.ws.hist_row;
    p.order = a.ws.order;
    p.score0 = a.ws.score[0];
    p.score_stride = a.ws.score[1] - a.ws.score[0];
    p.backptr = a.ws.backptr;
    p.history = reinterpret_cast<float *>(a.ws.backptr);
    p.block_max = a.ws.block_max;
    p.block_stride = pad_batch(a.B) * (tiled_ldp(a.S) / 16);
    p.posterior = a.posterior;
    p.final_state = a.ws.final_state;
    p.sync = a.ws.sync;
    p.frames_ready = a.frames_ready;
    p.S = (int)a.S;
    p.LDP = (int)tiled_ldp(a.S);
    const int SJ = (int)tiled_sj(a.S);
    p.NJT = SJ / TN;
    p.num_batches = (int)((a.B + TM - 1) / TM);
    p.NG = std::min(a.sm_count / p.NJT, p.num_batches);
    if ((p.NG + 1) * 32 > (int)kBandWordOffset) return cudaErrorInvalidValue;

    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    if (kBanded) {
        // the analysis words live in the (zeroed) tail of the sync area
        unsigned *band = a.ws.sync + kBandWordOffset;
        float *outside = reinterpret_cast<float *>(band + 2 * p.NJT);
        if (2 * p.NJT + 1 > (int)(kSyncWords - kBandWordOffset)) return cudaErrorInvalidValue;
        analyze_band_kernel<<<a.sm_count, 256, 0, stream>>>(a.transition, p.S, p.NJT, band, outside);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        p.band = band;
        p.outside = outside;
    }

    auto kernel = forward_tiled_kernel<M, kFlags, kBanded>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(M));
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&p};
    // cooperative launch: all CTAs of a group must be co-resident for the
    // per-step counter barrier; the launch fails instead of deadlocking
    return cudaLaunchCooperativeKernel((void *)kernel, dim3(p.NG * p.NJT), dim3(kThreads), args,
                                       smem_bytes(M), stream);
}

template <int M>
static cudaError_t launch_tiled_m(const ForwardArgs &a, cudaStream_t stream) {
    const bool banded = a.flags & TORBI_B200_BANDED;
    switch (a.flags & 3u) {
        case 0: return banded ? launch_tiled_t<M, 0, true>(a, stream) : launch_tiled_t<M, 0, false>(a, stream);
        case 1: return banded ? launch_tiled_t<M, 1, true>(a, stream) : launch_tiled_t<M, 1, false>(a, stream);
        case 2: return banded ? launch_tiled_t<M, 2, true>(a, stream) : launch_tiled_t<M, 2, false>(a, stream);
        default: return banded ? launch_tiled_t<M, 3, true>(a, stream) : launch_tiled_t<M, 3, false>(a, stream);
    }
}

#if !TORBI_TILED_HIST
// Relative cost of one timestep of a TM-row batch (measured on the B200, S = 1440).
static double step_cost(int tm) { return tm == 64 ? 1.0 : tm == 32 ? 0.61 : 0.53; }

// Batch height for this call: simulate longest-first list scheduling of the
// length-sorted batches over the CTA groups for TM = 64 / 32 / 16 and take the
// smallest makespan.  `frames_sorted` is descending.
int tiled_choose_rows(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count,
                      double *seconds) {
    const int njt = (int)(tiled_sj(S) / TN);
    const int groups_max = std::max(1, sm_count / njt);
    int best_tm = 64;
    double best = 1e300;
    for (int tm : {64, 32, 16}) {
        const int batches = (int)((B + tm - 1) / tm);
        const int groups = std::min(groups_max, batches);
        std::vector<double> load(groups, 0.0);
        for (int b = 0; b < batches; ++b) {     // batches arrive longest first: greedy = LPT
            auto it = std::min_element(load.begin(), load.end());
            *it += frames_sorted[(int64_t)b * tm] * step_cost(tm);
        }
        const double makespan = *std::max_element(load.begin(), load.end());
        if (makespan < best * 0.97) { best = makespan; best_tm = tm; }   // prefer the taller tile on near-ties
    }
    // one step of a 64-row batch takes 69 us at S = 1440 in score-history mode (42 / 37 us for 32 / 16 rows;
    // backpointer mode 86 / 49 / 40 us) and scales with the K loop
    if (seconds) *seconds = best * 69.0e-6 * (double)tiled_ldp(S) / 1440.0;
    return best_tm;
}

#endif

#if TORBI_TILED_HIST
cudaError_t launch_forward_tiled_hist(const ForwardArgs &a, cudaStream_t stream) {
#else
cudaError_t launch_forward_tiled(const ForwardArgs &a, cudaStream_t stream) {
#endif
    switch (a.tile_rows) {
        case 16: return launch_tiled_m<1>(a, stream);
        case 32: return launch_tiled_m<2>(a, stream);
        default: return launch_tiled_m<4>(a, stream);
    }
}

}  // namespace torbi_b200
