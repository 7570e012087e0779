// Split-K small-tile forward kernel (score-history mode only): the decomposition for
// batches of a few dozen to a few hundred sequences -- the per-GPU share of a batch
// sharded over the GPUs of a box, the long tail of a ragged batch, one batch of files.
//
// forward_tiled.cu needs 64 sequences per group of 18 CTAs to fill its 64 x 80 tiles,
// i.e. 512 sequences to fill the chip; with 64 sequences it runs on 18 (64-row tile) or
// 72 (16-row tiles, at a poor math / shared-memory-load ratio) of the 148 SMs.  Here a
// CTA owns a 16-sequence x 40-state tile, so a batch of 16 spreads over ceil(S / 40)
// CTAs (36 at S = 1440) and 64 sequences fill 144 SMs.  The register tile stays 4 x 5 per
// thread (same 9 LDS.128 per 80 math instructions as the big kernel) because ONE warp
// covers the whole 16 x 40 tile (lanes 4 x 8) and the CTA's 8 warps split the SOURCE
// states of every pipeline stage between them (split-K); their partial maxima meet in
// shared memory once per step.  max is associative and commutative, every candidate is
// the single fp32 add P[s] + A[j][s], so the scores are bit-identical to any other scan
// order; the path comes from backtrace_recompute_kernel (backtrace.cu).
//
// Two tile shapes (template parameter):
//   * 32 sequences x 20 states, transition block RESIDENT: the CTA's 20 rows of A (115 KB at
//     S = 1440) are copied to shared memory once per launch and only the score rows stream
//     through the ring.  L2 -> SM traffic per step drops from 48 MB to 26 MB at S = 1440, which
//     is what bounds the streaming shape (measured 4.4 TB/s).  Needs S <= 1536 (shared memory).
//   * 16 sequences x 40 states, transition block STREAMED with the scores: any S <= 40 x #SM.
//
// In the resident shape every warp streams ITS eighth of the score columns through a private
// double buffer (cp.async groups + __syncwarp; no barrier couples the warps inside a step), in
// stages of 8 k-steps that run without a branch so the operand loads of the next k-step are
// scheduled under the math of the current one (worth 7 %).  The streamed shape uses a 6-stage
// mbarrier ring shared by the CTA, like forward_tiled.cu.
//
// Everything else follows forward_tiled.cu: persistent cooperative grid, groups of
// NJT = ceil(S / TN) CTAs per batch, longest-first batch queue, scores exchanged through
// rotating L2-resident buffers with one counter barrier per step, fused log-space conversion,
// fp32 score history.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "sync.cuh"

namespace torbi_b200 {

namespace {

constexpr int kThreads = 256, kWarps = kThreads / kWarp;
constexpr int M = 4, N = 5;           // register tile
constexpr int kLoaders = 240;         // loader threads: (KS / 4) float4 columns x (960 / KS) rows per pass
constexpr int kTileCells = 640;       // TM * TN of either shape

// KS = source states per pipeline stage; a stage's KS / 4 float4 k-steps are dealt round-robin to the
// 8 warps.  The resident shape stages scores only and takes twice as many sources per stage: the
// fixed cost of a stage (barrier waits, refill issue) is paid half as often.
struct Resident32x20 {                // lanes 8 (sequences) x 4 (states)
    static constexpr int TM = 32, TN = 20, LANES_B = 8, LANES_J = 4, KS = 96, NSTAGE = 6;
    static constexpr bool kResident = true;
    static constexpr int kMaxLDP = 1536;                       // 20 rows x (LDP + 4) floats must fit
};
struct Streamed16x40 {                // lanes 4 x 8
    static constexpr int TM = 16, TN = 40, LANES_B = 4, LANES_J = 8, KS = 96, NSTAGE = 6;
    static constexpr bool kResident = false;
    static constexpr int kMaxLDP = 1 << 30;
};

// Resident shape: every warp runs a PRIVATE cp.async pipeline over its own eighth of the source states
// (double-buffered stages of kWarpSources sources, rows kWarpStride floats apart: 8 rows -> 8 distinct
// bank groups), synchronised with cp.async.wait_group + __syncwarp only -- no barrier couples the warps
// inside a step.  A lane copies 16-byte column lane % 8 of rows lane / 8 + 4 q: every copy instruction
// covers four full 128-byte lines.
constexpr int kWarpSources = 32, kWarpStride = 36, kWarpStages = 2;
template <typename Shape>
constexpr int stage_floats() { return (Shape::TM + (Shape::kResident ? 0 : Shape::TN)) * (Shape::KS + 4); }
template <typename Shape>
constexpr int ring_floats() {
    return Shape::kResident ? kWarps * kWarpStages * Shape::TM * kWarpStride : Shape::NSTAGE * stage_floats<Shape>();
}
template <typename Shape>
size_t smem_bytes(int ldp) {
    return sizeof(float) * (ring_floats<Shape>() + kTileCells + kWarps * kTileCells +
                            (Shape::kResident ? Shape::TN * (ldp + 4) : 0)) +
           sizeof(int) * 6 * Shape::TM + sizeof(u64) * 2 * Shape::NSTAGE;
}

__device__ __forceinline__ u64 add2(u64 a, u64 b) {     // two fp32 adds (FADD2), IEEE rn
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
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

// per-phase cycle counters of the step loop (instrumented builds of tools/profile_phases.sh only)
#ifdef TORBI_TILED_PROFILE
#define PROF_DECL long long prof_t0 = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_MARK(i) { const long long prof_now = clock64(); prof_acc[i] += prof_now - prof_t0; prof_t0 = prof_now; }
#define PROF_PRINT(steps)                                                                              \
    if (threadIdx.x == 0 && blockIdx.x % 37 == 0)                                                      \
        printf("cta %d steps %d cycles/step: score loads + first-stage wait %lld, K loop %lld, partial maxima + new scores %lld, " \
               "publish %lld, arrive %lld, prefetch %lld, exchange wait %lld\n", blockIdx.x, (steps),             \
               prof_acc[0] / (steps), prof_acc[1] / (steps), prof_acc[2] / (steps), prof_acc[3] / (steps),      \
               prof_acc[4] / (steps), prof_acc[5] / (steps), prof_acc[6] / (steps));
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_PRINT(steps)
#endif

struct SplitParams {
    const float *observation;   // (rows,S), item b at rows item_row[b] ..; 2-byte elements when obs_dtype != 0
    int obs_dtype;              // 0 fp32, 1 fp16, 2 bf16
    const float *transition;    // padded copy (SJ, LDP), -inf outside (S,S)
    const float *initial;       // (S,)
    const int32_t *frames;      // (B,)
    const int64_t *item_row;    // (B,) first row of each item in observation
    const int64_t *hist_row;    // (B,) first row of each item in the (compact) history
    const int32_t *order;       // (Bpad,) items longest first, -1 = empty slot
    float *score0;              // 3 rotating buffers (Bpad, LDP), `score_stride` floats apart
    int64_t score_stride;
    float *history;             // (rows,S) fp32: every P_t
    float *posterior;           // (B,S) or null
    int32_t *final_state;       // (B,)
    unsigned *sync;             // per group: barrier counter (+0), current batch (+1); queue head after the groups
    const int32_t *frames_ready;
    int S, LDP, NJT, NG, num_batches;
};

// One stage: the score rows and -- streamed shape -- the transition rows.  Thread tid < 240 owns
// float4 column tid % (KS / 4) of rows tid / (KS / 4) + kRows q, kRows = 960 / KS rows per pass;
// `in_row` is false for columns past the end of the (padded) row (last stage of the resident shape).
template <typename Shape>
__device__ __forceinline__ void load_stage_A(float *stage, const float *gA, int64_t pass_stride) {
    if constexpr (!Shape::kResident) {
        constexpr int kRows = 960 / Shape::KS, LDS_ = Shape::KS + 4;
        if (threadIdx.x < kLoaders) {
            float *sA = stage + Shape::TM * LDS_;
#pragma unroll
            for (int q = 0; q < Shape::TN / kRows; ++q) cp_async16(sA + q * kRows * LDS_, gA + q * pass_stride);
        }
    }
}
template <typename Shape>
__device__ __forceinline__ void load_stage_P(float *stage, const float *gP, int64_t pass_stride, int r0, bool in_row) {
    constexpr int kRows = 960 / Shape::KS, LDS_ = Shape::KS + 4;
    if (threadIdx.x < kLoaders && in_row) {
#pragma unroll
        for (int q = 0; q < (Shape::TM + kRows - 1) / kRows; ++q)
            if (q * kRows + kRows - 1 < Shape::TM || r0 + q * kRows < Shape::TM)
                cp_async16(stage + q * kRows * LDS_, gP + q * pass_stride);
    }
}

template <typename Shape, uint32_t kFlags>
__global__ void __launch_bounds__(kThreads, 1) forward_split_kernel(const SplitParams p) {
    constexpr int TM = Shape::TM, TN = Shape::TN, LANES_B = Shape::LANES_B, LANES_J = Shape::LANES_J;
    constexpr int KS = Shape::KS, LDS_ = KS + 4, NSTAGE = Shape::NSTAGE, kSteps = KS / 4;
    constexpr int kStageFloats = stage_floats<Shape>();
    static_assert(TM == LANES_B * M && TN == LANES_J * N && TM * TN == kTileCells, "tile shape");
    static_assert(!Shape::kResident || (TM == 32 && kWarpStages == 2), "per-warp pipeline: lane / 8 + 4 q covers 32 rows");
    static_assert(kSteps % kWarps == 0 && 960 % KS == 0 && (Shape::kResident || TN % (960 / KS) == 0), "stage shape");
    extern __shared__ __align__(16) float smem[];
    float *ring = smem;
    float *tile = smem + ring_floats<Shape>();                  // [TM][TN] emission -> new score
    float *partial = tile + TM * TN;                            // [warp][TM][TN] per-warp maxima
    float *A_res = partial + kWarps * TM * TN;                  // resident shape: [TN][LDP + 4] transition block
    const int LDA = p.LDP + 4;                                  // rows 4 banks apart, like the ring's
    int64_t *slot_row = reinterpret_cast<int64_t *>(A_res + (Shape::kResident ? TN * LDA : 0));   // [TM] first row of the slot's item
    int64_t *slot_hrow = slot_row + TM;                         // [TM] its first row in the history
    int *slot_item = reinterpret_cast<int *>(slot_hrow + TM);
    int *slot_frames = slot_item + TM;
    u64 *full_bar = reinterpret_cast<u64 *>(slot_frames + TM);  // [NSTAGE] one arrival per thread: its copies landed
    u64 *empty_bar = full_bar + NSTAGE;                         // [NSTAGE] one arrival per warp: done reading
    if (threadIdx.x == 0) {
        for (int k = 0; k < NSTAGE; ++k) {
            mbar_init(full_bar + k, kThreads);
            mbar_init(empty_bar + k, kWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // ring position of the next stage to consume, kept as wrap-around counters (no modulo in the
    // stage loop): its slot, and the parity of the full-barrier phase that stage completes
    unsigned slot = 0, full_parity = 0;
    unsigned empty_parity = 0;       // bit k: parity of the next phase of empty_bar[k] to wait for
    __shared__ int ready_shared;
    int ready = p.frames_ready ? 0 : 0x7fffffff;
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
    unsigned *counter = p.sync + group * 32;
    unsigned barriers = 0;

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int lane_b = lane / LANES_J, lane_j = lane % LANES_J;      // rows lane_b + 4 i, cols lane_j + 8 c
    const int S = p.S, LDP = p.LDP;
    const int nstages = (LDP + KS - 1) / KS;           // (resident shape: the last stage may be half empty)
    const bool obs_vec = p.obs_dtype == 0 && (S % 4) == 0 && (reinterpret_cast<uintptr_t>(p.observation) & 15) == 0;
    const bool hist_vec = (S % 4) == 0;
    const int ld_c4 = tid % (KS / 4), ld_r0 = tid / (KS / 4);
    const int ld_smem = ld_r0 * LDS_ + ld_c4 * 4;
    const float *ld_A = p.transition + (int64_t)(j0 + ld_r0) * LDP + ld_c4 * 4;
    const int64_t ld_stride10 = (int64_t)(960 / KS) * LDP;     // one loader pass further down
    // stages whose loads are issued before their step's first stage is consumed
    const int ahead = min(nstages, NSTAGE - 1);
    if constexpr (Shape::kResident) {
        // the CTA's transition block, once per launch (plain loads + stores: 115 KB)
        for (int idx = tid; idx < TN * (LDP / 4); idx += kThreads) {
            const int r = idx / (LDP / 4), q = idx % (LDP / 4);
            *reinterpret_cast<float4 *>(A_res + r * LDA + q * 4) =
                __ldg(reinterpret_cast<const float4 *>(p.transition + (int64_t)(j0 + r) * LDP) + q);
        }
        __syncthreads();
    }

    unsigned *queue_head = p.sync + p.NG * 32;
    for (;;) {
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
            for (int r = warp; r < TM; r += kWarps) {
                const int item = slot_item[r];
                float *dst0 = p.score0 + (row0 + r) * LDP;
                float *dst1 = dst0 + p.score_stride;
                float *dst2 = dst1 + p.score_stride;
                const int64_t obs0 = slot_row[r] * S;
                for (int c = j0 + lane; c < c_end; c += kWarp) {
                    if (item >= 0 && c < S) {
                        const float v = convert_observation<kFlags>(load_observation(p.observation, obs0 + c, p.obs_dtype)) + p.initial[c];
                        dst0[c] = v;
                        p.history[slot_hrow[r] * S + c] = v;
                    } else {
                        dst0[c] = 0.f;                   // padding stays finite, in all buffers
                        dst1[c] = 0.f;
                        dst2[c] = 0.f;
                    }
                }
            }
        }
        barriers += p.NJT;
        group_arrive(counter);
        group_wait(counter, barriers);

        // what step t needs that does not depend on the other CTAs: its emission tile (raw) and the
        // transition rows of the stages issued ahead; completion is covered by the full barrier of
        // the step's first stage (cp.async.mbarrier.arrive tracks every earlier cp.async of the thread)
        auto prefetch_step = [&](int t) {
            wait_for_frame(t);
            if (obs_vec) {
                if (tid < TM * (TN / 4)) {
                    const int r = tid / (TN / 4), c4 = tid % (TN / 4);
                    const int item = slot_item[r];
                    if (item >= 0 && t < slot_frames[r] && j0 + c4 * 4 < S)
                        cp_async16(tile + r * TN + c4 * 4, p.observation + (slot_row[r] + t) * S + j0 + c4 * 4);
                }
            } else {
                load_emission_tile_plain<TM, TN, kThreads>(tile, p.observation, p.obs_dtype, slot_item, slot_frames,
                                                           slot_row, S, t, j0);
            }
            for (int s = 0; s < ahead; ++s)
                load_stage_A<Shape>(ring + ((slot + s) % NSTAGE) * kStageFloats + ld_smem, ld_A + s * KS, ld_stride10);
        };
        if (batch_frames > 1) prefetch_step(1);

        int buf = 0;                                    // buffer holding P_{t-1}
        PROF_DECL
        for (int t = 1; t < batch_frames; ++t) {
            const int buf_new = buf == 2 ? 0 : buf + 1;
            const float *P_cur = p.score0 + buf * p.score_stride;
            float *P_new = p.score0 + buf_new * p.score_stride;
            buf = buf_new;
            const float *ld_P = P_cur + (row0 + ld_r0) * LDP + ld_c4 * 4;

            float cur[M][N];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c) cur[i][c] = -INFINITY;

            if constexpr (Shape::kResident) {
                // ---- per-warp private pipelines over the warp's eighth of the sources ----
                constexpr int KW = kWarpSources, WLD = kWarpStride, WST = kWarpStages;
                const int L = LDP / kWarps;                       // multiple of 12: LDP is a multiple of 96
                const int wstages = (L + KW - 1) / KW;            // (the last one may be partial)
                float *wring = ring + warp * (WST * TM * WLD);
                const int ld_chunk = (lane % 8) * 4, ld_row = lane / 8;
                const float *gP = P_cur + (row0 + ld_row) * LDP + warp * L + ld_chunk;
                float *sP = wring + ld_row * WLD + ld_chunk;
                auto issue = [&](int st) {
                    if (st < wstages && st * KW + ld_chunk < L) {
                        float *dst = sP + (st & 1) * (TM * WLD);
                        const float *src = gP + st * KW;
#pragma unroll
                        for (int q = 0; q < TM / 4; ++q) cp_async16(dst + q * 4 * WLD, src + (int64_t)q * 4 * LDP);
                    }
                    cp_async_commit();                            // (empty groups keep the count uniform)
                };
                issue(0);
                const float *Aw = A_res + lane_j * LDA + warp * L;
                for (int st = 0; st < wstages; ++st) {
                    cp_async_wait<0>();                           // this lane's copies of stage st have landed ...
                    __syncwarp();                                 // ... and so have the other lanes'
                    if (st == 0) { PROF_MARK(0) }
                    issue(st + 1);             // into the buffer read one iteration ago (every lane is past it)
                    const float *Ps = wring + (st & 1) * (TM * WLD) + lane_b * WLD;
                    const float *As = Aw + st * KW;
                    const int kmax = min(KW, L - st * KW);
                    // full stages run without a branch between k-steps, so the operand loads of the next
                    // k-step are scheduled under the math of the current one; only a partial last stage
                    // pays for the (warp-uniform) bound check
                    auto k_steps = [&](auto partial_stage) {
#pragma unroll
                        for (int k = 0; k < KW; k += 4) {
                            if constexpr (decltype(partial_stage)::value) {
                                if (k >= kmax) break;
                            }
                            ulonglong2 pv[M], av[N];
#pragma unroll
                            for (int i = 0; i < M; ++i)
                                pv[i] = *reinterpret_cast<const ulonglong2 *>(Ps + i * LANES_B * WLD + k);
#pragma unroll
                            for (int c = 0; c < N; ++c)
                                av[c] = *reinterpret_cast<const ulonglong2 *>(As + c * LANES_J * LDA + k);
#pragma unroll
                            for (int i = 0; i < M; ++i)
#pragma unroll
                                for (int c = 0; c < N; ++c) {
                                    const u64 v0 = add2(pv[i].x, av[c].x);
                                    const u64 v1 = add2(pv[i].y, av[c].y);
                                    cur[i][c] = fmaxf(fmaxf(fmaxf(fmaxf(cur[i][c], lo(v0)), hi(v0)), lo(v1)), hi(v1));
                                }
                        }
                    };
                    if (kmax == KW) k_steps(std::false_type{});
                    else k_steps(std::true_type{});
                }
                cp_async_wait<0>();            // the emission tile's copies as well
            } else {
                // the score rows of the stages issued ahead (their transition rows went out before the wait)
                for (int s = 0; s < ahead; ++s) {
                    const unsigned aslot = (slot + s) % NSTAGE;
                    load_stage_P<Shape>(ring + aslot * kStageFloats + ld_smem, ld_P + s * KS, ld_stride10, ld_r0,
                                        KS == kStageSources || s * KS + ld_c4 * 4 < LDP);
                    ring_full_arrive(full_bar + aslot);
                }
                for (int s = 0; s < nstages; ++s) {
                    ring_full_wait(full_bar + slot, full_parity);                       // everyone's copies have landed
                    if (s == 0) { PROF_MARK(0) }
                    // refill the slot the previous stage used with stage s + NSTAGE - 1
                    if (s + NSTAGE - 1 < nstages) {
                        const unsigned rslot = slot == 0 ? NSTAGE - 1 : slot - 1;
                        if (s > 0) {    // (s == 0: its last readers belong to the previous step, behind CTA barriers)
                            ring_empty_wait(empty_bar + rslot, (empty_parity >> rslot) & 1u);
                            empty_parity ^= 1u << rslot;
                        }
                        float *dst = ring + rslot * kStageFloats + ld_smem;
                        load_stage_A<Shape>(dst, ld_A + (s + NSTAGE - 1) * KS, ld_stride10);
                        load_stage_P<Shape>(dst, ld_P + (s + NSTAGE - 1) * KS, ld_stride10, ld_r0,
                                            KS == kStageSources || (s + NSTAGE - 1) * KS + ld_c4 * 4 < LDP);
                        ring_full_arrive(full_bar + rslot);
                    }
                    const float *Ps = ring + slot * kStageFloats + lane_b * LDS_;
                    const float *As = Shape::kResident ? A_res + lane_j * LDA + s * KS
                                                       : ring + slot * kStageFloats + (TM + lane_j) * LDS_;
                    const int lda = Shape::kResident ? LDA : LDS_;
    #pragma unroll
                    for (int kk = 0; kk < kSteps / kWarps; ++kk) {
                        const int k = (warp + kk * kWarps) * 4;          // this warp's share of the stage's sources
                        // stages wider than the 96-state row padding: the second half of the last one may lie past
                        // the end of the row (CTA-uniform condition; k >= 96 exactly for the upper half of kk)
                        if (KS > kStageSources && kk >= kSteps / kWarps / 2 && s == nstages - 1 && LDP % KS != 0) continue;
                        ulonglong2 pv[M], av[N];
    #pragma unroll
                        for (int i = 0; i < M; ++i)
                            pv[i] = *reinterpret_cast<const ulonglong2 *>(Ps + i * LANES_B * LDS_ + k);
    #pragma unroll
                        for (int c = 0; c < N; ++c)
                            av[c] = *reinterpret_cast<const ulonglong2 *>(As + c * LANES_J * lda + k);
    #pragma unroll
                        for (int i = 0; i < M; ++i)
    #pragma unroll
                            for (int c = 0; c < N; ++c) {
                                const u64 v0 = add2(pv[i].x, av[c].x);
                                const u64 v1 = add2(pv[i].y, av[c].y);
                                cur[i][c] = fmaxf(fmaxf(fmaxf(fmaxf(cur[i][c], lo(v0)), hi(v0)), lo(v1)), hi(v1));
                            }
                    }
                    // signalled only when a refill (during the next stage) is going to wait for it
                    if (s + NSTAGE < nstages) {
                        __syncwarp();
                        if (lane == 0) ring_empty_arrive(empty_bar + slot);
                    }
                    if (++slot == NSTAGE) { slot = 0; full_parity ^= 1u; }
                }
            }
            if (!obs_vec) __syncthreads();              // plain-store emission tile: make it visible
            PROF_MARK(1)

            // ---- the 8 per-warp maxima meet; new scores: obs[t][j] + max  (viterbi.cpp:102) ----
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c)
                    partial[warp * (TM * TN) + (lane_b + i * LANES_B) * TN + lane_j + c * LANES_J] = cur[i][c];
            __syncthreads();
            for (int o = tid; o < TM * TN; o += kThreads) {
                const int r = o / TN, col = o % TN;
                if (t < slot_frames[r] && j0 + col < S) {
                    float m = partial[o];
#pragma unroll
                    for (int w = 1; w < kWarps; ++w) m = fmaxf(m, partial[w * (TM * TN) + o]);
                    tile[o] = convert_observation<kFlags>(tile[o]) + m;
                }
            }
            __syncthreads();
            PROF_MARK(2)

            // ---- publish the new scores (finished rows are carried over) + history, then arrive ----
            if (tid < TM * (TN / 4)) {
                const int r = tid / (TN / 4), c4 = tid % (TN / 4);
                const int col = j0 + c4 * 4;
                if (col < S && slot_item[r] >= 0) {                       // padding never changes
                    const int64_t off = (row0 + r) * LDP + col;
                    const bool live = t < slot_frames[r];
                    float *hist = live ? p.history + (slot_hrow[r] + t) * S + col : nullptr;
                    if (col + 3 < S) {
                        const float4 v = live ? *reinterpret_cast<const float4 *>(tile + r * TN + c4 * 4)
                                              : __ldcg(reinterpret_cast<const float4 *>(P_cur + off));
                        *reinterpret_cast<float4 *>(P_new + off) = v;
                        if (live) {
                            if (hist_vec) __stcs(reinterpret_cast<float4 *>(hist), v);
                            else { hist[0] = v.x; hist[1] = v.y; hist[2] = v.z; hist[3] = v.w; }
                        }
                    } else {
                        for (int e = 0; col + e < S; ++e) {
                            const float v = live ? tile[r * TN + c4 * 4 + e] : __ldcg(P_cur + off + e);
                            P_new[off + e] = v;
                            if (live) hist[e] = v;
                        }
                    }
                }
            }
            barriers += p.NJT;
            PROF_MARK(3)
            group_arrive(counter);
            PROF_MARK(4)
            if (t + 1 < batch_frames) prefetch_step(t + 1);   // `tile` is free: the publish above was its last reader
            PROF_MARK(5)
            group_wait(counter, barriers);
            PROF_MARK(6)
        }
        PROF_PRINT(max(batch_frames - 1, 1))

        // ---- final scores out (viterbi.cpp:111-113) + first-occurrence argmax (:218) ----
        const float *P_fin = p.score0 + buf * p.score_stride;
        for (int r = jt * kWarps + warp; r < TM; r += p.NJT * kWarps) {
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
    }
}

}  // namespace

static bool resident_shape(int64_t S, int sm_count) {
    return padded_sources(S) <= Resident32x20::kMaxLDP && (S + Resident32x20::TN - 1) / Resident32x20::TN <= sm_count;
}

bool forward_split_supported(const ForwardArgs &a) {
    const bool streamed = (a.S + Streamed16x40::TN - 1) / Streamed16x40::TN <= a.sm_count;
    return (resident_shape(a.S, a.sm_count) || streamed) && a.ws.transition_pad != nullptr;
}

// Estimated seconds: longest-first list scheduling of the batches over the CTA groups; one step
// of a batch measured on the B200 at S = 1440 (scales with the K loop).
double split_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count) {
    const bool resident = resident_shape(S, sm_count);
    const int tm = resident ? Resident32x20::TM : Streamed16x40::TM;
    const int tn = resident ? Resident32x20::TN : Streamed16x40::TN;
    const double step_seconds = resident ? 12.7e-6 : 14.1e-6;      // per 32-row / 16-row batch
    const int njt = (int)((S + tn - 1) / tn);
    const int batches = (int)((B + tm - 1) / tm);
    const int groups = std::max(1, std::min(sm_count / std::max(njt, 1), batches));
    std::vector<double> load(groups, 0.0);
    for (int b = 0; b < batches; ++b) {
        auto it = std::min_element(load.begin(), load.end());
        *it += frames_sorted[(int64_t)b * tm];
    }
    return *std::max_element(load.begin(), load.end()) * step_seconds * (double)padded_sources(S) / 1440.0;
}

template <typename Shape>
static cudaError_t launch_split_t(const ForwardArgs &a, cudaStream_t stream) {
    SplitParams p{};
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
    p.history = reinterpret_cast<float *>(a.ws.backptr);
    p.posterior = a.posterior;
    p.final_state = a.ws.final_state;
    p.sync = a.ws.sync;
    p.frames_ready = a.frames_ready;
    p.S = (int)a.S;
    p.LDP = (int)padded_sources(a.S);
    p.NJT = (int)((a.S + Shape::TN - 1) / Shape::TN);
    p.num_batches = (int)((a.B + Shape::TM - 1) / Shape::TM);
    p.NG = std::min(a.sm_count / p.NJT, p.num_batches);
    if ((p.NG + 1) * 32 > (int)kBandWordOffset) return cudaErrorInvalidValue;

    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    const size_t smem = smem_bytes<Shape>(p.LDP);
    void *args[] = {(void *)&p};
    auto launch = [&](auto kernel) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        return cudaLaunchCooperativeKernel((void *)kernel, dim3(p.NG * p.NJT), dim3(kThreads), args, smem, stream);
    };
    switch (a.flags & 3u) {
        case 0: return launch(forward_split_kernel<Shape, 0>);
        case 1: return launch(forward_split_kernel<Shape, 1>);
        case 2: return launch(forward_split_kernel<Shape, 2>);
        default: return launch(forward_split_kernel<Shape, 3>);
    }
}

cudaError_t launch_forward_split(const ForwardArgs &a, cudaStream_t stream) {
    // TORBI_B200_SPLIT_STREAMED: force the streamed shape where the resident one would be used (tests)
    if (resident_shape(a.S, a.sm_count) && !(a.flags & TORBI_B200_SPLIT_STREAMED))
        return launch_split_t<Resident32x20>(a, stream);
    return launch_split_t<Streamed16x40>(a, stream);
}

}  // namespace torbi_b200
