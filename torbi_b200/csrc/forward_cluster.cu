// Cluster split-K forward kernel (score-history mode): the decomposition for 32 ... ~400 sequences
// per GPU -- the per-GPU share of the headline batch sharded over 8 GPUs (64 sequences), one batch
// of files, the tail of a ragged batch.
//
// forward_split.cu gives a CTA a 32-sequence x 20-state tile and lets its 8 warps split the source
// states; per timestep it pays a group barrier through L2 plus the first score loads after it
// (~5 k of its ~25 k cycles per step) with nothing to overlap them, because 64 sequences x 1440
// states are exactly one 640-cell tile per SM.  Here the SAME tile belongs to a thread-block CLUSTER
// of two CTAs (two SMs of one TPC):
//   * split-K 16 ways: CTA rank r of the pair owns the source states [r LDP/2, (r+1) LDP/2), its 8
//     warps an eighth of that each; the pair's half of the transition block (20 rows x LDP/2, 58 KB
//     at S = 1440) stays resident in shared memory for the whole launch;
//   * so a pair has room for TWO tiles, i.e. it serves every batch of 32 sequences of its group in
//     turn: step t of batch 0, step t of batch 1, ..., step t+1 of batch 0 ...  While batch b's new
//     scores travel through L2 to the other pairs (publish, counter, their loads), the pair is busy
//     with the other batches: the exchange latency that bounded forward_split is off the critical
//     path as soon as a group holds two batches (64 sequences);
//   * the 16 partial maxima of a cell meet in two hops: the 8 warps of a CTA through its own shared
//     memory, the two CTAs through DISTRIBUTED shared memory -- each CTA pushes the half of its
//     reduced tile that its peer finalises straight into the peer's inbox (st.shared::cluster) and
//     arrives on the peer's mbarrier (mbarrier.arrive.release.cluster); no cluster-wide barrier;
//   * every warp streams its source columns of the score rows through a private cp.async ring that
//     runs ACROSS batch-steps (the loads of the next batch's first stage are issued during the current
//     batch's last stage), gated by the per-batch publish counters.
// max is associative and commutative and every candidate is the single fp32 add P[s] + A[j][s], so
// the scores are bit-identical to any other scan order (viterbi.cpp:91-104); the path comes from
// backtrace_recompute_kernel (backtrace.cu) like for the other score-history kernels.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "sync.cuh"

namespace torbi_b200 {

namespace {

constexpr int kThreads = 256, kWarps = kThreads / kWarp;
constexpr int M = 4, N = 5;                       // register tile
constexpr int TM = 32, TN = 20, LANES_B = 8, LANES_J = 4;
constexpr int kStageSteps = 8;                    // k-steps (float4 groups of sources) per ring stage
constexpr int kRowStride = kStageSteps * 4 + 4;   // 36 floats: the 8 rows of an LDS.128 fall in 8 distinct bank groups
constexpr int kSlots = 2;                         // ring slots per warp: one stage in flight beside the one being read
constexpr int kSlotFloats = TM * kRowStride;
constexpr int kQuads = M * N / 4;                 // float4 groups of a thread's 20 cells
constexpr int kReduceThreads = kQuads * kWarp;    // 160: one per (quad, lane) slot of the tile
constexpr int kMaxBatches = 24;                   // batches a group interleaves at a time
constexpr int kSyncStride = 8;                    // words between counters: one 32-byte sector each

struct ClusterParams {
    const float *observation;   // (rows,S), item b at rows item_row[b] ..; 2-byte elements when obs_dtype != 0
    int obs_dtype;
    const float *transition;    // padded copy (SJ, LDP), -inf outside (S,S)
    const float *initial;       // (S,)
    const int32_t *frames;      // (B,)
    const int64_t *item_row;    // (B,) first row of each item in observation
    const int64_t *hist_row;    // (B,) first row of each item in the (compact) history
    const int32_t *order;       // (Bpad,) items longest first, -1 = empty slot
    float *score0;              // 2 alternating buffers (Bpad, LDP), `score_stride` floats apart
    int64_t score_stride;
    float *history;             // (rows,S) fp32: every P_t
    float *posterior;           // (B,S) or null
    int32_t *final_state;       // (B,)
    unsigned *sync;             // counters, kSyncStride words apart: [1 + batch] publish counts, then one per group
    const int32_t *frames_ready;
    int S, LDP, NJT, NG, num_batches;
    int nk[kWarps];             // k-steps of each warp within its CTA's half of the sources
};

size_t cluster_smem_bytes(int ldp) {
    return sizeof(float) * (kWarps * kSlots * kSlotFloats + TN * (ldp / 2 + 4)) +
           sizeof(float4) * (2 * kWarps * kReduceThreads + 2 * kReduceThreads) +
           sizeof(int) * (6 * kMaxBatches * TM + kMaxBatches + 1) + sizeof(u64) * 6 + 16;
}

__device__ __forceinline__ u64 add2(u64 a, u64 b) {     // two fp32 adds (FADD2), IEEE rn
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    // .cg: L2 only -- the score buffers are rewritten every step by other SMs
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> the same offset in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ unsigned map_to_rank(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(unsigned addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// one arrival on a peer CTA's mbarrier; release at cluster scope orders this thread's earlier DSMEM stores
__device__ __forceinline__ void mbar_arrive_remote(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(u64 *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// The batch-steps of a group in execution order: rounds tau = 1, 2, ...; in round tau every batch
// whose longest item has more than tau frames advances one step, in batch order.  Batches are sorted
// longest first, so the running ones are always the first `n`.  Every warp of every CTA of the group
// walks the same sequence, once for its loads (ahead) and once for its math.
struct Cursor {
    int tau, b, n;
};
__device__ __forceinline__ void cursor_first(Cursor &c, const int *len, int batches) {
    c.tau = 1;
    c.b = 0;
    c.n = batches;
    while (c.n > 0 && len[c.n - 1] <= 1) --c.n;
}
__device__ __forceinline__ void cursor_next(Cursor &c, const int *len) {
    if (++c.b >= c.n) {
        c.b = 0;
        ++c.tau;
        while (c.n > 0 && len[c.n - 1] <= c.tau) --c.n;
    }
}

__device__ __forceinline__ void group_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire(counter) < target) {}
    }
    __syncthreads();
}

#ifdef TORBI_TILED_PROFILE
#define CPROF_DECL long long prof_t0 = clock64(), prof_acc[6] = {0, 0, 0, 0, 0, 0}; long long prof_steps = 0;
#define CPROF_REDUCER_PRINT                                                                                              \
    if (lane == 0 && blockIdx.x % 37 == 0 && warp == kWarps && prof_steps > 0)                                           \
        printf("cta %d reducer warp batch-steps %lld cycles/batch-step: emission requests %lld, wait for the partials %lld, " \
               "8-way maximum %lld, DSMEM exchange + publish %lld, fence + counter %lld\n", blockIdx.x, prof_steps,         \
               prof_acc[3] / prof_steps, prof_acc[4] / prof_steps, prof_acc[5] / prof_steps, prof_acc[0] / prof_steps,     \
               prof_acc[1] / prof_steps);
#define CPROF_MARK(i) { const long long prof_now = clock64(); prof_acc[i] += prof_now - prof_t0; prof_t0 = prof_now; }
#else
#define CPROF_DECL
#define CPROF_REDUCER_PRINT
#define CPROF_MARK(i)
#endif

// kWS (warp specialised): a ninth warp does everything after the K loop -- it collects the eight warps' partial
// maxima (handed over through double-buffered shared memory and mbarriers), runs the DSMEM exchange with the
// peer CTA, adds the emissions, publishes and signals the batch's counter -- while the eight K-loop warps go
// straight on to the next batch of the group.  Without it warps 0..4 carry that tail themselves (5.6 k of
// 15.1 k cycles per batch-step, profiles/r02_ab_cluster_vs_split.txt) and the other three idle behind them.
// Measured: bit-exact, but slower still -- one warp sharing a scheduler with two K-loop warps takes ~12 k
// cycles for the tail (same file), so it is off by default (TORBI_B200_CLUSTER_WS=1 selects it).
template <uint32_t kFlags, bool kWS>
__global__ void __launch_bounds__(kWS ? kThreads + kWarp : kThreads, 1) forward_cluster_kernel(const ClusterParams p) {
    constexpr int kAllThreads = kWS ? kThreads + kWarp : kThreads;
    extern __shared__ __align__(16) float smem[];
    const int S = p.S, LDP = p.LDP;
    const int KH4 = LDP / 2;                                    // sources of this CTA's half
    const int LDA = KH4 + 4;                                    // rows 4 banks apart
    float *ring = smem;                                         // [warp][slot][TM][kRowStride]
    float *A_res = ring + kWarps * kSlots * kSlotFloats;        // [TN][LDA]
    float4 *partial = reinterpret_cast<float4 *>(A_res + TN * LDA);   // [2][warp][160] per-warp maxima
    float4 *inbox = partial + 2 * kWarps * kReduceThreads;      // [2][160] the peer's reduced half
    int64_t *s_row = reinterpret_cast<int64_t *>(inbox + 2 * kReduceThreads);   // [kMaxBatches][TM] first row of the item
    int64_t *s_hrow = s_row + kMaxBatches * TM;                          // [kMaxBatches][TM] its first row in the history
    int *s_item = reinterpret_cast<int *>(s_hrow + kMaxBatches * TM);    // [kMaxBatches][TM]
    int *s_frames = s_item + kMaxBatches * TM;                  // [kMaxBatches][TM]
    int *s_len = s_frames + kMaxBatches * TM;                   // [kMaxBatches] frames of the batch's longest item
    u64 *inbox_bar = reinterpret_cast<u64 *>(s_len + kMaxBatches + (kMaxBatches & 1));   // [2]
    u64 *part_full = inbox_bar + 2;                             // [2] kWS: the eight warps' partials of a batch-step are in
    u64 *part_empty = part_full + 2;                            // [2] kWS: ... and have been read by the ninth warp

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int lane_b = lane / LANES_J, lane_j = lane % LANES_J;
    const unsigned rank = cluster_rank(), peer = rank ^ 1u;
    const int cluster_id = blockIdx.x / 2;
    const int group = cluster_id / p.NJT, jt = cluster_id % p.NJT;
    const int j0 = jt * TN;
    const unsigned group_ctas = 2u * p.NJT;

    if (tid == 0) {
        // arrivals per phase: one per pushing thread (80), or one per pushing LANE of the ninth warp (16)
        mbar_init(inbox_bar + 0, kWS ? kWarp / 2 : kReduceThreads / 2);
        mbar_init(inbox_bar + 1, kWS ? kWarp / 2 : kReduceThreads / 2);
        for (int k = 0; k < 2; ++k) {
            mbar_init(part_full + k, kWarps);
            mbar_init(part_empty + k, 1);
        }
        mbar_fence_init();
    }
    // the pair's half of the transition block, once per launch
    for (int idx = tid; idx < TN * (KH4 / 4); idx += kAllThreads) {
        const int r = idx / (KH4 / 4), q = idx % (KH4 / 4);
        *reinterpret_cast<float4 *>(A_res + r * LDA + q * 4) =
            __ldg(reinterpret_cast<const float4 *>(p.transition + (int64_t)(j0 + r) * LDP + rank * KH4) + q);
    }
    cluster_sync_all();            // the peer's mbarriers exist before anything arrives on them
    if (group >= p.NG) return;     // (never: the grid is NG * NJT clusters)

    // this warp's share of the CTA's sources, in k-steps (float4 groups)
    const bool compute_warp = warp < kWarps;                    // (kWS: warp 8 is the reducer)
    int kofs = 0;
    for (int w = 0; w < min(warp, kWarps); ++w) kofs += p.nk[w];
    const int nk = compute_warp ? p.nk[warp] : 0;
    const int ns = (nk + kStageSteps - 1) / kStageSteps;        // ring stages per batch-step
    float *wring = ring + min(warp, kWarps - 1) * (kSlots * kSlotFloats);
    const int ld_chunk = (lane % 8) * 4, ld_row = lane / 8;     // cp.async mapping: 16-byte column, rows ld_row + 4 q
    const int64_t src_col0 = (int64_t)rank * KH4 + kofs * 4;    // first source column of this warp
    const float *Aw = A_res + lane_j * LDA + kofs * 4;

    // ownership of the reduced tile: slot u = quad * 32 + lane of the compute mapping; the pair member
    // whose rank equals the parity of lane_b finalises it (all four lanes of a column group agree)
    const int u_quad = tid / kWarp, u_lane = lane;
    const bool reducer = !kWS && tid < kReduceThreads;
    const bool mine_lane = ((u_lane >> 2) & 1) == (int)rank;
    const bool mine = reducer && mine_lane;
    // after the 4 x 4 transpose an owner holds element e = 4 quad + lane_j of lanes lane_j' = 0..3:
    // row lane_b + 8 (e / 5), columns 4 (e % 5) .. + 3 of the tile
    const int fin_e = 4 * u_quad + lane_j;
    const int fin_row = lane_b + LANES_B * (fin_e / N), fin_col = 4 * (fin_e % N);
    const unsigned fin_mask = rank == 0 ? 0x0f0f0f0fu : 0xf0f0f0f0u;
    const bool obs_vec = p.obs_dtype == 0 && (S % 4) == 0 && (reinterpret_cast<uintptr_t>(p.observation) & 15) == 0;
    const bool hist_vec = (S % 4) == 0;
    int ready = p.frames_ready ? 0 : 0x7fffffff;               // frames of `observation` known to have landed
    auto wait_for_frame = [&](int t) {
        if (t >= ready) {
            int r;
            while ((r = ld_acquire_sys(p.frames_ready)) <= t) __nanosleep(200);
            ready = r;
        }
    };

    unsigned *group_counter = p.sync + (size_t)kSyncStride * (1 + p.num_batches + group);
    unsigned group_target = 0;
    unsigned reduce_count = 0;                                  // batch-steps reduced so far: inbox buffer and phase

    const int group_batches = (p.num_batches - group + p.NG - 1) / p.NG;     // batches group, group + NG, ...
    for (int wave0 = 0; wave0 < group_batches; wave0 += kMaxBatches) {
        const int batches = min(kMaxBatches, group_batches - wave0);
        auto global_batch = [&](int b) { return group + p.NG * (wave0 + b); };
        __syncthreads();
        for (int idx = tid; idx < batches * TM; idx += kAllThreads) {
            const int item = p.order[(int64_t)global_batch(idx / TM) * TM + idx % TM];
            s_item[idx] = item;
            s_frames[idx] = item >= 0 ? p.frames[item] : 0;
            s_row[idx] = item >= 0 ? p.item_row[item] : 0;
            s_hrow[idx] = item >= 0 ? p.hist_row[item] : 0;
        }
        __syncthreads();
        if (tid < batches) s_len[tid] = s_frames[tid * TM];    // longest first
        __syncthreads();

        // ---- t = 0: P_0[s] = obs[0][s] + initial[s]   (viterbi.cpp:72-76); rank r takes rows 16 r .. ----
        wait_for_frame(0);
        {
            const int c_end = (jt == p.NJT - 1) ? LDP : min(j0 + TN, LDP);
            const int cols = c_end - j0;
            for (int idx = tid; idx < batches * (TM / 2) * cols; idx += kAllThreads) {
                const int c = j0 + idx % cols, rr = idx / cols;
                const int b = rr / (TM / 2), r = rank * (TM / 2) + rr % (TM / 2);
                const int item = s_item[b * TM + r];
                float *dst0 = p.score0 + ((int64_t)global_batch(b) * TM + r) * LDP + c;
                if (item >= 0 && c < S) {
                    const int64_t e0 = s_row[b * TM + r] * S + c;
                    const float v = convert_observation<kFlags>(load_observation(p.observation, e0, p.obs_dtype)) + p.initial[c];
                    *dst0 = v;
                    p.history[s_hrow[b * TM + r] * S + c] = v;
                } else {
                    dst0[0] = 0.f;                               // padding stays finite, in both buffers
                    dst0[p.score_stride] = 0.f;
                }
            }
        }
        group_target += group_ctas;
        group_barrier(group_counter, group_target);

        // ---- the interleaved recursion ----
        Cursor cc, lc;
        cursor_first(cc, s_len, batches);
        lc = cc;
        int lstage = 0;                                         // stage of `lc`'s batch-step to issue next
        unsigned issued = 0, consumed = 0;                      // ring stages; slot = count % kSlots
        // Issue the next ring stage of this warp (nothing once the schedule is exhausted).  The first
        // stage of a batch-step needs the batch's previous step published by EVERY CTA of the group --
        // this one included.  While a group still interleaves several batches that is some other
        // batch's step; once a single batch is left it is the very step this warp is computing, so the
        // look-ahead issue (`may_block` false, from inside the K loop) is skipped and the stage is
        // issued at the top of the next batch-step instead, after this CTA's own publish.
        auto issue_next = [&](bool may_block) {
            if (lc.n > 0 && ns > 0) {
                const int64_t row0 = (int64_t)global_batch(lc.b) * TM;
                if (lstage == 0 && lc.tau > 1) {
                    if (!may_block && lc.b == cc.b) return;
                    // P_{tau-1} of this batch must have been published by every CTA of the group
                    const unsigned *counter = p.sync + (size_t)kSyncStride * (1 + global_batch(lc.b));
                    const unsigned target = group_ctas * (unsigned)(lc.tau - 1);
                    if (lane == 0)
                        while (ld_acquire(counter) < target) {}
                    __syncwarp();
                }
                const int steps = min(kStageSteps, nk - lstage * kStageSteps);
                if (ld_chunk < steps * 4) {
                    const float *src = p.score0 + ((lc.tau - 1) & 1) * p.score_stride + (row0 + ld_row) * LDP + src_col0 +
                                       lstage * (kStageSteps * 4) + ld_chunk;
                    float *dst = wring + (issued % kSlots) * kSlotFloats + ld_row * kRowStride + ld_chunk;
#pragma unroll
                    for (int q = 0; q < TM / 4; ++q) cp_async16(dst + q * 4 * kRowStride, src + (int64_t)q * 4 * LDP);
                }
                if (++lstage == ns) {
                    lstage = 0;
                    cursor_next(lc, s_len);
                }
                cp_async_commit();
                ++issued;
            }
        };

        CPROF_DECL
        if (kWS && !compute_warp) {
            // ================= the ninth warp: reduce, exchange, publish =================
            Cursor rc;
            cursor_first(rc, s_len, batches);
            while (rc.n > 0) {
                const int b = rc.b, tau = rc.tau;
                const int64_t row0 = (int64_t)global_batch(b) * TM;
                // emissions of the cells this lane finalises (one float4 per quad), requested before the wait
                float4 emit[kQuads];
                bool live[kQuads];
                int64_t hrow[kQuads];
#pragma unroll
                for (int q = 0; q < kQuads; ++q) {
                    const int e = 4 * q + lane_j;
                    const int frow = lane_b + LANES_B * (e / N), fcol = 4 * (e % N);
                    const int item = mine_lane ? s_item[b * TM + frow] : -1;
                    live[q] = mine_lane && item >= 0 && tau < s_frames[b * TM + frow] && j0 + fcol < S;
                    hrow[q] = mine_lane ? s_hrow[b * TM + frow] : 0;
                    emit[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live[q]) {
                        if (p.frames_ready) wait_for_frame(tau);
                        const int64_t e0 = (s_row[b * TM + frow] + tau) * S + j0 + fcol;
                        if (obs_vec) {
                            emit[q] = __ldcg(reinterpret_cast<const float4 *>(p.observation + e0));
                        } else {
                            emit[q].x = load_observation(p.observation, e0, p.obs_dtype);
                            if (j0 + fcol + 1 < S) emit[q].y = load_observation(p.observation, e0 + 1, p.obs_dtype);
                            if (j0 + fcol + 2 < S) emit[q].z = load_observation(p.observation, e0 + 2, p.obs_dtype);
                            if (j0 + fcol + 3 < S) emit[q].w = load_observation(p.observation, e0 + 3, p.obs_dtype);
                        }
                    }
                }
                const unsigned par = reduce_count & 1u, phase = (reduce_count >> 1) & 1u;
                CPROF_MARK(3)
                mbar_wait(part_full + par, phase);                  // the eight warps' partials are in
                CPROF_MARK(4)
                float4 m[kQuads];
#pragma unroll
                for (int q = 0; q < kQuads; ++q) {
                    const float4 *col = partial + par * kWarps * kReduceThreads + q * kWarp + lane;
                    m[q] = col[0];
#pragma unroll
                    for (int w = 1; w < kWarps; ++w) m[q] = max4(m[q], col[w * kReduceThreads]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(part_empty + par);       // the buffer may be refilled
                CPROF_MARK(5)
                if (!mine_lane) {
#pragma unroll
                    for (int q = 0; q < kQuads; ++q)
                        st_cluster_f4(map_to_rank(smem_u32(inbox + par * kReduceThreads + q * kWarp + lane), peer), m[q]);
                    mbar_arrive_remote(map_to_rank(smem_u32(inbox_bar + par), peer));
                } else {
                    mbar_wait_cluster(inbox_bar + par, phase);
#pragma unroll
                    for (int q = 0; q < kQuads; ++q) {
                        float4 v = max4(m[q], inbox[par * kReduceThreads + q * kWarp + lane]);
                        // 4 x 4 transpose across the column group (see the 256-thread path below)
                        float s0 = (lane_j & 1) ? v.x : v.y, s1 = (lane_j & 1) ? v.z : v.w;
                        s0 = __shfl_xor_sync(fin_mask, s0, 1);
                        s1 = __shfl_xor_sync(fin_mask, s1, 1);
                        if (lane_j & 1) { v.x = s0; v.z = s1; } else { v.y = s0; v.w = s1; }
                        float t0 = (lane_j & 2) ? v.x : v.z, t1 = (lane_j & 2) ? v.y : v.w;
                        t0 = __shfl_xor_sync(fin_mask, t0, 2);
                        t1 = __shfl_xor_sync(fin_mask, t1, 2);
                        if (lane_j & 2) { v.x = t0; v.y = t1; } else { v.z = t0; v.w = t1; }
                        if (live[q]) {
                            const int e = 4 * q + lane_j;
                            const int frow = lane_b + LANES_B * (e / N), col0 = j0 + 4 * (e % N);
                            float4 o;
                            o.x = convert_observation<kFlags>(emit[q].x) + v.x;
                            o.y = convert_observation<kFlags>(emit[q].y) + v.y;
                            o.z = convert_observation<kFlags>(emit[q].z) + v.z;
                            o.w = convert_observation<kFlags>(emit[q].w) + v.w;
                            float *dst = p.score0 + (tau & 1) * p.score_stride + (row0 + frow) * LDP + col0;
                            float *hist = p.history + (hrow[q] + tau) * S + col0;
                            if (col0 + 3 < S) {
                                *reinterpret_cast<float4 *>(dst) = o;
                                if (hist_vec) __stcs(reinterpret_cast<float4 *>(hist), o);
                                else { hist[0] = o.x; hist[1] = o.y; hist[2] = o.z; hist[3] = o.w; }
                            } else {
                                const float ee[4] = {o.x, o.y, o.z, o.w};
                                for (int x = 0; col0 + x < S; ++x) { dst[x] = ee[x]; hist[x] = ee[x]; }
                            }
                        }
                    }
                }
                // one arrival per CTA on the batch's publish counter, after all of its stores
                __syncwarp();
                CPROF_MARK(0)
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(p.sync + (size_t)kSyncStride * (1 + global_batch(b)), 1u);
                }
                CPROF_MARK(1)
                ++reduce_count;
                cursor_next(rc, s_len);
#ifdef TORBI_TILED_PROFILE
                ++prof_steps;
#endif
            }
        } else
        while (cc.n > 0) {
            const int b = cc.b, tau = cc.tau;
            const int64_t row0 = (int64_t)global_batch(b) * TM;
            // emission of the cells this thread finalises, fetched now and used after the K loop
            float4 emit = make_float4(0.f, 0.f, 0.f, 0.f);
            const int fin_item = mine ? s_item[b * TM + fin_row] : -1;
            const int64_t fin_row0 = mine ? s_row[b * TM + fin_row] : 0;
            const int64_t fin_hrow0 = mine ? s_hrow[b * TM + fin_row] : 0;
            const bool fin_live = mine && fin_item >= 0 && tau < s_frames[b * TM + fin_row] && j0 + fin_col < S;
            if (fin_live) {
                if (p.frames_ready) wait_for_frame(tau);
                const int64_t e0 = (fin_row0 + tau) * S + j0 + fin_col;
                if (obs_vec) {
                    emit = __ldcg(reinterpret_cast<const float4 *>(p.observation + e0));
                } else {
                    emit.x = load_observation(p.observation, e0, p.obs_dtype);
                    if (j0 + fin_col + 1 < S) emit.y = load_observation(p.observation, e0 + 1, p.obs_dtype);
                    if (j0 + fin_col + 2 < S) emit.z = load_observation(p.observation, e0 + 2, p.obs_dtype);
                    if (j0 + fin_col + 3 < S) emit.w = load_observation(p.observation, e0 + 3, p.obs_dtype);
                }
            }

            float cur[M][N];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c) cur[i][c] = -INFINITY;

            for (int st = 0; st < ns; ++st) {
                if (issued == consumed) issue_next(true);       // nothing in flight: this very stage (see issue_next)
                cp_async_wait_all();                            // this lane's copies of the stage have landed ...
                __syncwarp();                                   // ... and so have the other lanes'; all lanes are past the previous stage
                issue_next(false);                              // into the slot read one stage ago
                const float *Ps = wring + (consumed % kSlots) * kSlotFloats + lane_b * kRowStride;
                const float *As = Aw + st * (kStageSteps * 4);
                ++consumed;
                const int kmax = min(kStageSteps, nk - st * kStageSteps);
                // full stages run without a branch between k-steps, so the operand loads of the next
                // k-step are scheduled under the math of the current one
                auto k_steps = [&](auto partial_stage) {
#pragma unroll
                    for (int k = 0; k < kStageSteps; ++k) {
                        if constexpr (decltype(partial_stage)::value) {
                            if (k >= kmax) break;
                        }
                        ulonglong2 pv[M], av[N];
#pragma unroll
                        for (int i = 0; i < M; ++i)
                            pv[i] = *reinterpret_cast<const ulonglong2 *>(Ps + i * LANES_B * kRowStride + k * 4);
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            av[c] = *reinterpret_cast<const ulonglong2 *>(As + c * LANES_J * LDA + k * 4);
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
                if (kmax == kStageSteps) k_steps(std::false_type{});
                else k_steps(std::true_type{});
            }
            CPROF_MARK(0)

            // ---- the 16 partial maxima of every cell meet: 8 warps through shared memory ... ----
            const unsigned par = reduce_count & 1u;
            float4 *my_partial = partial + (par * kWarps + warp) * kReduceThreads;
            if (kWS && reduce_count >= 2)                       // the ninth warp has read what this buffer held
                mbar_wait(part_empty + par, ((reduce_count >> 1) - 1u) & 1u);
#pragma unroll
            for (int q = 0; q < kQuads; ++q) {
                // element e = i * N + c of the thread's tile; quad q holds e = 4 q .. 4 q + 3
                float v[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) v[x] = cur[(4 * q + x) / N][(4 * q + x) % N];
                my_partial[q * kWarp + lane] = make_float4(v[0], v[1], v[2], v[3]);
            }
            if constexpr (kWS) {
                __syncwarp();
                if (lane == 0) mbar_arrive(part_full + par);    // (release: the partials are visible to the waiter)
            } else {
                __syncthreads();
            }
            CPROF_MARK(1)
            if (reducer) {
                const float4 *col = partial + par * kWarps * kReduceThreads + tid;
                float4 m = col[0];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) m = max4(m, col[w * kReduceThreads]);
                // ---- ... and the two CTAs of the pair through distributed shared memory ----
                if (!mine) {
                    st_cluster_f4(map_to_rank(smem_u32(inbox + par * kReduceThreads + tid), peer), m);
                    mbar_arrive_remote(map_to_rank(smem_u32(inbox_bar + par), peer));
                } else {
                    mbar_wait_cluster(inbox_bar + par, (reduce_count >> 1) & 1u);
                    m = max4(m, inbox[par * kReduceThreads + tid]);
                    // 4 x 4 transpose across the column group: this lane gets element 4 quad + lane_j of all four lanes,
                    // i.e. four CONSECUTIVE destination states of one sequence
                    {
                        float s0 = (lane_j & 1) ? m.x : m.y, s1 = (lane_j & 1) ? m.z : m.w;
                        s0 = __shfl_xor_sync(fin_mask, s0, 1);
                        s1 = __shfl_xor_sync(fin_mask, s1, 1);
                        if (lane_j & 1) { m.x = s0; m.z = s1; } else { m.y = s0; m.w = s1; }
                        float t0 = (lane_j & 2) ? m.x : m.z, t1 = (lane_j & 2) ? m.y : m.w;
                        t0 = __shfl_xor_sync(fin_mask, t0, 2);
                        t1 = __shfl_xor_sync(fin_mask, t1, 2);
                        if (lane_j & 2) { m.x = t0; m.y = t1; } else { m.z = t0; m.w = t1; }
                    }
                    // new scores: obs[t][j] + max  (viterbi.cpp:102), published for the other pairs + history
                    if (fin_live) {
                        const int col0 = j0 + fin_col;
                        float4 v;
                        v.x = convert_observation<kFlags>(emit.x) + m.x;
                        v.y = convert_observation<kFlags>(emit.y) + m.y;
                        v.z = convert_observation<kFlags>(emit.z) + m.z;
                        v.w = convert_observation<kFlags>(emit.w) + m.w;
                        float *dst = p.score0 + (tau & 1) * p.score_stride + (row0 + fin_row) * LDP + col0;
                        float *hist = p.history + (fin_hrow0 + tau) * S + col0;
                        if (col0 + 3 < S) {
                            *reinterpret_cast<float4 *>(dst) = v;
                            if (hist_vec) __stcs(reinterpret_cast<float4 *>(hist), v);
                            else { hist[0] = v.x; hist[1] = v.y; hist[2] = v.z; hist[3] = v.w; }
                        } else {
                            const float e[4] = {v.x, v.y, v.z, v.w};
                            for (int x = 0; col0 + x < S; ++x) { dst[x] = e[x]; hist[x] = e[x]; }
                        }
                    }
                }
                // one arrival per CTA on the batch's publish counter, after all of its stores
                named_barrier(1, kReduceThreads);
                if (tid == 0) {
                    __threadfence();
                    atomicAdd(p.sync + (size_t)kSyncStride * (1 + global_batch(b)), 1u);
                }
            }
            ++reduce_count;
            cursor_next(cc, s_len);
            CPROF_MARK(2)
#ifdef TORBI_TILED_PROFILE
            ++prof_steps;
#endif
        }
        cp_async_wait_all();
        CPROF_REDUCER_PRINT
#ifdef TORBI_TILED_PROFILE
        if (lane == 0 && blockIdx.x % 37 == 0 && (warp == 0 || warp == 7) && prof_steps > 0)
            printf("cta %d warp %d batch-steps %lld cycles/batch-step: K loop %lld, partial store + CTA barrier %lld, reduce + DSMEM + publish %lld\n",
                   blockIdx.x, warp, prof_steps, prof_acc[0] / prof_steps, prof_acc[1] / prof_steps, prof_acc[2] / prof_steps);
#endif

        // ---- final scores out (viterbi.cpp:111-113) + first-occurrence argmax (:218) ----
        group_target += group_ctas;
        group_barrier(group_counter, group_target);
        const int cta_in_group = jt * 2 + (int)rank;
        constexpr int kAllWarps = kAllThreads / kWarp;
        for (int rr = cta_in_group * kAllWarps + warp; rr < batches * TM; rr += (int)group_ctas * kAllWarps) {
            const int item = s_item[rr];
            if (item < 0) continue;
            const int b = rr / TM, r = rr % TM;
            // a sequence's last scores are in the buffer of ITS last step (finished rows are not touched again)
            const float *row = p.score0 + ((s_frames[rr] - 1) & 1) * p.score_stride + ((int64_t)global_batch(b) * TM + r) * LDP;
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
    // no CTA of the pair may exit while its peer can still push into its shared memory
    cluster_sync_all();
}

int cluster_njt(int64_t S) { return (int)((S + TN - 1) / TN); }

// k-steps per warp within a CTA's half: warps 0..4 also reduce / publish after every batch-step, so
// the other three take `bonus` k-steps more (TORBI_B200_CLUSTER_BONUS, default 1)
void split_k_steps(int half_steps, int bonus, int *nk) {
    const int reducers = kReduceThreads / kWarp, others = kWarps - reducers;
    if (half_steps < kWarps + others * bonus) bonus = 0;
    const int base = (half_steps - others * bonus) / kWarps;
    int rest = half_steps;
    for (int w = 0; w < kWarps; ++w) {
        nk[w] = base + (w >= reducers ? bonus : 0);
        rest -= nk[w];
    }
    for (int w = kWarps - 1; rest > 0; w = (w + kWarps - 1) % kWarps) {     // remainder: from the last warp backwards
        ++nk[w];
        --rest;
    }
}

}  // namespace

bool forward_cluster_supported(const ForwardArgs &a) {
    const int njt = cluster_njt(a.S);
    const int64_t batches = (a.B + TM - 1) / TM;
    if (2 * njt > a.sm_count || a.ws.transition_pad == nullptr) return false;
    const int ng = (int)std::min<int64_t>(a.sm_count / 2 / njt, batches);
    // publish counters (one per batch) and group counters must fit the scheduling words below the band area
    return (size_t)kSyncStride * (1 + batches + ng + 1) <= kBandWordOffset &&
           cluster_smem_bytes((int)padded_sources(a.S)) <= 227 * 1024;
}

// Estimated seconds: the batches of a group advance in rounds; a round costs one batch-step per running
// batch (measured on the B200 at S = 1440, scales with the K loop), plus the exposed exchange when a
// group is down to a single batch.
double cluster_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count) {
    const int njt = cluster_njt(S);
    const int batches = (int)((B + TM - 1) / TM);
    const int groups = std::max(1, std::min(sm_count / 2 / std::max(njt, 1), batches));
    const double step = 4.7e-6 * (double)padded_sources(S) / 1440.0, exposed = 2.0e-6;
    double worst = 0.0;
    for (int g = 0; g < groups; ++g) {
        double seconds = 0.0;
        int prev = 0;
        // batches g, g + groups, ... (longest first): between consecutive lengths the number running is constant
        std::vector<int> lens;
        for (int b = g; b < batches; b += groups) lens.push_back(frames_sorted[(int64_t)b * TM]);
        for (int k = (int)lens.size() - 1; k >= 0; --k) {
            const int rounds = lens[k] - prev;                  // rounds during which k + 1 batches run
            prev = lens[k];
            seconds += rounds * ((k + 1) * step + (k == 0 ? exposed : 0.0));
        }
        worst = std::max(worst, seconds);
    }
    return worst;
}

cudaError_t launch_forward_cluster(const ForwardArgs &a, cudaStream_t stream) {
    ClusterParams p{};
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
    p.NJT = cluster_njt(a.S);
    p.num_batches = (int)((a.B + TM - 1) / TM);
    p.NG = std::min(a.sm_count / 2 / p.NJT, p.num_batches);
    // TORBI_B200_CLUSTER_WS=1: ninth (reducer) warp.  Measured slower than the 256-thread kernel (3.32 M vs 3.81 M
    // timesteps/s at 64 sequences): the single reducing warp needs ~12 k cycles per batch-step -- more than a K loop
    bool ws = false;
    if (const char *env = std::getenv("TORBI_B200_CLUSTER_WS")) ws = std::atoi(env) != 0;
    int bonus = ws ? 0 : 1;
    if (const char *env = std::getenv("TORBI_B200_CLUSTER_BONUS")) bonus = std::max(0, std::min(8, std::atoi(env)));
    split_k_steps(p.LDP / 8, bonus, p.nk);

    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    const size_t smem = cluster_smem_bytes(p.LDP);
    auto launch = [&](auto kernel) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        cudaLaunchConfig_t config{};
        config.gridDim = dim3(2 * p.NG * p.NJT);
        config.blockDim = dim3(ws ? kThreads + kWarp : kThreads);
        config.dynamicSmemBytes = smem;
        config.stream = stream;
        cudaLaunchAttribute attrs[2];
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = 2;
        attrs[0].val.clusterDim.y = 1;
        attrs[0].val.clusterDim.z = 1;
        // all pairs of a group spin on each other's counters: they must be co-resident, and the
        // cooperative attribute makes the launch fail instead of deadlocking when they cannot be
        attrs[1].id = cudaLaunchAttributeCooperative;
        attrs[1].val.cooperative = 1;
        config.attrs = attrs;
        config.numAttrs = 2;
        void *args[] = {(void *)&p};
        return cudaLaunchKernelExC(&config, (const void *)kernel, args);
    };
    switch (a.flags & 3u) {
        case 0: return ws ? launch(forward_cluster_kernel<0, true>) : launch(forward_cluster_kernel<0, false>);
        case 1: return ws ? launch(forward_cluster_kernel<1, true>) : launch(forward_cluster_kernel<1, false>);
        case 2: return ws ? launch(forward_cluster_kernel<2, true>) : launch(forward_cluster_kernel<2, false>);
        default: return ws ? launch(forward_cluster_kernel<3, true>) : launch(forward_cluster_kernel<3, false>);
    }
}

}  // namespace torbi_b200
