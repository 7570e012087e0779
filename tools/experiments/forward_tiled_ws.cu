// EXPERIMENT, not built into libtorbi_b200.so: warp-specialised variant of the tiled kernel (round 1).
// Measured slower than the classic kernel (profiles/r01_ws_experiment.txt); kept as the record of the A/B.
// Warp-specialised variant of the batch-tiled forward kernel (forward_tiled.cu)
// for full 64-sequence batches.  Same tiling, same arithmetic, same results; the
// difference is WHO does what, so that nothing but the max-plus inner loop sits on
// the compute warps' critical path:
//
//   warps 0-7   compute   register-tiled FADD2 / FMNMX3 loop, emission + publish
//   warp  8     loader    cp.async of the A / P / emission tiles into a 2-stage ring,
//                         completion signalled through mbarriers (no __syncthreads
//                         per stage); also waits for the group's score exchange
//   warps 9-11  re-scan   exact lowest-index argmax of step t and the coalesced
//                         uint16 backpointer stores, WHILE the compute warps already
//                         run the inner loop of step t+1
//
// Registers are redistributed with setmaxnreg (compute 192, the two auxiliary
// warpgroups 64: 256 x 192 + 256 x 64 = 512 x 128, the CTA's launch allocation --
// setmaxnreg can only move registers inside that pool).
//
// STATUS: experimental, opt-in (TORBI_B200_TILED_WS), bit-exact, but SLOWER than
// the classic kernel on the B200 (189.9 ms vs 174.4 ms at S=1440, B=512, T=2000).
// Measured cause: the recursion needs a group-wide score exchange every timestep;
// the classic kernel fills that ~10 us gap with the re-scan, whereas here the
// compute warps idle through it (skeleton without any re-scan work: 171.8 ms vs
// 152.0 ms classic) and the re-scan warps then compete for the same issue slots.
// Kept as the measured record of that experiment (profiles/r01_ws_experiment.txt).
// The max / winning-chunk tiles are double buffered and the score buffers rotate
// over four slots, so a CTA that runs ahead never overwrites what a slower CTA of
// its group (or its own re-scan warps) still reads.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

namespace {

constexpr int TM = 64;
constexpr int TN = kTileDests;
constexpr int KC = 32;
constexpr int KS = kStageSources;
constexpr int LDS_ = KS + 4;
constexpr int NSTAGE = 2;
constexpr int M = 4, N = 5;
constexpr int LANES_B = 8, LANES_J = 4, WARPS_J = 4;
constexpr int kComputeThreads = 256, kComputeWarps = 8;
// warps 8-11: re-scan (one per SM sub-partition, so the extra issue load is even),
// warp 12: loader, warps 13-15 only take part in the CTA-wide barriers
constexpr int kFirstRescanWarp = 8;
constexpr int kRescanWarps = 4, kRescanThreads = kRescanWarps * 32;
constexpr int kLoaderWarp = 12;
constexpr int kThreadsWS = 512;
constexpr int kScoreBuffers = 4;

constexpr int kStageFloats = (TM + TN) * LDS_;
constexpr size_t kSmemWS =
    sizeof(float) * (NSTAGE * kStageFloats + TM * TN + 2 * TM * TN)   // ring, emission/score tile, 2 max tiles
    + sizeof(uint16_t) * 2 * TM * TN                                   // 2 chunk -> backpointer tiles
    + sizeof(int) * 2 * TM                                             // slot item / frames
    + sizeof(unsigned long long) * 16;                                 // mbarriers
static_assert(kSmemWS <= 227 * 1024, "shared memory budget");

typedef unsigned long long u64;

__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem));
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(u64 *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bulk asynchronous copy (TMA engine, async proxy): one instruction moves a whole row;
// completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_copy(void *smem, const void *gmem, unsigned bytes, u64 *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_acquire_sys(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

struct WSParams {
    const float *observation;
    const float *transition;    // padded copy (SJ, LDP)
    const float *initial;
    const int32_t *frames;
    const int32_t *order;
    float *score0;              // kScoreBuffers rotating buffers, `score_stride` floats apart
    int64_t score_stride;
    uint16_t *backptr;
    float *posterior;
    int32_t *final_state;
    unsigned *sync;
    const int32_t *frames_ready;
    int64_t T;
    int S, LDP, NJT, NG, num_batches;
};

// mbarrier slots
enum { kFull0 = 0, kFull1, kEmpty0, kEmpty1, kEpiDone0, kEpiDone1, kRescanDone0, kRescanDone1, kNumBars };

template <uint32_t kFlags>
__global__ void __launch_bounds__(kThreadsWS, 1) forward_tiled_ws_kernel(const WSParams p) {
    extern __shared__ __align__(16) float smem[];
    float *ring = smem;
    float *tile = smem + NSTAGE * kStageFloats;                 // [TM][TN] emission -> new score
    float *best_tiles = tile + TM * TN;                         // 2 x [TM][TN]
    uint16_t *bp_tiles = reinterpret_cast<uint16_t *>(best_tiles + 2 * TM * TN);   // 2 x [TM][TN]
    int *slot_item = reinterpret_cast<int *>(bp_tiles + 2 * TM * TN);
    int *slot_frames = slot_item + TM;
    u64 *bars = reinterpret_cast<u64 *>(slot_frames + TM);
    __shared__ int batch_shared;

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int group = blockIdx.x / p.NJT, jt = blockIdx.x % p.NJT;
    const int j0 = jt * TN;
    unsigned *counter = p.sync + group * 32;
    unsigned *queue_head = p.sync + p.NG * 32;
    const int S = p.S, LDP = p.LDP;
    const int nstages = LDP / KS;
    // (launch-time precondition: S % 8 == 0 and 16-byte aligned observations, so the
    //  emission tile and the backpointers move as 16-byte vectors)

    if (tid == 0) {
        mbar_init(bars + kFull0, 32);          mbar_init(bars + kFull1, 32);
        mbar_init(bars + kEmpty0, kComputeWarps); mbar_init(bars + kEmpty1, kComputeWarps);
        mbar_init(bars + kEpiDone0, kComputeWarps); mbar_init(bars + kEpiDone1, kComputeWarps);
        mbar_init(bars + kRescanDone0, kRescanWarps); mbar_init(bars + kRescanDone1, kRescanWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Every role runs its own copy of the batch loop, entirely inside the branch
    // that set its register budget (ptxas sizes each region by the setmaxnreg that
    // dominates it).  The roles meet at CTA-wide barriers at the batch boundaries.
    unsigned barriers = 0;         // group barrier arrivals expected so far (tracked by every role)

    // next batch off the longest-first queue; returns false when the queue is empty
    auto next_batch = [&](int64_t &row0, int &batch_frames) {
        if (tid == 0) {
            if (jt == 0) counter[1] = atomicAdd(queue_head, 1u);
            __threadfence();
            atomicAdd(counter, 1u);
            while (ld_acquire(counter) < barriers + p.NJT) {}
            batch_shared = (int)__ldcg(counter + 1);
        }
        barriers += p.NJT;
        __syncthreads();
        const int batch = batch_shared;
        if (batch >= p.num_batches) return false;
        row0 = (int64_t)batch * TM;
        if (tid < TM) {
            const int item = p.order[row0 + tid];
            slot_item[tid] = item;
            slot_frames[tid] = item >= 0 ? p.frames[item] : 0;
        }
        __syncthreads();
        batch_frames = slot_frames[0];
        return true;
    };

    if (warp < kComputeWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");   // 256 x 192 + 256 x 64 = 512 x 128
        unsigned stage_count = 0;
        unsigned step_count[2] = {0, 0};
        int64_t row0;
        int batch_frames;
        while (next_batch(row0, batch_frames)) {
            // =========================== COMPUTE WARPS ===========================
            const int wb = warp / WARPS_J, wj = warp % WARPS_J;
            const int lane_b = lane / LANES_J, lane_j = lane % LANES_J;
            const int r_base = wb * (LANES_B * M) + lane_b;
            const int c_base = wj * (LANES_J * N) + lane_j;

            // t = 0 (viterbi.cpp:72-76); the loader has already waited for frame 0
            if (p.frames_ready) {
                if (tid == 0) while (ld_acquire_sys(p.frames_ready) <= 0) __nanosleep(200);
                named_barrier(1, kComputeThreads);
            }
            {
                const int c_end = (jt == p.NJT - 1) ? LDP : min(j0 + TN, LDP);
                for (int r = warp; r < TM; r += kComputeWarps) {
                    const int item = slot_item[r];
                    float *dst0 = p.score0 + (row0 + r) * LDP;
                    const float *obs = p.observation + (int64_t)max(item, 0) * p.T * S;
                    for (int c = j0 + lane; c < c_end; c += kWarp) {
                        if (item >= 0 && c < S) {
                            dst0[c] = convert_observation<kFlags>(obs[c]) + p.initial[c];
                        } else {
#pragma unroll
                            for (int k = 0; k < kScoreBuffers; ++k) dst0[c + k * p.score_stride] = 0.f;
                        }
                    }
                }
            }
            barriers += p.NJT;
            named_barrier(1, kComputeThreads);
            if (tid == 0) { __threadfence(); atomicAdd(counter, 1u); }

            int buf = 0;
            for (int t = 1; t < batch_frames; ++t) {
                const int buf_new = (buf + 1) % kScoreBuffers;
                float *P_new = p.score0 + buf_new * p.score_stride;
                buf = buf_new;
                const int pb = t & 1;                       // tile buffer of this step
                float *best_tile = best_tiles + pb * TM * TN;
                uint16_t *bp_tile = bp_tiles + pb * TM * TN;

                float cur[M][N];
                int bestc[M][N];
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int c = 0; c < N; ++c) { cur[i][c] = -INFINITY; bestc[i][c] = 0; }

                for (int s = 0; s < nstages; ++s, ++stage_count) {
                    const int slot = stage_count & 1;
                    mbar_wait(bars + kFull0 + slot, (stage_count >> 1) & 1);
                    const float *Ps = ring + slot * kStageFloats + r_base * LDS_;
                    const float *As = ring + slot * kStageFloats + (TM + c_base) * LDS_;
#pragma unroll
                    for (int ch = 0; ch < KS / KC; ++ch) {
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
                                    const float a0 = (k == 0) ? cur[i][c] : acc[i][c];
                                    acc[i][c] = fmaxf(fmaxf(fmaxf(fmaxf(a0, lo(v0)), hi(v0)), lo(v1)), hi(v1));
                                }
                        }
                        const int chunk = s * (KS / KC) + ch;
#pragma unroll
                        for (int i = 0; i < M; ++i)
#pragma unroll
                            for (int c = 0; c < N; ++c) {
                                if (acc[i][c] > cur[i][c]) bestc[i][c] = chunk;
                                cur[i][c] = acc[i][c];
                            }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + kEmpty0 + slot);      // this warp is done with the slot
                }

                // the re-scan warps must be done with this tile buffer (used two steps ago)
                if (step_count[pb] > 0) mbar_wait(bars + kRescanDone0 + pb, (step_count[pb] - 1) & 1);

#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const int r = r_base + i * LANES_B;
                    if (t < slot_frames[r]) {
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            const int col = c_base + c * LANES_J;
                            if (j0 + col < S) {
                                const float emit = convert_observation<kFlags>(tile[r * TN + col]);
                                tile[r * TN + col] = emit + cur[i][c];
                                best_tile[r * TN + col] = cur[i][c];
                                bp_tile[r * TN + col] = (uint16_t)bestc[i][c];
                            }
                        }
                    }
                }
                named_barrier(1, kComputeThreads);
                if (lane == 0) mbar_arrive(bars + kEpiDone0 + pb);          // re-scan of step t may start
                ++step_count[pb];

                // publish (finished rows carried over from the previous buffer)
                const float *P_cur = p.score0 + ((buf + kScoreBuffers - 1) % kScoreBuffers) * p.score_stride;
                for (int idx = tid; idx < TM * (TN / 4); idx += kComputeThreads) {
                    const int r = idx / (TN / 4), c4 = idx % (TN / 4);
                    const int col = j0 + c4 * 4;
                    if (col >= S || slot_item[r] < 0) continue;
                    const int64_t off = (row0 + r) * LDP + col;
                    const bool live = t < slot_frames[r];
                    if (col + 3 < S) {
                        const float4 v = live ? *reinterpret_cast<const float4 *>(tile + r * TN + c4 * 4)
                                              : __ldcg(reinterpret_cast<const float4 *>(P_cur + off));
                        *reinterpret_cast<float4 *>(P_new + off) = v;
                    } else {
                        for (int e = 0; col + e < S; ++e)
                            P_new[off + e] = live ? tile[r * TN + c4 * 4 + e] : __ldcg(P_cur + off + e);
                    }
                }
                barriers += p.NJT;
                named_barrier(1, kComputeThreads);
                if (tid == 0) { __threadfence(); atomicAdd(counter, 1u); }
            }

            // final scores + first-occurrence argmax (viterbi.cpp:111-113, :218)
            if (tid == 0) while (ld_acquire(counter) < barriers) {}
            named_barrier(1, kComputeThreads);
            const float *P_fin = p.score0 + buf * p.score_stride;
            for (int r = jt * kComputeWarps + warp; r < TM; r += p.NJT * kComputeWarps) {
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
            __syncthreads();       // the batch is completely finished (tiles, ring and slots reusable)
        }
    } else {
      // one setmaxnreg for the whole third warpgroup (loader + re-scan warps)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
      if (warp == kLoaderWarp) {
        unsigned stage_count = 0;
        int64_t row0;
        int batch_frames;
        while (next_batch(row0, batch_frames)) {
            // ============================ LOADER WARP ============================
            barriers += p.NJT;                                   // the t = 0 exchange
            int ready = p.frames_ready ? 0 : 0x7fffffff;
            int buf = 0;
            // Lane c4 < 24 owns float4 column c4 of a [rows][24] tile and walks the rows with
            // constant strides: no divisions, no divergence, ~4 instructions per copy.
            auto copy_rows = [&](float *dst_tile, const float *src, int rows) {
                if (lane < KS / 4) {
                    float *d = dst_tile + lane * 4;
                    const float *g = src + lane * 4;
#pragma unroll 4
                    for (int r = 0; r < rows; ++r) {
                        cp_async16(d, g);
                        d += LDS_;
                        g += LDP;
                    }
                }
            };
            for (int t = 1; t < batch_frames; ++t) {
                const float *P_cur = p.score0 + buf * p.score_stride + row0 * LDP;
                buf = (buf + 1) % kScoreBuffers;
                for (int s = 0; s < nstages; ++s, ++stage_count) {
                    const int slot = stage_count & 1;
                    mbar_wait(bars + kEmpty0 + slot, ((stage_count >> 1) & 1) ^ 1);   // slot drained
                    float *stage = ring + slot * kStageFloats;
                    const int k0 = s * KS;
                    // the transition rows do not depend on the other CTAs: issue them first
                    copy_rows(stage + TM * LDS_, p.transition + (int64_t)j0 * LDP + k0, TN);
                    if (s == 0) {
                        // scores of step t-1 published by the whole group; frame t copied in
                        if (lane == 0) {
                            while (ld_acquire(counter) < barriers) {}
                            if (t >= ready) {
                                int r;
                                while ((r = ld_acquire_sys(p.frames_ready)) <= t) __nanosleep(200);
                                ready = r;
                            }
                        }
                        ready = __shfl_sync(kFullMask, ready, 0);
                        barriers += p.NJT;
                        // emission tile travels with the first stage: lane c4 < 20 owns a float4 column
                        if (lane < TN / 4 && j0 + lane * 4 < S) {
                            const float *obs_t = p.observation + (int64_t)t * S + j0 + lane * 4;
                            const int64_t item_stride = p.T * S;
#pragma unroll 4
                            for (int r = 0; r < TM; ++r) {
                                const int item = slot_item[r];
                                if (item >= 0 && t < slot_frames[r])
                                    cp_async16(tile + r * TN + lane * 4, obs_t + item * item_stride);
                            }
                        }
                    }
                    copy_rows(stage, P_cur + k0, TM);
                    cp_async_arrive(bars + kFull0 + slot);
                }
            }
            __syncthreads();
        }
      } else if (warp >= kLoaderWarp + 1) {
        // spare warps of the loader's warpgroup: only the CTA-wide barriers
        int64_t row0;
        int batch_frames;
        while (next_batch(row0, batch_frames)) {
            barriers += (unsigned)p.NJT * batch_frames;
            __syncthreads();
        }
      } else {
        unsigned step_count[2] = {0, 0};
        int64_t row0;
        int batch_frames;
        while (next_batch(row0, batch_frames)) {
            // =========================== RE-SCAN WARPS ===========================
            const int rw = warp - kFirstRescanWarp;              // 0..3
            const int rtid = tid - kFirstRescanWarp * 32;        // 0..127
            const int sub = lane >> 3, q4 = (lane & 7) * 4;
            const bool full_tile = j0 + TN <= S;
            const int cols_valid = S - j0;
            const float *Aq = p.transition + (int64_t)(j0 + sub) * LDP + q4;
            int buf = 0;
            for (int t = 1; t < batch_frames; ++t) {
                const int pb = t & 1;
                const float *best_tile = best_tiles + pb * TM * TN;
                uint16_t *bp_tile = bp_tiles + pb * TM * TN;
                const float *P_rows = p.score0 + buf * p.score_stride + row0 * LDP;
                buf = (buf + 1) % kScoreBuffers;
                mbar_wait(bars + kEpiDone0 + pb, step_count[pb] & 1);
                ++step_count[pb];

                constexpr int kBatch = 4;
#ifndef TORBI_EXPERIMENT_SKIP_RESCAN
                for (int r = rw; r < TM; r += kRescanWarps) {
                    if (t >= slot_frames[r]) continue;
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
                            const bool valid = full_tile || (cb + u) * 4 + sub < cols_valid;
                            if (!valid) k0[u] = 0;
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
                            key = min(key, __shfl_xor_sync(kFullMask, key, 1));
                            key = min(key, __shfl_xor_sync(kFullMask, key, 2));
                            key = min(key, __shfl_xor_sync(kFullMask, key, 4));
                            k0[u] = min(k0[u] + (int)key, S - 1);
                        }
                        __syncwarp();
#pragma unroll
                        for (int u = 0; u < kBatch; ++u)
                            if ((lane & 7) == 0 && (full_tile || (cb + u) * 4 + sub < cols_valid))
                                bp_row[(cb + u) * 4] = (uint16_t)k0[u];
                        __syncwarp();
                    }
                }
#endif
                named_barrier(2, kRescanThreads);
                for (int idx = rtid; idx < TM * (TN / 8); idx += kRescanThreads) {
                    const int r = idx / (TN / 8), c8 = idx % (TN / 8);
                    if (t < slot_frames[r] && j0 + c8 * 8 < S)
                        *reinterpret_cast<uint4 *>(p.backptr + ((int64_t)slot_item[r] * p.T + t) * S + j0 + c8 * 8) =
                            *reinterpret_cast<const uint4 *>(bp_tile + r * TN + c8 * 8);
                }
                named_barrier(2, kRescanThreads);
                if (lane == 0) mbar_arrive(bars + kRescanDone0 + pb);
            }
            barriers += (unsigned)p.NJT * batch_frames;         // t = 0 exchange + one per step
            __syncthreads();
        }
      }
    }
}

}  // namespace

bool forward_tiled_ws_supported(const ForwardArgs &a) {
    return forward_tiled_supported(a) && a.tile_rows == 64 && a.ws.score[3] != nullptr &&
           a.S % 8 == 0 && (reinterpret_cast<uintptr_t>(a.observation) & 15) == 0;
}

template <uint32_t kFlags>
static cudaError_t launch_ws_t(const ForwardArgs &a, cudaStream_t stream) {
    WSParams p{};
    p.observation = a.observation;
    p.transition = a.ws.transition_pad;
    p.initial = a.initial;
    p.frames = a.ws.frames;
    p.order = a.ws.order;
    p.score0 = a.ws.score[0];
    p.score_stride = a.ws.score[1] - a.ws.score[0];
    p.backptr = a.ws.backptr;
    p.posterior = a.posterior;
    p.final_state = a.ws.final_state;
    p.sync = a.ws.sync;
    p.frames_ready = a.frames_ready;
    p.T = a.T;
    p.S = (int)a.S;
    p.LDP = (int)padded_sources(a.S);
    p.NJT = (int)(padded_dests(a.S) / TN);
    p.num_batches = (int)((a.B + TM - 1) / TM);
    p.NG = std::min(a.sm_count / p.NJT, p.num_batches);
    if ((p.NG + 1) * 32 > (int)kBandWordOffset) return cudaErrorInvalidValue;
    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    auto kernel = forward_tiled_ws_kernel<kFlags>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemWS);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&p};
    return cudaLaunchCooperativeKernel((void *)kernel, dim3(p.NG * p.NJT), dim3(kThreadsWS), args,
                                       kSmemWS, stream);
}

cudaError_t launch_forward_tiled_ws(const ForwardArgs &a, cudaStream_t stream) {
    switch (a.flags & 3u) {
        case 0: return launch_ws_t<0>(a, stream);
        case 1: return launch_ws_t<1>(a, stream);
        case 2: return launch_ws_t<2>(a, stream);
        default: return launch_ws_t<3>(a, stream);
    }
}

}  // namespace torbi_b200
