// "Wide" forward kernel for small batches and single long sequences: instead of
// giving each group of SMs its own batch (forward_tiled.cu), ALL SMs cooperate on
// every timestep of a few sequences.
//
//   * persistent cooperative grid of NC <= #SM CTAs; CTA c owns the destination
//     states [c*R, c*R+R) and keeps as many rows of its transition slice as fit
//     RESIDENT in shared memory for the whole call (S = 1440: all 10 rows, 58 KB;
//     S = 4096: 11 of 28 rows), streaming the rest from L2 every step;
//   * up to BC <= 16 sequences advance in lockstep; each step every CTA pulls their
//     previous scores (BC x S floats) from L2 into shared memory, one warp computes
//     a destination row for all BC sequences, and the grid meets at a counter
//     barrier -- the posterior vector is exchanged through L2 once per timestep;
//   * more than BC sequences are processed group after group, longest first.
//
// Same exactness rules as the tiled kernel: candidates are the single fp32 add
// P[s] + A[j][s]; each lane keeps a running maximum over its (ascending) float4
// groups and the FIRST group that strictly raised it, then resolves the lowest
// index inside that group by equality, and the warp reduces on (value desc,
// index asc) -- the reference's strict-`>` ascending scan (viterbi.cpp:94-101).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

namespace {

constexpr int kWideThreads = 512;
constexpr int kWideWarps = kWideThreads / kWarp;

struct WideParams {
    const float *observation;   // (rows,S), item b at rows item_row[b] ..; 2-byte elements when obs_dtype != 0
    int obs_dtype;              // 0 fp32, 1 fp16, 2 bf16
    float *checkpoints;         // (B, ck_rows, S) or null: P_t for t = k * checkpoint_every
    int64_t checkpoint_every, ck_rows;
    const float *transition;    // padded copy (SJ, LDP), -inf outside (S,S)
    const float *initial;       // (S,)
    const int32_t *frames;      // (B,)
    const int64_t *item_row;    // (B,) first row of each item in observation
    const int64_t *hist_row;    // (B,) first row of each item in the (compact) backpointers
    const int32_t *order;       // (Bpad,) items longest first
    float *score0;              // 2 alternating buffers (Bpad, LDP), `score_stride` floats apart
    int64_t score_stride;
    uint16_t *backptr;          // (rows,S)
    float *posterior;           // (B,S) or null
    int32_t *final_state;       // (B,)
    unsigned *sync;             // grid barrier counter
    const int32_t *frames_ready;
    int B, S, LDP, NC, R;
    int resident_rows;          // leading rows of each CTA's transition slice kept in shared memory
};

__device__ __forceinline__ unsigned wide_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int wide_ld_acquire_sys(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (wide_ld_acquire(counter) < target) {}
    }
    __syncthreads();
}

template <int BC, uint32_t kFlags>
__global__ void __launch_bounds__(kWideThreads, 1) forward_wide_kernel(const WideParams p) {
    constexpr int BW = BC < 4 ? BC : 4;                  // sequences scanned together by one warp
    extern __shared__ __align__(16) float smem[];
    const int S = p.S, LDP = p.LDP, R = p.R;
    float *Ps = smem;                                    // [BC][LDP] previous scores
    float *As = smem + BC * LDP;                         // [resident_rows][LDP] transition slice
    __shared__ int slot_item[BC], slot_frames[BC];
    __shared__ int64_t slot_row[BC], slot_hrow[BC];
    __shared__ int ready_shared;

    const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
    const int j_begin = blockIdx.x * R, j_end = min(j_begin + R, S);
    unsigned barriers = 0;
    int ready = p.frames_ready ? 0 : 0x7fffffff;
    auto wait_for_frame = [&](int t) {
        if (t >= ready) {
            if (tid == 0) {
                int r;
                while ((r = wide_ld_acquire_sys(p.frames_ready)) <= t) __nanosleep(200);
                ready_shared = r;
            }
            __syncthreads();
            ready = ready_shared;
            __syncthreads();
        }
    };

    const int resident_rows = min(p.resident_rows, j_end - j_begin);
    {
        for (int idx = tid; idx < resident_rows * (LDP / 4); idx += kWideThreads) {
            const int r = idx / (LDP / 4), q = idx % (LDP / 4);
            reinterpret_cast<float4 *>(As + r * LDP)[q] =
                __ldg(reinterpret_cast<const float4 *>(p.transition + (int64_t)(j_begin + r) * LDP) + q);
        }
    }

    for (int first = 0; first < p.B; first += BC) {
        const int nb = min(BC, p.B - first);
        __syncthreads();
        if (tid < BC) {
            const int item = tid < nb ? p.order[first + tid] : -1;
            slot_item[tid] = item;
            slot_frames[tid] = item >= 0 ? p.frames[item] : 0;
            slot_row[tid] = item >= 0 ? p.item_row[item] : 0;
            slot_hrow[tid] = item >= 0 ? p.hist_row[item] : 0;
        }
        __syncthreads();
        const int group_frames = slot_frames[0];         // longest first

        // ---- t = 0 (viterbi.cpp:72-76); every CTA writes its own columns ----
        wait_for_frame(0);
        {
            const int c_end = (blockIdx.x == p.NC - 1) ? LDP : j_end;
            for (int b = warp; b < nb; b += kWideWarps) {
                const int item = slot_item[b];
                float *dst0 = p.score0 + (int64_t)(first + b) * LDP;
                const int64_t obs0 = slot_row[b] * S;
                for (int c = j_begin + lane; c < c_end; c += kWarp) {
                    if (c < S) {
                        dst0[c] = convert_observation<kFlags>(load_observation(p.observation, obs0 + c, p.obs_dtype)) + p.initial[c];
                    } else {
                        dst0[c] = 0.f;                   // padding stays finite in both buffers
                        dst0[c + p.score_stride] = 0.f;
                    }
                }
            }
        }
        barriers += p.NC;
        grid_barrier(p.sync, barriers);

        for (int t = 1; t < group_frames; ++t) {
            const float *P_cur = p.score0 + ((t - 1) & 1) * p.score_stride + (int64_t)first * LDP;
            float *P_new = p.score0 + (t & 1) * p.score_stride + (int64_t)first * LDP;
            wait_for_frame(t);

            // previous scores of the group -> shared memory (L2 only: rewritten every step)
#pragma unroll 4
            for (int idx = tid; idx < nb * (LDP / 4); idx += kWideThreads)
                reinterpret_cast<float4 *>(Ps)[idx] = __ldcg(reinterpret_cast<const float4 *>(P_cur) + idx);
            __syncthreads();

            // Work items = (destination row, sub-group of BW sequences), dealt round-robin to the
            // warps: with R = 10 rows and 16 sequences that is 40 items for 16 warps instead of
            // 10 busy warps scanning 16 sequences each.
            const int subgroups = (nb + BW - 1) / BW;
            const int items = (j_end - j_begin) * subgroups;
            for (int item_id = warp; item_id < items; item_id += kWideWarps) {
                const int j = j_begin + item_id / subgroups;
                const int b0 = (item_id % subgroups) * BW;
                const int nbw = min(BW, nb - b0);                 // sequences of this item
                // emissions (and carried scores of finished sequences) requested early by lane b -- RAW: the
                // conversion branches on the value, so converting here made the warp wait out the HBM latency
                // of the observation before it could start its scan (ncu source view: 6 % of the samples)
                float emit = 0.f;
                if (lane < nbw) {
                    emit = t < slot_frames[b0 + lane]
                        ? load_observation(p.observation, (slot_row[b0 + lane] + t) * S + j, p.obs_dtype)
                        : __ldcg(P_cur + (int64_t)(b0 + lane) * LDP + j);
                }
                float best[BW];
                int grp[BW];
                const float *prow[BW];
                const bool has_groups = lane < LDP / 4;          // S <= 96: lanes 24..31 own no source state
#pragma unroll
                for (int b = 0; b < BW; ++b) {
                    best[b] = -INFINITY;
                    grp[b] = has_groups ? lane : 0;
                    // slots past the item's last sequence re-read its row: no branch between the loads
                    prow[b] = Ps + (b0 + (b < nbw ? b : nbw - 1)) * LDP;
                }
                int arg_b[BW];
                // the scan, parametrised on where the transition row lives: resident rows are
                // read with LDS, the rest of the slice streams from L2
                // (issuing 8 loads per lane ahead of the scan on streamed rows was measured SLOWER: 10.8 vs
                // 8.4 us per step at S = 4096, profiles/r02_wide_kernel_notes.txt)
                auto scan_row = [&](auto load_a) {
                    for (int q = lane; q < LDP / 4; q += kWarp) {
                        const float4 a = load_a(q);
#pragma unroll
                        for (int b = 0; b < BW; ++b) {
                            const float4 pv = reinterpret_cast<const float4 *>(prow[b])[q];
                            const float m = fmaxf(fmaxf(pv.x + a.x, pv.y + a.y), fmaxf(pv.z + a.z, pv.w + a.w));
                            if (m > best[b]) { best[b] = m; grp[b] = q; }         // first group reaching the max
                        }
                    }
#pragma unroll
                    for (int b = 0; b < BW; ++b) {
                        // lowest index inside the lane's winning group
                        const float4 a = load_a(grp[b]);
                        const float4 pv = reinterpret_cast<const float4 *>(prow[b])[grp[b]];
                        int arg = 4 * grp[b] + 3;
                        if (pv.z + a.z == best[b]) arg = 4 * grp[b] + 2;
                        if (pv.y + a.y == best[b]) arg = 4 * grp[b] + 1;
                        if (pv.x + a.x == best[b]) arg = 4 * grp[b];
                        arg_b[b] = arg;
                    }
                };
                if (j - j_begin < resident_rows) {
                    const float4 *arow = reinterpret_cast<const float4 *>(As + (j - j_begin) * LDP);
                    scan_row([&](int q) { return arow[q]; });
                } else {
                    const float4 *arow = reinterpret_cast<const float4 *>(p.transition + (int64_t)j * LDP);
                    scan_row([&](int q) { return arow[q]; });   // plain LDG: measured 2x faster here than the .nc path
                }
#pragma unroll
                for (int b = 0; b < BW; ++b) {
                    if (b >= nbw) continue;
                    // ... then across lanes
                    int arg = arg_b[b];
                    if (!has_groups) arg = 0x7fffffff;            // (-inf, INT_MAX) never wins a tie
                    float v = best[b];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(kFullMask, v, off);
                        const int oi = __shfl_xor_sync(kFullMask, arg, off);
                        if (better(ov, oi, v, arg)) { v = ov; arg = oi; }
                    }
                    const float e_raw = __shfl_sync(kFullMask, emit, b);
                    if (lane == 0) {
                        const int bb = b0 + b;
                        const bool live = t < slot_frames[bb];
                        const float e = live ? convert_observation<kFlags>(e_raw) : e_raw;
                        P_new[(int64_t)bb * LDP + j] = live ? e + v : e;         // viterbi.cpp:102 / carried
                        if (live) {
                            p.backptr[(slot_hrow[bb] + t) * S + j] = (uint16_t)min(arg, S - 1);
                            if (p.checkpoints && t % p.checkpoint_every == 0)
                                p.checkpoints[((int64_t)slot_item[bb] * p.ck_rows + (t / p.checkpoint_every - 1)) * S + j] = e + v;
                        }
                    }
                }
            }
            barriers += p.NC;
            grid_barrier(p.sync, barriers);
        }

        // ---- final scores out (viterbi.cpp:111-113) + first-occurrence argmax (:218) ----
        const float *P_fin = p.score0 + ((group_frames - 1) & 1) * p.score_stride + (int64_t)first * LDP;
        for (int b = blockIdx.x * kWideWarps + warp; b < nb; b += p.NC * kWideWarps) {
            const int item = slot_item[b];
            const float *row = P_fin + (int64_t)b * LDP;
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

struct WideGeometry {
    int NC, R, BC;
    int resident_rows;     // of the R rows a CTA owns
    bool resident;         // all of them
    size_t smem;
};

WideGeometry wide_geometry(int64_t B, int64_t S, int sm_count) {
    WideGeometry g{};
    g.R = (int)((S + sm_count - 1) / sm_count);
    g.NC = (int)((S + g.R - 1) / g.R);
    const size_t ldp = (size_t)padded_sources(S);
    const size_t slice = sizeof(float) * g.R * ldp, budget = 226 * 1024;   // (227 KB per CTA minus the static arrays)
    // as many sequences in lockstep as the batch has and shared memory holds
    g.BC = B > 8 ? 16 : B > 4 ? 8 : B > 2 ? 4 : B > 1 ? 2 : 1;
    while (g.BC > 1 && sizeof(float) * g.BC * ldp > budget) g.BC /= 2;
    const size_t scores = sizeof(float) * g.BC * ldp;
    (void)slice;
    g.resident_rows = scores < budget ? (int)std::min<size_t>(g.R, (budget - scores) / (sizeof(float) * ldp)) : 0;
    g.resident = g.resident_rows == g.R;
    g.smem = scores + sizeof(float) * ldp * g.resident_rows;
    return g;
}

}  // namespace

bool forward_wide_supported(const ForwardArgs &a) {
    const WideGeometry g = wide_geometry(a.B, a.S, a.sm_count);
    return g.smem <= 226 * 1024 && a.ws.transition_pad != nullptr;
}

// Estimated seconds for the wide kernel: per step a fixed exchange + barrier cost
// plus the max-plus work of the sequences still running (calibrated on the B200).
double wide_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count) {
    // B200 measurements (tools/calibrate_kernels.sh): S = 1440 fully resident: 3.2 / 5.3 / 13.9 us
    // per step for 1 / 4 / 16 sequences; S = 4096 (11 of 28 rows resident): 8.5 / 22 us for 1 / 4.
    const WideGeometry g = wide_geometry(B, S, sm_count);
    const double ldp = (double)padded_sources(S);
    const double rounds = (g.R + kWideWarps - 1) / kWideWarps;             // rows per warp
    const double per_sequence_us = rounds * ldp * 0.0005;
    const double streamed = 1.0 - (double)g.resident_rows / g.R;
    const double slice_stream_us = 4.0 * (double)padded_dests(S) * ldp * streamed / 5.0e6;   // ~5 TB/s of L2
    double total = 0.0;
    for (int64_t first = 0; first < B; first += g.BC) {
        const int nb = (int)std::min<int64_t>(g.BC, B - first);
        // streaming the non-resident rows overlaps the per-sequence scans
        // (fixed part: grid barrier 1.25 us + score reload + reduction latency; 2.45 us measured at S = 256, batch 1)
        total += frames_sorted[first] * (2.3 + std::max(slice_stream_us, per_sequence_us * nb));
    }
    return total * 1e-6;
}

template <int BC>
static cudaError_t launch_wide_t(const ForwardArgs &a, const WideParams &p, size_t smem,
                                 cudaStream_t stream) {
    void *args[] = {(void *)&p};
    auto launch = [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaLaunchCooperativeKernel((void *)kernel, dim3(p.NC), dim3(kWideThreads), args, smem, stream);
    };
    switch (a.flags & 3u) {
        case 0: return launch(forward_wide_kernel<BC, 0>);
        case 1: return launch(forward_wide_kernel<BC, 1>);
        case 2: return launch(forward_wide_kernel<BC, 2>);
        default: return launch(forward_wide_kernel<BC, 3>);
    }
}

cudaError_t launch_forward_wide(const ForwardArgs &a, cudaStream_t stream) {
    const WideGeometry g = wide_geometry(a.B, a.S, a.sm_count);
    WideParams p{};
    p.observation = a.observation;
    p.obs_dtype = a.obs_dtype;
    p.checkpoints = a.checkpoints;
    p.checkpoint_every = a.checkpoint_every > 0 ? a.checkpoint_every : 1;
    p.ck_rows = a.checkpoints ? (a.T - 1) / a.checkpoint_every : 0;
    p.transition = a.ws.transition_pad;
    p.initial = a.initial;
    p.frames = a.ws.frames;
    p.item_row = a.ws.item_row;
    p.hist_row = a.ws.hist_row;
    p.order = a.ws.order;
    p.score0 = a.ws.score[0];
    p.score_stride = a.ws.score[1] - a.ws.score[0];
    p.backptr = a.ws.backptr;
    p.posterior = a.posterior;
    p.final_state = a.ws.final_state;
    p.sync = a.ws.sync;
    p.frames_ready = a.frames_ready;
    p.B = (int)a.B;
    p.S = (int)a.S;
    p.LDP = (int)padded_sources(a.S);
    p.NC = g.NC;
    p.R = g.R;
    p.resident_rows = g.resident_rows;
    cudaError_t e = launch_pad_transition(a, stream);
    if (e != cudaSuccess) return e;
    switch (g.BC) {
        case 1: return launch_wide_t<1>(a, p, g.smem, stream);
        case 2: return launch_wide_t<2>(a, p, g.smem, stream);
        case 4: return launch_wide_t<4>(a, p, g.smem, stream);
        case 8: return launch_wide_t<8>(a, p, g.smem, stream);
        default: return launch_wide_t<16>(a, p, g.smem, stream);
    }
}

}  // namespace torbi_b200
