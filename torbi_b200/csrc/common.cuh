// Shared device/host helpers for the torbi_b200 kernels (sm_100a only).
#pragma once

#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/torbi_b200.h"

namespace torbi_b200 {

constexpr int kWarp = 32;
constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------------------
// Fused log-space conversion of one observation element.
// Restates /root/reference/torbi/core.py:189-197:
//     if not log_probs: observation = torch.log(observation)
//     torch.exp_(observation); observation += finfo(float32).tiny; torch.log_(observation)
// with the CUDA math library's expf/logf (what torch's CUDA exp/log kernels
// call), no fast-math, no flush-to-zero.
//
// eps_floor(): logf(expf(x) + FLT_MIN) is the identity, bit for bit, for every fp32 x in
// (-64, -4] -- where log-probabilities of a few thousand states live -- so the two
// transcendental calls are skipped there.  This is not an approximation: the range was found
// by evaluating all 2^32 inputs on the device (tools/probe_conversion.cu,
// profiles/r01_conversion_identity_probe.txt) and torbi_b200_selftest_conversion re-checks
// eps_floor against the plain formula for all 2^32 inputs (tests/test_gpu_parity.py).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float eps_floor_reference(float x) { return logf(expf(x) + FLT_MIN); }

__device__ __forceinline__ float eps_floor(float x) {
    if (x <= -4.f && x > -64.f) return x;
    return eps_floor_reference(x);
}

template <uint32_t kFlags>
__device__ __forceinline__ float convert_observation(float x) {
    if constexpr (kFlags & TORBI_B200_OBS_IS_PROBABILITY) x = logf(x);
    if constexpr (kFlags & TORBI_B200_EPS_FLOOR) x = eps_floor(x);
    return x;
}

__device__ __forceinline__ float convert_observation_rt(float x, uint32_t flags) {
    if (flags & TORBI_B200_OBS_IS_PROBABILITY) x = logf(x);
    if (flags & TORBI_B200_EPS_FLOOR) x = eps_floor(x);
    return x;
}

// ---------------------------------------------------------------------------
// Observation loads.  `observation` is fp32, or fp16 / bf16 with TORBI_B200_OBS_F16 / _BF16 (dtype 1 / 2):
// the reference casts whatever dtype it is given to fp32 on the host (core.py:191); here the 2-byte
// elements are upcast at the load -- exact, every half / bfloat16 value is an fp32 value.
// All loads are L2-coherent (.cg): the host entry may still be copying later frames of the same buffer
// while the kernel runs, and a non-coherent (.nc) or L1-cached load could return a stale sector that
// straddles a frame boundary.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float half_bits_to_float(unsigned short raw, int dtype) {
    return dtype == 1 ? __half2float(__ushort_as_half(raw)) : __uint_as_float((unsigned)raw << 16);
}
__device__ __forceinline__ float load_observation(const void *base, int64_t index, int dtype) {
    if (dtype == 0) return __ldcg(static_cast<const float *>(base) + index);
    return half_bits_to_float(__ldcg(static_cast<const unsigned short *>(base) + index), dtype);
}

// Emission tile of one timestep without cp.async (2-byte observations, S % 4 != 0 or an unaligned
// base): tile[r][c] = raw observation of slot r, state j0 + c, for live slots; plain shared-memory
// stores, the caller synchronises before the tile is consumed.
template <int TM, int TN, int kThreads>
__device__ __forceinline__ void load_emission_tile_plain(float *tile, const void *observation, int dtype,
                                                         const int *slot_item, const int *slot_frames,
                                                         const int64_t *slot_row, int S, int t, int j0) {
    const bool half4 = dtype != 0 && (S % 4) == 0 && (reinterpret_cast<uintptr_t>(observation) & 7) == 0;
    if (half4) {
        // four 2-byte elements per 8-byte load
        for (int idx = threadIdx.x; idx < TM * (TN / 4); idx += kThreads) {
            const int r = idx / (TN / 4), c4 = idx % (TN / 4);
            const int item = slot_item[r];
            if (item >= 0 && t < slot_frames[r] && j0 + c4 * 4 < S) {
                const uint2 raw = __ldcg(reinterpret_cast<const uint2 *>(
                    static_cast<const unsigned short *>(observation) + (slot_row[r] + t) * S + j0 + c4 * 4));
                float4 v;
                v.x = half_bits_to_float((unsigned short)(raw.x & 0xffffu), dtype);
                v.y = half_bits_to_float((unsigned short)(raw.x >> 16), dtype);
                v.z = half_bits_to_float((unsigned short)(raw.y & 0xffffu), dtype);
                v.w = half_bits_to_float((unsigned short)(raw.y >> 16), dtype);
                *reinterpret_cast<float4 *>(tile + r * TN + c4 * 4) = v;
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < TM * TN; idx += kThreads) {
            const int r = idx / TN, c = idx % TN;
            const int item = slot_item[r];
            if (item >= 0 && t < slot_frames[r] && j0 + c < S)
                tile[idx] = load_observation(observation, (slot_row[r] + t) * S + j0 + c, dtype);
        }
    }
}

// (value, index) ordering of the reference's strict-`>` ascending scan
// (/root/reference/torbi/csrc/viterbi.cpp:94-101): larger value wins, equal
// values keep the LOWER source index.
__device__ __forceinline__ bool better(float v, int i, float best, int best_i) {
    return v > best || (v == best && i < best_i);
}

// Workspace layout shared by api.cu and the kernels.  All offsets 256-B aligned.
struct Workspace {
    int32_t *frames;        // (B,)  device copy of batch_frames
    int32_t *order;         // (B,)  items sorted by length, longest first
    int32_t *final_state;   // (B,)  argmax of the final forward scores
    int64_t *item_row;      // (B,)  first row of each item in `observation` and `indices`:
                            //       b * T in the padded (B,T,S) layout, the caller's offsets in the packed one
    int64_t *hist_row;      // (B,)  first row of each item in the score history / backpointers, which are always
                            //       compact: sum(batch_frames) rows, item after item
    uint32_t *sync;         // scheduling / barrier words (zeroed per call)
    float *score[4];        // rotating forward-score buffers, (pad_batch, padded_sources) each
    float *block_max;       // 3 x (pad_batch, padded_sources / 16): max of each 16-state block of the rotating
                            // score rows (banded mode: the outside-window maximum comes from these)
    float *transition_pad;  // (padded_dests, padded_sources) copy, -inf outside (S,S)
    uint16_t *backptr;      // (B, T, S) uint16 backpointers, row 0 unused -- or, in score-history mode,
                            // the same region as fp32 (B, T, S) forward scores (4 bytes per entry)
    size_t bytes;
};

// scheduling words, zeroed per call: [0, 6144) one 128-byte line per CTA group (<= 148 groups + queue
// head), [6144, 8192) the banded-mode analysis (2 words per destination tile + the outside value)
constexpr size_t kSyncWords = 8192;
constexpr size_t kBandWordOffset = 6144;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Tiling constants of forward_tiled.cu that fix the padded layouts.
constexpr int kTileDests = 80;     // destination states per CTA
constexpr int kStageSources = 96;  // source states per pipeline stage

// row length of the score buffers and of the padded transition copy
inline int64_t padded_sources(int64_t S) { return (S + kStageSources - 1) / kStageSources * kStageSources; }
// rows of the padded transition copy
inline int64_t padded_dests(int64_t S) { return (S + kTileDests - 1) / kTileDests * kTileDests; }
inline int64_t pad_batch(int64_t B) { return (B + 127) / 128 * 128; }

// `history`: size the backpointer region for fp32 scores (unless TORBI_B200_BACKPOINTERS is set)
inline bool workspace_has_history(uint32_t flags) { return !(flags & TORBI_B200_BACKPOINTERS); }

// `history` = the trailing region holds fp32 forward scores (4 bytes per (b,t,s): tiled kernel in
// score-history mode, split / cluster kernels) rather than uint16 backpointers (2 bytes: wide and generic
// kernels, tiled kernel in backpointer mode).  Everything before that region has the same layout
// either way, so the region can be sized after the kernel has been chosen.  `rows` = rows of that
// region: sum(batch_frames) (B * T as the upper bound when the lengths are not known yet).
inline Workspace carve_workspace(void *base, int64_t B, int64_t rows, int64_t S, bool history) {
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = base ? static_cast<char *>(base) + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    const int64_t Sp = padded_sources(S), Bp = pad_batch(B);
    w.frames = static_cast<int32_t *>(take(sizeof(int32_t) * B));
    w.order = static_cast<int32_t *>(take(sizeof(int32_t) * Bp));
    w.final_state = static_cast<int32_t *>(take(sizeof(int32_t) * B));
    w.item_row = static_cast<int64_t *>(take(sizeof(int64_t) * B));
    w.hist_row = static_cast<int64_t *>(take(sizeof(int64_t) * B));
    w.sync = static_cast<uint32_t *>(take(sizeof(uint32_t) * kSyncWords));
    w.score[0] = static_cast<float *>(take(sizeof(float) * Bp * Sp));
    w.score[1] = static_cast<float *>(take(sizeof(float) * Bp * Sp));
    w.score[2] = static_cast<float *>(take(sizeof(float) * Bp * Sp));
    w.score[3] = static_cast<float *>(take(sizeof(float) * Bp * Sp));
    w.block_max = static_cast<float *>(take(sizeof(float) * 3 * Bp * (Sp / 16)));
    w.transition_pad = static_cast<float *>(take(sizeof(float) * padded_dests(S) * Sp));
    w.backptr = static_cast<uint16_t *>(
        take((history ? sizeof(float) : sizeof(uint16_t)) * rows * S));
    w.bytes = off;
    return w;
}

}  // namespace torbi_b200
