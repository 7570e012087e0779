// Launcher interface between api.cu and the kernel translation units.
#pragma once

#include "common.cuh"

namespace torbi_b200 {

struct ForwardArgs {
    const float *observation;  // device (B,T,S); fp32, or 2-byte elements when obs_dtype != 0
    int obs_dtype;             // 0 fp32, 1 fp16, 2 bf16 (TORBI_B200_OBS_F16 / _BF16)
    const float *transition;   // device (S,S) [dest][src]
    const float *initial;      // device (S,)
    float *posterior;          // device (B,S) or nullptr
    // device word holding how many leading frames of `observation` are valid so far
    // (the host entry streams the batch in while the recursion runs); nullptr = all
    const int32_t *frames_ready;
    int64_t B, T, S;
    int64_t max_frames;        // max(batch_frames)
    uint32_t flags;
    int sm_count;
    int tile_rows;             // batch height of the tiled kernel: 64, 32 or 16
    // torbi_b200_decode_checkpointed: device (B, (T - 1) / checkpoint_every, S) receiving P_t for
    // t = k * checkpoint_every; nullptr / 0 otherwise
    float *checkpoints;
    int64_t checkpoint_every;
    Workspace ws;
};

// forward_generic.cu -- one CTA per sequence, any S
cudaError_t launch_forward_generic(const ForwardArgs &a, cudaStream_t stream);

// forward_tiled.cu -- batch-tiled max-plus "GEMM" per timestep, persistent grid
bool forward_tiled_supported(const ForwardArgs &a);
cudaError_t launch_forward_tiled(const ForwardArgs &a, cudaStream_t stream);
// forward_tiled_hist.cu -- the same kernel in score-history mode (no argmax bookkeeping; the
// workspace's backpointer region holds fp32 (B,T,S) forward scores for launch_backtrace_recompute)
cudaError_t launch_forward_tiled_hist(const ForwardArgs &a, cudaStream_t stream);
int tiled_choose_rows(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count,
                      double *seconds);
cudaError_t launch_pad_transition(const ForwardArgs &a, cudaStream_t stream);

// forward_split.cu -- split-K small tiles (32 x 20 with the transition block resident in shared memory,
// or 16 x 40 streamed) for batches too small to fill the chip with 64 x 80 tiles (score-history mode only)
bool forward_split_supported(const ForwardArgs &a);
double split_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count);
cudaError_t launch_forward_split(const ForwardArgs &a, cudaStream_t stream);

// forward_cluster.cu -- split-K over clusters of 2 CTAs (DSMEM partial-maximum exchange), batches of a
// group interleaved so one batch's L2 score exchange hides behind the other's K loop (score-history mode)
bool forward_cluster_supported(const ForwardArgs &a);
double cluster_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count);
cudaError_t launch_forward_cluster(const ForwardArgs &a, cudaStream_t stream);

// forward_wide.cu -- all SMs on every timestep of a few sequences (small batches, long sequences)
bool forward_wide_supported(const ForwardArgs &a);
double wide_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count);
cudaError_t launch_forward_wide(const ForwardArgs &a, cudaStream_t stream);

// forward_flow.cu -- one sequence at a time on all SMs, timesteps exchanged as tagged 8-byte words
// through L2 (no grid barrier), transition block resident in registers + shared memory
bool forward_flow_supported(const ForwardArgs &a);
double flow_estimate_seconds(const int32_t *frames_sorted, int64_t B, int64_t S, int sm_count);
cudaError_t launch_forward_flow(const ForwardArgs &a, cudaStream_t stream);

// backtrace.cu -- ragged backtrace + tail fill (viterbi.cpp:140-160, :219-221)
cudaError_t launch_backtrace(const Workspace &ws, int32_t *indices, int64_t B, int64_t T,
                             int64_t S, cudaStream_t stream);

// long sequences with stored backpointers (config 5): blocked backtrace -- per block of frames the
// composition of the backpointer maps for ALL states (parallel over blocks), a short serial pass over
// the block boundaries, then every block chases its own segment (parallel); bit-identical
constexpr int64_t kWindowedBacktraceMinFrames = 16384;
cudaError_t launch_backtrace_blocked(const Workspace &ws, int32_t *indices, int64_t B, int64_t T, int64_t S,
                                     int64_t max_frames, int *launches, cudaStream_t stream);

// score-history kernels + torbi_b200_decode_checkpointed: copy the requested rows of the history
cudaError_t launch_gather_checkpoints(const ForwardArgs &a, cudaStream_t stream);

// backtrace by recomputation from the fp32 score history (tiled kernel, score-history mode)
cudaError_t launch_backtrace_recompute(const Workspace &ws, int32_t *indices, int64_t B, int64_t T,
                                       int64_t S, cudaStream_t stream);

// entropy.cu -- framewise normalised entropy of chunk.py:81-85; out = device (rows, 2) fp32:
// [H, sum of |terms| / log(states)]
cudaError_t launch_entropy(const void *observation, int obs_dtype, int64_t rows, int64_t S, float *out,
                           int sm_count, cudaStream_t stream);

// selftest.cu
cudaError_t launch_selftest_conversion(unsigned long long *mismatches, int sm_count, cudaStream_t stream);

}  // namespace torbi_b200
