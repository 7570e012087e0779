/*
 * torbi_b200 -- C ABI of the B200-native Viterbi decoder.
 *
 * This is the drop-in boundary for torbi's one hot path.  It replaces the
 * PyTorch custom operator
 *
 *     torbi::viterbi_decode(Tensor observation, Tensor batch_frames,
 *                           Tensor transition, Tensor initial) -> Tensor
 *
 * declared at /root/reference/torbi/csrc/ops.cpp:16-18 and implemented by
 * viterbi_decode_cpu (/root/reference/torbi/csrc/viterbi.cpp:182-234) and
 * viterbi_decode_cuda (/root/reference/torbi/csrc/cuda/viterbi.cu:309-362),
 * whose only caller is torbi.decode (/root/reference/torbi/viterbi.py:53).
 * The bindings a maintainer adds on the reference side are in INTEGRATION.md.
 *
 * Plain C: pointers and sizes only, no torch / ATen types.  Results follow the
 * reference CPU operator exactly: decoded indices bit-exact, fp32 forward
 * scores bit-exact (same add / max order), lowest source index wins ties
 * (viterbi.cpp:94-101), first-occurrence final argmax (viterbi.cpp:218),
 * positions t >= batch_frames[b]-1 hold the final state (viterbi.cpp:219-221).
 *
 * There is no CPU fallback: every entry point needs a CUDA device of compute
 * capability 10.x and fails with TORBI_B200_ERR_CUDA otherwise.
 */
#ifndef TORBI_B200_H_
#define TORBI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TORBI_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define TORBI_B200_API __attribute__((visibility("default")))
#else
#define TORBI_B200_API
#endif

/* status codes (0 = success, negative = failure; text via torbi_b200_last_error) */
#define TORBI_B200_OK 0
#define TORBI_B200_ERR_INVALID_ARGUMENT (-1) /* shapes, NULLs, frames outside [1,T], S > 65535 */
#define TORBI_B200_ERR_WORKSPACE (-2)        /* workspace too small or misaligned            */
#define TORBI_B200_ERR_CUDA (-3)             /* launch / runtime failure, no sm_100 device   */
#define TORBI_B200_ERR_OUT_OF_MEMORY (-4)    /* host-entry arena could not be allocated      */

/* flags: the log-space conversion torbi.from_probabilities applies on the host
 * (/root/reference/torbi/core.py:189-197), fused into the forward kernel's
 * observation load instead of three extra passes over (B,T,S). */
#define TORBI_B200_OBS_IS_PROBABILITY (1u << 0) /* x <- logf(x) first (core.py:189-190)        */
#define TORBI_B200_EPS_FLOOR (1u << 1)          /* x <- logf(expf(x) + FLT_MIN) (core.py:193-197) */
/* element type of `observation`: default fp32; with one of these it points at IEEE half / bfloat16
 * elements (2 bytes each, same (batch, frames, states) layout).  The reference casts any input dtype to
 * fp32 on the host before the copy (core.py:191); here half-precision inputs cross PCIe / HBM at 2 bytes
 * per element and the kernels upcast in their observation load -- exact, the decode runs in fp32. */
#define TORBI_B200_OBS_F16 (1u << 4)
#define TORBI_B200_OBS_BF16 (1u << 5)
/* kernel selection, for tests and profiling; default 0 lets the library pick */
#define TORBI_B200_FORCE_GENERIC (1u << 8)      /* one CTA per sequence, any S                 */
#define TORBI_B200_FORCE_TILED (1u << 9)        /* batch-tiled max-plus kernel                 */
#define TORBI_B200_FORCE_WIDE (1u << 11)        /* all SMs per timestep (small batch / long sequence) */
/* record CUDA events around the forward and backtrace kernels on the caller's
 * stream so torbi_b200_last_timing can report their device time */
#define TORBI_B200_TIME_KERNELS (1u << 10)
/* force the batch height of the tiled kernel (tests / tuning): bits 12-13 = 1, 2, 3 for 16, 32, 64 rows */
#define TORBI_B200_TILE_ROWS_MASK (3u << 12)
#define TORBI_B200_TILE_ROWS_16 (1u << 12)
#define TORBI_B200_TILE_ROWS_32 (2u << 12)
#define TORBI_B200_TILE_ROWS_64 (3u << 12)
/* exact fast path for banded transition matrices (tiled kernel): a device-side pass finds,
 * per tile of destination states, the source window outside which every entry equals one
 * constant; the dense loop runs over the window only and the rest is a first-occurrence max
 * of P[s] + constant.  Bit-identical results for ANY matrix, faster when it is banded. */
#define TORBI_B200_BANDED (1u << 15)
/* The tiled kernel has two ways of producing the path.  Default (score-history mode): the forward
 * pass is a pure max-plus product that keeps every fp32 score vector P_t (4 bytes per (b,t,s) in
 * the workspace) and the backtrace recomputes, per step, the one argmax that lies on the decoded
 * path -- idx_{t-1} = lowest s maximising P_{t-1}[s] + A[idx_t][s], the forward pass's own fp32 add,
 * hence the same result bit for bit.  With this flag the kernel tracks the argmax of every state
 * and stores uint16 backpointers (2 bytes per (b,t,s)) like the reference's trellis
 * (viterbi.cpp:91-104); slower, half the workspace.  Pass the same flag to
 * torbi_b200_workspace_bytes.  The wide, flow and generic kernels always store backpointers. */
#define TORBI_B200_BACKPOINTERS (1u << 16)
/* force the split-K small-tile kernel (32 sequences x 20 states per CTA with the transition block
 * resident in shared memory, or 16 x 40 streamed; score-history mode): the decomposition for batches
 * too small to fill the chip with 64 x 80 tiles */
#define TORBI_B200_FORCE_SPLIT (1u << 17)
/* with the split kernel: stream the transition block (16 x 40 tiles) even where the resident
 * 32 x 20 shape would fit (states <= 1536); tests / tuning */
#define TORBI_B200_SPLIT_STREAMED (1u << 18)
/* force the cluster split-K kernel: pairs of CTAs (thread-block clusters of 2) share a 32-sequence x
 * 20-state tile and split its source states 16 ways (8 warps x 2 CTAs); the partial maxima meet through
 * distributed shared memory (st.shared::cluster + remote mbarrier arrive) and every CTA pair alternates
 * between the batches of its group, so one batch's score exchange through L2 is hidden behind the
 * other's max-plus loop.  Score-history mode, dense transitions, states <= 20 x (#SM / 2). */
#define TORBI_B200_FORCE_CLUSTER (1u << 19)
/* force the flow kernel: one sequence at a time on all SMs, every timestep exchanged as 8-byte
 * {score, step tag} words through L2 (single 64-bit stores polled by their consumers: no grid barrier,
 * no atomics), the transition block resident in registers + shared memory.  The decomposition for one
 * or a few sequences (config 1, config 5); states <= 64 x #SM.  Stores uint16 backpointers. */
#define TORBI_B200_FORCE_FLOW (1u << 20)
/* (bit 14 was the warp-specialised variant of the tiled kernel: measured slower, no longer built;
 * source and experiment log kept under tools/experiments/ and profiles/r01_ws_experiment.txt) */

TORBI_B200_API int torbi_b200_abi_version(void);

/* Thread-local description of the last failure on the calling thread. */
TORBI_B200_API const char *torbi_b200_last_error(void);

/* Number of CUDA devices usable by this library (compute capability 10.x). */
TORBI_B200_API int torbi_b200_device_count(void);

/* Upper bound of the device workspace torbi_b200_decode needs for this shape, whatever the lengths
 * (fp32 score history B*T*S*4 -- or uint16 backpointers B*T*S*2 with
 * TORBI_B200_BACKPOINTERS --, rotating forward scores, scheduling state).  The history / backpointers
 * only ever hold the frames that exist (sum of batch_frames rows): torbi_b200_workspace_bytes_planned
 * gives the exact figure for one call.  Returns 0 on invalid shape. */
TORBI_B200_API size_t torbi_b200_workspace_bytes(int64_t batch, int64_t frames, int64_t states, uint32_t flags);

/* The same for one concrete call: the library picks the forward kernel from the length profile first
 * (as torbi_b200_plan does) and sizes the trailing region for it -- uint16 backpointers when the wide or
 * generic kernel will run (small batches, one long sequence: half of the figure above), fp32 score
 * history otherwise.  torbi_b200_decode accepts a workspace of this size for the same arguments and the
 * same `sm_count` (cudaDevAttrMultiProcessorCount of the device it runs on).  Host logic only; 0 on
 * invalid arguments. */
TORBI_B200_API size_t torbi_b200_workspace_bytes_planned(const int32_t *batch_frames, int64_t batch,
                                                         int64_t frames, int64_t states, uint32_t flags,
                                                         int sm_count);

/*
 * Decode a batch resident in device memory.  Replaces viterbi_decode_cuda
 * (viterbi.cu:309-362); argument meaning follows viterbi.py:5-30.
 *
 *   observation   device (batch, frames, states) fp32 (fp16 / bf16 with TORBI_B200_OBS_F16 /
 *                 _BF16), contiguous.  Log probabilities, or whatever `flags` says.
 *                 Never written.
 *   batch_frames  HOST   (batch,) int32, each in [1, frames]  (the reference
 *                 trusts these, viterbi.cpp:66; here they are validated)
 *   transition    device (states, states) fp32 log-probs, [destination][source]
 *                 (viterbi.cpp:83-84)
 *   initial       device (states,) fp32 log-probs
 *   indices       device (batch, frames) int32, output
 *   posterior     device (batch, states) fp32 final forward scores, or NULL
 *                 (the reference keeps these internal, viterbi.cpp:204-206)
 *   workspace     device, >= torbi_b200_workspace_bytes(...), 256-byte aligned;
 *                 owned by the caller, contents undefined afterwards
 *   stream        cudaStream_t (NULL = legacy default stream).  All work is
 *                 enqueued on it, on the CURRENT device; the call does not
 *                 synchronise.
 *
 * The library never allocates or frees device memory in this entry point and
 * is re-entrant across host threads (one per device).
 */
TORBI_B200_API int torbi_b200_decode(const void *observation, const int32_t *batch_frames,
                      const float *transition, const float *initial, int64_t batch,
                      int64_t frames, int64_t states, uint32_t flags, int32_t *indices,
                      float *posterior, void *workspace, size_t workspace_bytes, void *stream);

/*
 * torbi_b200_decode that also hands out intermediate forward scores: `checkpoints` is a device array
 * (batch, (frames - 1) / checkpoint_every, states) fp32 receiving P_t (viterbi.cpp:102) for
 * t = checkpoint_every, 2 checkpoint_every, ... of every item (rows with t >= batch_frames[b] are left
 * untouched).  The reference keeps these internal; they let a long decode (config 5: 4096 states x 10^6
 * frames, which the reference cannot index, viterbi.cpp:99) be verified segment by segment: a CPU
 * restatement started from checkpoint k must reproduce checkpoint k+1 and the decoded path in between.
 */
TORBI_B200_API int torbi_b200_decode_checkpointed(const void *observation, const int32_t *batch_frames,
                      const float *transition, const float *initial, int64_t batch,
                      int64_t frames, int64_t states, uint32_t flags, int32_t *indices,
                      float *posterior, void *workspace, size_t workspace_bytes, void *stream,
                      float *checkpoints, int64_t checkpoint_every);

/*
 * The same decode on a PACKED (unpadded) batch -- the ragged layout of SURVEY 8b.  The reference pads
 * every batch to its longest item (data/collate.py:26-31) and allocates a (batch, frames, states)
 * trellis (viterbi.cpp:201-203); here item b occupies rows [item_row[b], item_row[b] + batch_frames[b])
 * of one (rows, states) observation buffer, its decoded states land in the same rows of the (rows,)
 * `indices`, and the score history / backpointers take rows * states entries instead of
 * batch * max_frames * states.  Items may be any non-overlapping row ranges in any order (gaps are
 * allowed: rows no item covers are neither read nor written), so the pieces of a chunked sequence
 * (chunk.py:12-49) or the files of a batch are decoded where they lie, without a padded copy.
 *
 *   observation   device (rows, states) fp32 (fp16 / bf16 with TORBI_B200_OBS_F16 / _BF16)
 *   item_row      HOST   (batch,) int64 first row of every item
 *   batch_frames  HOST   (batch,) int32 rows of every item, each >= 1
 *   indices       device (rows,) int32: indices[item_row[b] + t] = state of item b at frame t
 *   workspace     device, >= torbi_b200_workspace_bytes_packed(...) for the same arguments
 * Everything else as torbi_b200_decode.  Each item decodes exactly as it would alone.
 */
TORBI_B200_API int torbi_b200_decode_packed(const void *observation, const int64_t *item_row,
                      const int32_t *batch_frames, const float *transition, const float *initial,
                      int64_t batch, int64_t rows, int64_t states, uint32_t flags, int32_t *indices,
                      float *posterior, void *workspace, size_t workspace_bytes, void *stream);

/* Workspace for torbi_b200_decode_packed (host logic only; 0 on invalid arguments, see
 * torbi_b200_last_error). */
TORBI_B200_API size_t torbi_b200_workspace_bytes_packed(const int64_t *item_row, const int32_t *batch_frames,
                                                        int64_t batch, int64_t rows, int64_t states,
                                                        uint32_t flags, int sm_count);

/* torbi_b200_decode_packed for HOST buffers on one device (copies, decode and the copy back of the
 * (rows,) indices inside the call; rows no item covers keep the caller's values). */
TORBI_B200_API int torbi_b200_decode_host_packed(const void *observation, const int64_t *item_row,
                      const int32_t *batch_frames, const float *transition, const float *initial,
                      int64_t batch, int64_t rows, int64_t states, uint32_t flags, int32_t *indices,
                      float *posterior, int device);

/*
 * Decode a batch resident in HOST memory on one device: host->device copies,
 * decode and the device->host copy of the indices are all inside the call,
 * which returns when `indices` is complete.  Pinned host buffers copy at full
 * PCIe rate; pageable ones work too.  Device buffers come from a per-device
 * arena the library grows on demand and keeps (torbi_b200_release frees it).
 * This is the entry torbi.from_probabilities(gpu=k) uses for CPU tensors
 * (/root/reference/torbi/core.py:147-191 moves them with .to(device)).
 */
TORBI_B200_API int torbi_b200_decode_host(const void *observation, const int32_t *batch_frames,
                           const float *transition, const float *initial, int64_t batch,
                           int64_t frames, int64_t states, uint32_t flags, int32_t *indices,
                           float *posterior, int device);

/*
 * Same, sharding the independent batch items across several devices of one
 * box (longest-processing-time-first by batch_frames; one host thread and one
 * stream per device; no collective -- each device writes its own rows of the
 * caller's `indices`).  devices[i] are CUDA ordinals; n_devices >= 1.
 */
TORBI_B200_API int torbi_b200_decode_host_multi(const void *observation, const int32_t *batch_frames,
                                 const float *transition, const float *initial,
                                 int64_t batch, int64_t frames, int64_t states,
                                 uint32_t flags, int32_t *indices, float *posterior,
                                 const int *devices, int n_devices);

/*
 * Framewise normalised entropy for the chunked long-sequence mode -- replaces `entropy` of
 * /root/reference/torbi/chunk.py:81-85 (torch CPU ops over the whole (states, frames) transpose):
 *     H[t] = -( sum_s exp(x[t][s]) * x[t][s] ) / log(states)
 * on the observation exactly as stored (the reference's Dataset chunks before any conversion,
 * data/dataset.py:21-23).  One pass over HBM, one warp per frame, fixed summation order.
 *
 *   observation  device (rows, states) fp32 (fp16 / bf16 with TORBI_B200_OBS_F16 / _BF16 in `flags`)
 *   entropy      device (rows, 2) fp32, output: [t][0] = H[t]; [t][1] = sum_s |exp(x) x| / log(states),
 *                the magnitude that bounds the rounding difference to any other fp32 summation order
 *                (|H - H'| <= ~2 states 2^-24 [t][1]): a caller that must reproduce the reference's
 *                split decisions re-evaluates only the frames whose H is that close to the threshold
 *                (torbi_b200/chunk.py does, with the reference's own ops).
 * Enqueued on `stream` on the current device; does not synchronise.
 */
TORBI_B200_API int torbi_b200_entropy(const void *observation, int64_t rows, int64_t states, uint32_t flags,
                                      float *entropy, void *stream);

/* Free the host-entry arenas (all devices).  Safe to call at any time no decode
 * is in flight. */
TORBI_B200_API void torbi_b200_release(void);

/* Counters for bench.py: kernels launched by this library since load, and the
 * name of the forward kernel the last decode on this thread used. */
TORBI_B200_API uint64_t torbi_b200_launch_count(void);
TORBI_B200_API const char *torbi_b200_last_kernel(void);
/* "backtrace" (stored uint16 backpointers) or "backtrace_recompute" (score-history mode) */
TORBI_B200_API const char *torbi_b200_last_backtrace(void);

/* Which forward kernel a decode of this shape and length profile would use on a device with
 * `sm_count` SMs ("forward_tiled", "forward_tiled_32", "forward_wide", "forward_generic", ...).
 * Pure host logic -- no CUDA call -- so schedulers and tests can query the plan.  NULL on
 * invalid arguments (see torbi_b200_last_error). */
TORBI_B200_API const char *torbi_b200_plan(const int32_t *batch_frames, int64_t batch, int64_t frames,
                                           int64_t states, uint32_t flags, int sm_count);

/* Device time of the forward and the backtrace kernel of the last
 * torbi_b200_decode on this thread that was called with
 * TORBI_B200_TIME_KERNELS.  Waits for those kernels to finish.  Either pointer
 * may be NULL. */
TORBI_B200_API int torbi_b200_last_timing(float *forward_ms, float *backtrace_ms);

/*
 * Native file I/O for the file API (from_file / from_files_to_files, core.py:211-368): the
 * reference loads every input with torch.load and writes every result with torch.save from one
 * Python thread (data/dataset.py:17-22, core.py:466-473).  These three host-only entries parse the
 * torch.save container (a zip of STORED records) in C++, read the payloads with pread() straight
 * into the caller's (pinned) batch buffer from a pool of host threads, and write int32 results in
 * the same container format, readable by torch.load.  No CUDA call is made.
 */
#define TORBI_B200_PT_OK 0
#define TORBI_B200_PT_IO_ERROR 1     /* open / read / write failed                                   */
#define TORBI_B200_PT_NOT_A_ZIP 2    /* legacy (non-zip) torch.save file or damaged archive          */
#define TORBI_B200_PT_UNSUPPORTED 3  /* valid file the fast path does not cover (zip64, compression,
                                        several tensors, non-contiguous view, other pickled objects):
                                        the caller falls back to torch.load                          */
#define TORBI_B200_PT_F32 0
#define TORBI_B200_PT_F16 1
#define TORBI_B200_PT_BF16 2
#define TORBI_B200_PT_F64 3
#define TORBI_B200_PT_I32 4
#define TORBI_B200_PT_I64 5

typedef struct {
    int64_t data_offset; /* byte offset in the file of the tensor's first element             */
    int64_t nbytes;      /* bytes of the (contiguous, row-major) tensor                        */
    int64_t shape[4];
    int32_t ndim;        /* <= 4                                                               */
    int32_t dtype;       /* TORBI_B200_PT_F32 ...                                              */
    int32_t status;      /* TORBI_B200_PT_OK or why the fast path does not apply to this file  */
    int32_t reserved;
} torbi_b200_pt_info;

/* Describe `n` torch.save files (one tensor each).  Returns the number of files whose status is
 * not TORBI_B200_PT_OK (0 = all good), or a negative TORBI_B200_ERR_* for bad arguments.
 * `threads` <= 0 uses every host core. */
TORBI_B200_API int torbi_b200_pt_inspect(const char *const *paths, int64_t n, torbi_b200_pt_info *info,
                                         int threads);

/* Read the payload of file i (as described by info[i], status OK) to destinations[i] and zero the
 * rest of its capacities[i] bytes (the zero padding of collate(), data/collate.py:26-31).  Large
 * payloads are split across threads.  Returns the number of failed reads. */
TORBI_B200_API int torbi_b200_pt_read(const char *const *paths, const torbi_b200_pt_info *info, int64_t n,
                                      void *const *destinations, const int64_t *capacities, int threads);

/* Write n contiguous int32 tensors (shape of tensor i = shapes[4*i .. 4*i + ndims[i])) as torch.save
 * files.  Returns the number of failed writes. */
TORBI_B200_API int torbi_b200_pt_write_i32(const char *const *paths, const int32_t *const *data,
                                           const int64_t *shapes, const int32_t *ndims, int64_t n,
                                           int threads);

/* Self-test: the kernels skip logf(expf(x) + FLT_MIN) where it is the identity (every fp32 x in
 * (-64, -4], found exhaustively).  Evaluates the shortcut and the plain formula for ALL 2^32 fp32
 * bit patterns on the current device and reports how many disagree (must be 0).  Takes ~1 s;
 * allocates 8 bytes of device memory for the duration of the call. */
TORBI_B200_API int torbi_b200_selftest_conversion(uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* TORBI_B200_H_ */
