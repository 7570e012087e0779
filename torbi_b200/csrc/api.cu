// C ABI of libtorbi_b200.so (declared in include/torbi_b200.h): validation,
// workspace carving, kernel selection, and the host-buffer / multi-device
// entries.  Replaces viterbi_decode_cuda and its launch glue
// (/root/reference/torbi/csrc/cuda/viterbi.cu:203-362) -- but validates its
// arguments, checks every launch, runs on the caller's stream and never
// touches the current-device setting of the caller.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

using namespace torbi_b200;

namespace {

thread_local char g_error[512] = "";
thread_local const char *g_last_kernel = "none";
thread_local const char *g_last_backtrace = "none";
std::atomic<uint64_t> g_launches{0};

// events for TORBI_B200_TIME_KERNELS, per host thread, bound to one device
struct TimingEvents {
    int device = -1;
    bool valid = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};
thread_local TimingEvents g_timing;

cudaError_t timing_prepare(int device) {
    if (g_timing.device == device && g_timing.ev[0]) return cudaSuccess;
    for (auto &e : g_timing.ev) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    for (auto &e : g_timing.ev) {
        cudaError_t err = cudaEventCreate(&e);
        if (err != cudaSuccess) return err;
    }
    g_timing.device = device;
    return cudaSuccess;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int fail_cuda(cudaError_t e, const char *what) {
    return fail(TORBI_B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DeviceInfo {
    int sm_count = 0;
    int cc_major = 0;
    bool ok = false;
};

std::mutex g_dev_mutex;
std::map<int, DeviceInfo> g_devices;

// cudaGetDeviceProperties on every call is what the reference does
// (viterbi.cu:217-218, :271-272); here it is cached per device.
int device_info(int device, DeviceInfo *out) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    auto it = g_devices.find(device);
    if (it == g_devices.end()) {
        DeviceInfo d;
        cudaError_t e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess)
            e = cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, device);
        if (e != cudaSuccess) return fail_cuda(e, "query device");
        d.ok = true;
        it = g_devices.emplace(device, d).first;
    }
    *out = it->second;
    if (out->cc_major != 10)
        return fail(TORBI_B200_ERR_CUDA,
                    "device %d has compute capability %d.x; libtorbi_b200 carries sm_100a code only",
                    device, out->cc_major);
    return TORBI_B200_OK;
}

// No C++ exception may cross the C ABI (host allocations for a huge batch, thread creation).
template <typename Fn>
int guarded(Fn fn) {
    try {
        return fn();
    } catch (const std::bad_alloc &) {
        return fail(TORBI_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (const std::exception &error) {
        return fail(TORBI_B200_ERR_CUDA, "internal error: %s", error.what());
    } catch (...) {
        return fail(TORBI_B200_ERR_CUDA, "internal error");
    }
}

int validate_shape(int64_t B, int64_t T, int64_t S) {
    if (B < 1 || T < 1 || S < 1)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "batch, frames and states must be >= 1 (got %lld, %lld, %lld)",
                    (long long)B, (long long)T, (long long)S);
    if (S > 65535)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "states = %lld exceeds 65535 (uint16 backpointers)", (long long)S);
    if (B > 0x7fffffff || T > 0x7fffffff)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "batch and frames must fit in int32");
    return TORBI_B200_OK;
}

}  // namespace

extern "C" {

int torbi_b200_abi_version(void) { return TORBI_B200_ABI_VERSION; }

const char *torbi_b200_last_error(void) { return g_error; }

uint64_t torbi_b200_launch_count(void) { return g_launches.load(); }

const char *torbi_b200_last_kernel(void) { return g_last_kernel; }

const char *torbi_b200_last_backtrace(void) { return g_last_backtrace; }

int torbi_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
            ++usable;
    }
    return usable;
}

size_t torbi_b200_workspace_bytes(int64_t B, int64_t T, int64_t S, uint32_t flags) {
    if (validate_shape(B, T, S) != TORBI_B200_OK) return 0;
    return carve_workspace(nullptr, B, B * T, S, workspace_has_history(flags)).bytes;
}

}  // extern "C"

namespace {

enum Choice { kGeneric, kTiled, kWide, kSplit, kCluster, kFlow };
// whether the cost model may pick the cluster kernel by itself (else only TORBI_B200_FORCE_CLUSTER does)
constexpr bool kClusterAuto = false;
constexpr int64_t kFlowAutoMaxBatch = 16;

// Kernel selection: estimated time of the batch-tiled, split-K, cluster and wide decompositions for
// this length profile (`sorted` = batch_frames, longest first), unless a flag forces one.  Also fixes
// a.tile_rows.  `flags` may lose TORBI_B200_BANDED (a hint) where the tiled kernel cannot run.
int choose_kernel(ForwardArgs &a, const std::vector<int32_t> &sorted, uint32_t &flags, Choice *out) {
    const int64_t B = a.B, S = a.S;
    double tiled_seconds = 0.0;
    a.tile_rows = tiled_choose_rows(sorted.data(), B, S, a.sm_count, &tiled_seconds);
    if (flags & TORBI_B200_TILE_ROWS_MASK)
        a.tile_rows = 16 << (((flags & TORBI_B200_TILE_ROWS_MASK) >> 12) - 1);
    const bool can_tile = forward_tiled_supported(a), can_wide = forward_wide_supported(a);
    // TORBI_B200_BANDED is a speed hint (results are identical either way): where the tiled kernel cannot
    // run (more destination tiles than SMs, i.e. S > 80 x #SM) the flag is dropped and the selection
    // proceeds as without it.  Only FORCE_TILED / TILE_ROWS_* insist on the tiled kernel.
    if ((flags & TORBI_B200_BANDED) && !can_tile) {
        flags &= ~TORBI_B200_BANDED;
        a.flags = flags;
    }
    // the split-K kernels exist in score-history mode only, dense transitions only
    const bool history_ok = workspace_has_history(flags) && !(flags & TORBI_B200_BANDED);
    const bool can_split = forward_split_supported(a) && history_ok;
    const bool can_cluster = forward_cluster_supported(a) && history_ok;
    Choice choice = kGeneric;
    double best_seconds = 1e300;
    if (can_tile) { choice = kTiled; best_seconds = tiled_seconds; }
    if (can_split) {
        const double seconds = split_estimate_seconds(sorted.data(), B, S, a.sm_count);
        if (seconds < best_seconds) { choice = kSplit; best_seconds = seconds; }
    }
    if (can_cluster && kClusterAuto) {
        const double seconds = cluster_estimate_seconds(sorted.data(), B, S, a.sm_count);
        if (seconds < best_seconds) { choice = kCluster; best_seconds = seconds; }
    }
    if (can_wide) {
        const double seconds = wide_estimate_seconds(sorted.data(), B, S, a.sm_count);
        if (seconds < best_seconds) { choice = kWide; best_seconds = seconds; }
    }
    const bool can_flow = forward_flow_supported(a);
    // (sequences go through the flow kernel one after the other: chosen by itself only for a handful)
    if (can_flow && B <= kFlowAutoMaxBatch && flow_estimate_seconds(sorted.data(), B, S, a.sm_count) < best_seconds)
        choice = kFlow;
    if (flags & TORBI_B200_FORCE_SPLIT) {
        if (!can_split)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "split kernel does not support this shape / these flags");
        choice = kSplit;
    }
    if (flags & TORBI_B200_FORCE_CLUSTER) {
        if (!can_cluster)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "cluster kernel does not support this shape / these flags");
        choice = kCluster;
    }
    if (flags & (TORBI_B200_FORCE_TILED | TORBI_B200_TILE_ROWS_MASK | TORBI_B200_BANDED)) {
        if (!can_tile) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "tiled kernel does not support this shape");
        choice = kTiled;
    }
    if (flags & TORBI_B200_FORCE_WIDE) {
        if (!can_wide) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "wide kernel does not support this shape");
        choice = kWide;
    }
    if (flags & TORBI_B200_FORCE_FLOW) {
        if (!can_flow) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "flow kernel does not support this shape");
        choice = kFlow;
    }
    if (flags & TORBI_B200_FORCE_GENERIC) choice = kGeneric;
    *out = choice;
    return TORBI_B200_OK;
}

const char *kernel_name(int tile_rows, Choice choice, uint32_t flags) {
    if (choice == kTiled) {
        if (flags & TORBI_B200_BANDED)
            return tile_rows == 64 ? "forward_tiled_banded" : tile_rows == 32 ? "forward_tiled_banded_32" : "forward_tiled_banded_16";
        return tile_rows == 64 ? "forward_tiled" : tile_rows == 32 ? "forward_tiled_32" : "forward_tiled_16";
    }
    if (choice == kSplit) return "forward_split";
    if (choice == kCluster) return "forward_cluster";
    if (choice == kFlow) return "forward_flow";
    return choice == kWide ? "forward_wide" : "forward_generic";
}

// Everything about a call that is decided on the host before anything is enqueued: validated
// lengths, the longest-first order, the kernel, and whether the workspace tail holds the fp32 score
// history (4 bytes per (b,t,s)) or uint16 backpointers (2 bytes).
struct Plan {
    Choice choice = kGeneric;
    int tile_rows = 64;
    uint32_t flags = 0;          // caller's flags, minus hints that do not apply
    bool history = false;
    int64_t max_frames = 0;
    std::vector<int32_t> order;  // (pad_batch,) items longest first, -1 = empty slot
    // first row of every item in observation / history / backpointers / indices: b * T in the padded
    // (B,T,S) layout, the caller's offsets in the packed (rows,S) one
    std::vector<int64_t> item_row;
    // first row of every item in the score history / backpointers: compact, item after item
    std::vector<int64_t> hist_row;
    int64_t rows = 0;            // rows of the history / backpointer region = sum(batch_frames)
    bool packed = false;
    size_t workspace_bytes = 0;
};

// `item_row` == nullptr: padded layout, `T` frames per item.  Otherwise the packed layout: `rows` rows in
// all, item b occupies rows [item_row[b], item_row[b] + batch_frames[b]); T is ignored.
int make_plan(const int32_t *batch_frames, int64_t B, int64_t T, int64_t S, uint32_t flags, int sm_count,
              Plan *plan, const int64_t *item_row = nullptr, int64_t rows = 0) {
    if (!batch_frames) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    plan->packed = item_row != nullptr;
    if (plan->packed) {
        if (B < 1 || rows < 1)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "batch and rows must be >= 1 (got %lld, %lld)", (long long)B, (long long)rows);
        T = 1;
        for (int64_t b = 0; b < B; ++b) T = std::max<int64_t>(T, batch_frames[b]);
    }
    int rc = validate_shape(B, T, S);
    if (rc) return rc;
    if ((flags & TORBI_B200_OBS_F16) && (flags & TORBI_B200_OBS_BF16))
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "TORBI_B200_OBS_F16 and TORBI_B200_OBS_BF16 are exclusive");
    plan->max_frames = 0;
    for (int64_t b = 0; b < B; ++b) {
        // the reference trusts these and reads out of bounds for 0 (viterbi.cpp:152)
        if (batch_frames[b] < 1 || batch_frames[b] > T)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "batch_frames[%lld] = %d outside [1, %lld]",
                        (long long)b, batch_frames[b], (long long)T);
        plan->max_frames = std::max<int64_t>(plan->max_frames, batch_frames[b]);
    }
    plan->item_row.resize(B);
    if (plan->packed) {
        // every item inside the buffer, no two items sharing a row (the score history / backpointers
        // and the output use the same row numbering)
        std::vector<int64_t> by_row(B);
        std::iota(by_row.begin(), by_row.end(), 0);
        std::sort(by_row.begin(), by_row.end(), [&](int64_t x, int64_t y) { return item_row[x] < item_row[y]; });
        int64_t end = 0;
        for (int64_t k = 0; k < B; ++k) {
            const int64_t b = by_row[k];
            if (item_row[b] < end || item_row[b] + batch_frames[b] > rows)
                return fail(TORBI_B200_ERR_INVALID_ARGUMENT,
                            "item %lld: rows [%lld, %lld) overlap another item or leave the buffer of %lld rows",
                            (long long)b, (long long)item_row[b], (long long)(item_row[b] + batch_frames[b]), (long long)rows);
            end = item_row[b] + batch_frames[b];
            plan->item_row[b] = item_row[b];
        }
    } else {
        for (int64_t b = 0; b < B; ++b) plan->item_row[b] = b * T;
    }
    // the score history / backpointers only ever hold the frames that exist (the reference zero-fills a
    // (batch, frames, states) trellis, viterbi.cpp:201-203)
    plan->hist_row.resize(B);
    plan->rows = 0;
    for (int64_t b = 0; b < B; ++b) {
        plan->hist_row[b] = plan->rows;
        plan->rows += batch_frames[b];
    }
    plan->order.assign(pad_batch(B), -1);
    std::iota(plan->order.begin(), plan->order.begin() + B, 0);
    std::stable_sort(plan->order.begin(), plan->order.begin() + B,
                     [&](int32_t x, int32_t y) { return batch_frames[x] > batch_frames[y]; });
    std::vector<int32_t> sorted(B);
    for (int64_t i = 0; i < B; ++i) sorted[i] = batch_frames[plan->order[i]];
    ForwardArgs a{};
    a.B = B; a.T = T; a.S = S;
    a.flags = flags;
    a.sm_count = sm_count;
    // layout only, never dereferenced: the `supported` predicates only look at shapes
    a.ws = carve_workspace(reinterpret_cast<void *>(uintptr_t(256)), B, plan->rows, S, workspace_has_history(flags));
    a.observation = reinterpret_cast<const float *>(uintptr_t(256));
    rc = choose_kernel(a, sorted, flags, &plan->choice);
    if (rc) return rc;
    plan->flags = flags;
    plan->tile_rows = a.tile_rows;
    plan->history = plan->choice == kSplit || plan->choice == kCluster ||
                    (plan->choice == kTiled && workspace_has_history(flags));
    plan->workspace_bytes = carve_workspace(nullptr, B, plan->rows, S, plan->history).bytes;
    return TORBI_B200_OK;
}

struct DecodeOptions {
    // device word the forward kernel polls when the observations are still being copied in (host
    // entry); nullptr otherwise.  `*streamed` reports whether the chosen kernel honours it.
    const int32_t *frames_ready = nullptr;
    bool *streamed = nullptr;
    // device (B, (T - 1) / every, S): forward scores P_t for t = every, 2 every, ... (t < frames_b)
    float *checkpoints = nullptr;
    int64_t checkpoint_every = 0;
    const Plan *plan = nullptr;  // made by the caller already (host entry)
    // packed layout (torbi_b200_decode_packed): host (B,) first rows and the row count; T is ignored
    const int64_t *item_row = nullptr;
    int64_t rows = 0;
};

int decode_device(const void *observation, const int32_t *batch_frames,
                  const float *transition, const float *initial, int64_t B, int64_t T,
                  int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                  void *workspace, size_t workspace_bytes, void *stream_, const DecodeOptions &opt) {
    g_error[0] = 0;
    if (!observation || !batch_frames || !transition || !initial || !indices || !workspace)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    DeviceInfo info;
    int rc = device_info(device, &info);
    if (rc) return rc;
    Plan own_plan;
    if (!opt.plan) {
        rc = make_plan(batch_frames, B, T, S, flags, info.sm_count, &own_plan, opt.item_row, opt.rows);
        if (rc) return rc;
    }
    const Plan &plan = opt.plan ? *opt.plan : own_plan;
    flags = plan.flags;
    if (plan.packed) {
        T = plan.max_frames;
        if (opt.checkpoints)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "checkpoints are not available in the packed layout");
    }
    if (opt.checkpoints && opt.checkpoint_every < 1)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "checkpoint_every must be >= 1");
    if (reinterpret_cast<uintptr_t>(workspace) % 256)
        return fail(TORBI_B200_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    // the kernel is chosen first, so a call that ends up on the wide / generic kernel (uint16
    // backpointers) only needs half the trailing region of a score-history call
    Workspace ws = carve_workspace(workspace, B, plan.rows, S, plan.history);
    if (workspace_bytes < ws.bytes)
        return fail(TORBI_B200_ERR_WORKSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, ws.bytes);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    // lengths and the longest-first processing order go to the device
    e = cudaMemcpyAsync(ws.frames, batch_frames, sizeof(int32_t) * B, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ws.order, plan.order.data(), sizeof(int32_t) * plan.order.size(),
                            cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ws.item_row, plan.item_row.data(), sizeof(int64_t) * B, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ws.hist_row, plan.hist_row.data(), sizeof(int64_t) * B, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.sync, 0, sizeof(uint32_t) * kSyncWords, stream);
    if (e != cudaSuccess) return fail_cuda(e, "stage batch_frames");
    // (pageable host memory: the async copies have been staged by the time cudaMemcpyAsync returns)

    ForwardArgs a{};
    a.observation = static_cast<const float *>(observation);
    a.obs_dtype = (flags & TORBI_B200_OBS_F16) ? 1 : (flags & TORBI_B200_OBS_BF16) ? 2 : 0;
    a.transition = transition;
    a.initial = initial;
    a.posterior = posterior;
    a.B = B; a.T = T; a.S = S;
    a.max_frames = plan.max_frames;
    a.flags = flags;
    a.sm_count = info.sm_count;
    a.tile_rows = plan.tile_rows;
    a.ws = ws;
    a.frames_ready = opt.frames_ready;
    a.checkpoints = opt.checkpoints;
    a.checkpoint_every = opt.checkpoints ? opt.checkpoint_every : 0;
    if (opt.checkpoints && plan.choice == kTiled && !plan.history)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT,
                    "checkpoints are not available from the tiled kernel in backpointer mode");
    const bool timed = flags & TORBI_B200_TIME_KERNELS;
    g_timing.valid = false;
    if (timed) {
        e = timing_prepare(device);
        if (e == cudaSuccess) e = cudaEventRecord(g_timing.ev[0], stream);
        if (e != cudaSuccess) return fail_cuda(e, "timing events");
    }

    const Choice choice = plan.choice;
    const bool tiled = choice != kGeneric && choice != kFlow;   // kernels that honour frames_ready
    if (opt.streamed) *opt.streamed = tiled;
    g_last_kernel = kernel_name(plan.tile_rows, choice, flags);
    if (!tiled && opt.frames_ready)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "internal: this kernel cannot stream");
    if (choice == kGeneric) {
        e = launch_forward_generic(a, stream);
    } else if (choice == kFlow) {
        e = launch_forward_flow(a, stream);
    } else if (choice == kWide) {
        e = launch_forward_wide(a, stream);
    } else if (choice == kSplit) {
        e = launch_forward_split(a, stream);
    } else if (choice == kCluster) {
        e = launch_forward_cluster(a, stream);
    } else if (plan.history) {
        // score-history mode (default): pure max-plus forward, argmax recomputed along the path
        e = launch_forward_tiled_hist(a, stream);
    } else {
        e = launch_forward_tiled(a, stream);
    }
    if (e != cudaSuccess) return fail_cuda(e, g_last_kernel);
    int launches = choice == kGeneric ? 1 : (choice == kTiled && (flags & TORBI_B200_BANDED) ? 3 : 2);   // (pad_transition + [analyze_band] +) forward
    if (opt.checkpoints && plan.history) {
        // the score history holds every P_t already: gather the requested rows
        e = launch_gather_checkpoints(a, stream);
        if (e != cudaSuccess) return fail_cuda(e, "gather_checkpoints");
        ++launches;
    }
    if (timed) cudaEventRecord(g_timing.ev[1], stream);
    const bool history = plan.history;
    const int64_t T_out = plan.packed ? 0 : T;           // 0: indices are packed like the observations
    int backtrace_launches = 1;
    g_last_backtrace = history ? "backtrace_recompute" : "backtrace";
    if (history) {
        e = launch_backtrace_recompute(ws, indices, B, T_out, S, stream);
    } else if (plan.max_frames >= kWindowedBacktraceMinFrames) {
        g_last_backtrace = "backtrace_blocked";
        e = launch_backtrace_blocked(ws, indices, B, T_out, S, plan.max_frames, &backtrace_launches, stream);
    } else {
        e = launch_backtrace(ws, indices, B, T_out, S, stream);
    }
    if (e != cudaSuccess) return fail_cuda(e, g_last_backtrace);
    if (timed) {
        cudaEventRecord(g_timing.ev[2], stream);
        g_timing.valid = true;
    }
    g_launches += launches + backtrace_launches;
    return TORBI_B200_OK;
}

}  // namespace

extern "C" {

int torbi_b200_decode(const void *observation, const int32_t *batch_frames,
                      const float *transition, const float *initial, int64_t B, int64_t T,
                      int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                      void *workspace, size_t workspace_bytes, void *stream) {
    return guarded([&] {
        return decode_device(observation, batch_frames, transition, initial, B, T, S, flags, indices,
                             posterior, workspace, workspace_bytes, stream, DecodeOptions{});
    });
}

int torbi_b200_decode_checkpointed(const void *observation, const int32_t *batch_frames,
                                   const float *transition, const float *initial, int64_t B, int64_t T,
                                   int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                                   void *workspace, size_t workspace_bytes, void *stream,
                                   float *checkpoints, int64_t checkpoint_every) {
    return guarded([&] {
        if (!checkpoints) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
        DecodeOptions opt;
        opt.checkpoints = checkpoints;
        opt.checkpoint_every = checkpoint_every;
        return decode_device(observation, batch_frames, transition, initial, B, T, S, flags, indices,
                             posterior, workspace, workspace_bytes, stream, opt);
    });
}

int torbi_b200_decode_packed(const void *observation, const int64_t *item_row, const int32_t *batch_frames,
                             const float *transition, const float *initial, int64_t B, int64_t rows,
                             int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                             void *workspace, size_t workspace_bytes, void *stream) {
    return guarded([&] {
        if (!item_row) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
        DecodeOptions opt;
        opt.item_row = item_row;
        opt.rows = rows;
        return decode_device(observation, batch_frames, transition, initial, B, /*T=*/1, S, flags, indices,
                             posterior, workspace, workspace_bytes, stream, opt);
    });
}

size_t torbi_b200_workspace_bytes_packed(const int64_t *item_row, const int32_t *batch_frames, int64_t B,
                                         int64_t rows, int64_t S, uint32_t flags, int sm_count) {
    g_error[0] = 0;
    if (sm_count < 1 || !item_row) return 0;
    try {
        Plan plan;
        if (make_plan(batch_frames, B, 1, S, flags, sm_count, &plan, item_row, rows) != TORBI_B200_OK) return 0;
        return plan.workspace_bytes;
    } catch (...) {
        fail(TORBI_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
        return 0;
    }
}

size_t torbi_b200_workspace_bytes_planned(const int32_t *batch_frames, int64_t B, int64_t T, int64_t S,
                                          uint32_t flags, int sm_count) {
    g_error[0] = 0;
    if (sm_count < 1) return 0;
    try {
        Plan plan;
        if (make_plan(batch_frames, B, T, S, flags, sm_count, &plan) != TORBI_B200_OK) return 0;
        return plan.workspace_bytes;
    } catch (...) {
        fail(TORBI_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
        return 0;
    }
}

const char *torbi_b200_plan(const int32_t *batch_frames, int64_t B, int64_t T, int64_t S,
                            uint32_t flags, int sm_count) {
    g_error[0] = 0;
    if (sm_count < 1) return nullptr;
    try {
        Plan plan;
        if (make_plan(batch_frames, B, T, S, flags, sm_count, &plan) != TORBI_B200_OK) return nullptr;
        return kernel_name(plan.tile_rows, plan.choice, plan.flags);
    } catch (...) {
        fail(TORBI_B200_ERR_OUT_OF_MEMORY, "host allocation failed");
        return nullptr;
    }
}

int torbi_b200_last_timing(float *forward_ms, float *backtrace_ms) {
    if (!g_timing.valid)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT,
                    "no decode with TORBI_B200_TIME_KERNELS on this thread");
    cudaError_t e = cudaEventSynchronize(g_timing.ev[2]);
    float fwd = 0.f, back = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&fwd, g_timing.ev[0], g_timing.ev[1]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&back, g_timing.ev[1], g_timing.ev[2]);
    if (e != cudaSuccess) return fail_cuda(e, "timing events");
    if (forward_ms) *forward_ms = fwd;
    if (backtrace_ms) *backtrace_ms = back;
    return TORBI_B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Host-buffer entries
// ---------------------------------------------------------------------------
namespace {

constexpr int kMaxChunks = 64;

struct Arena {
    void *base = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;       // kernels, small copies, results
    cudaStream_t copy_stream = nullptr;  // observation chunks streamed in under the kernel
    cudaEvent_t reset_event = nullptr;
    int32_t *ready_host = nullptr;       // pinned: frames available after each chunk
    std::mutex busy;  // one decode at a time per device arena
};

std::mutex g_arena_mutex;
std::map<int, Arena *> g_arenas;

Arena *arena_for(int device) {
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    auto it = g_arenas.find(device);
    if (it == g_arenas.end()) it = g_arenas.emplace(device, new Arena()).first;
    return it->second;
}

struct DeviceGuard {
    int previous = -1;
    explicit DeviceGuard(int device) {
        cudaGetDevice(&previous);
        cudaSetDevice(device);
    }
    ~DeviceGuard() {
        if (previous >= 0) cudaSetDevice(previous);
    }
};

int arena_reserve(Arena *arena, int device, size_t total);

// Decode `count` items (given by index into the caller's batch) on one device.
int decode_items_on_device(const void *observation_, const int32_t *batch_frames,
                           const float *transition, const float *initial,
                           const int32_t *items, int64_t count, int64_t T, int64_t S,
                           uint32_t flags, int32_t *indices, float *posterior, int device) {
    DeviceGuard guard(device);
    DeviceInfo info;
    int rc = device_info(device, &info);
    if (rc) return rc;
    Arena *arena = arena_for(device);
    std::lock_guard<std::mutex> lock(arena->busy);

    // everything that can be rejected is rejected before a single copy is enqueued: lengths, flags,
    // the kernel choice (a forced kernel that does not support the shape) and the workspace size
    std::vector<int32_t> frames(count);
    for (int64_t i = 0; i < count; ++i) frames[i] = batch_frames[items[i]];
    Plan plan;
    rc = make_plan(frames.data(), count, T, S, flags, info.sm_count, &plan);
    if (rc) return rc;

    // fp16 / bf16 observations stay 2 bytes per element on the wire and in HBM; the kernels upcast
    // in their observation load (exact: every half value is an fp32 value)
    const size_t elem = (flags & (TORBI_B200_OBS_F16 | TORBI_B200_OBS_BF16)) ? 2 : 4;
    const char *observation = static_cast<const char *>(observation_);
    const size_t obs_bytes = align_up(elem * count * T * S, 256);
    const size_t tr_bytes = align_up(sizeof(float) * S * S, 256);
    const size_t init_bytes = align_up(sizeof(float) * S, 256);
    const size_t idx_bytes = align_up(sizeof(int32_t) * count * T, 256);
    const size_t post_bytes = posterior ? align_up(sizeof(float) * count * S, 256) : 0;
    const size_t ws_bytes = plan.workspace_bytes;
    const size_t total = obs_bytes + tr_bytes + init_bytes + idx_bytes + post_bytes + ws_bytes + 256;

    cudaError_t e;
    rc = arena_reserve(arena, device, total);
    if (rc) return rc;
    char *p = static_cast<char *>(arena->base);
    char *d_obs = p; p += obs_bytes;
    float *d_tr = reinterpret_cast<float *>(p); p += tr_bytes;
    float *d_init = reinterpret_cast<float *>(p); p += init_bytes;
    int32_t *d_idx = reinterpret_cast<int32_t *>(p); p += idx_bytes;
    float *d_post = posterior ? reinterpret_cast<float *>(p) : nullptr; p += post_bytes;
    void *d_ws = p; p += ws_bytes;
    int32_t *d_ready = reinterpret_cast<int32_t *>(p);
    cudaStream_t stream = arena->stream;

    e = cudaMemcpyAsync(d_tr, transition, sizeof(float) * S * S, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_init, initial, sizeof(float) * S, cudaMemcpyHostToDevice, stream);
    // once anything is in flight, every exit waits for both streams: the next call on this arena
    // must not overlap leftover copies, and the caller may free its (pinned) buffers on return
    struct Drain {
        Arena *arena;
        bool armed = true;
        ~Drain() {
            if (!armed) return;
            cudaStreamSynchronize(arena->stream);
            cudaStreamSynchronize(arena->copy_stream);
            cudaGetLastError();
        }
    } drain{arena};

    // Large pinned batches are streamed in along T on a second stream while the
    // forward kernel already runs: step t only needs frame t (the kernel polls
    // `d_ready`).  Small, pageable or generic-kernel batches are copied up front.
    const size_t row_bytes = elem * T * S;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, observation) == cudaSuccess &&
                        attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const bool stream_in = pinned && plan.choice != kGeneric && plan.choice != kFlow && T >= 64 && row_bytes < (size_t(1) << 31) &&
                           (size_t)count * row_bytes >= (size_t(64) << 20);
    const int chunks = stream_in ? (int)std::min<int64_t>(kMaxChunks, std::max<int64_t>(2, T / 16)) : 1;
    // Chunk boundaries along T.  The forward kernel trails the copy by one chunk (step t needs the whole
    // chunk holding frame t), and at config 2 the copy is only ~1.3x faster than the kernel consumes
    // frames, so what is exposed is roughly the FIRST chunk's copy time: many small chunks (64: 1.7 ms
    // each at config 2 instead of 6.6 ms with 16), the first two halved again.  Chunks must not grow
    // faster than that 1.3x either: a geometric ramp T/128, T/64, ... stalls the kernel for 7 ms at the
    // first big chunk (profiles/r01_e2e_copy_experiments.txt).
    std::vector<int64_t> bound(chunks + 1);
    for (int c = 0; c <= chunks; ++c) bound[c] = T * c / chunks;
    if (chunks == kMaxChunks && T >= 4 * kMaxChunks) {
        bound[1] = T / (4 * chunks);
        bound[2] = T / (2 * chunks);
        bound[3] = T / chunks;
        for (int c = 4; c < chunks; ++c) bound[c] = bound[3] + (T - bound[3]) * (c - 3) / (chunks - 3);
    }
    if (stream_in) {
        e = cudaMemsetAsync(d_ready, 0, sizeof(int32_t), arena->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(arena->reset_event, arena->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, arena->reset_event, 0);
    }
    for (int c = 0; c < chunks && e == cudaSuccess; ++c) {
        const int64_t t0 = bound[c], t1 = bound[c + 1];
        // items that are consecutive in the caller's batch travel in one copy
        for (int64_t i = 0; i < count && e == cudaSuccess;) {
            int64_t run = 1;
            while (i + run < count && items[i + run] == items[i] + run) ++run;
            const char *src = observation + ((int64_t)items[i] * T * S + t0 * S) * elem;
            char *dst = d_obs + (i * T * S + t0 * S) * elem;
            if (chunks == 1)
                e = cudaMemcpyAsync(dst, src, row_bytes * run, cudaMemcpyHostToDevice, stream);
            else
                e = cudaMemcpy2DAsync(dst, row_bytes, src, row_bytes, elem * (t1 - t0) * S,
                                      run, cudaMemcpyHostToDevice, arena->copy_stream);
            i += run;
        }
        if (stream_in && e == cudaSuccess) {
            arena->ready_host[c] = (int32_t)t1;
            e = cudaMemcpyAsync(d_ready, &arena->ready_host[c], sizeof(int32_t),
                                cudaMemcpyHostToDevice, arena->copy_stream);
        }
    }
    if (e != cudaSuccess) return fail_cuda(e, "host->device copy");

    bool streamed = false;
    DecodeOptions opt;
    opt.frames_ready = stream_in ? d_ready : nullptr;
    opt.streamed = &streamed;
    opt.plan = &plan;
    rc = decode_device(d_obs, frames.data(), d_tr, d_init, count, T, S, flags, d_idx, d_post, d_ws,
                       ws_bytes, stream, opt);
    if (rc) return rc;

    for (int64_t i = 0; i < count && e == cudaSuccess;) {
        int64_t run = 1;
        while (i + run < count && items[i + run] == items[i] + run) ++run;
        e = cudaMemcpyAsync(indices + (int64_t)items[i] * T, d_idx + i * T,
                            sizeof(int32_t) * run * T, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess && posterior)
            e = cudaMemcpyAsync(posterior + (int64_t)items[i] * S, d_post + i * S,
                                sizeof(float) * run * S, cudaMemcpyDeviceToHost, stream);
        i += run;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess && stream_in) e = cudaStreamSynchronize(arena->copy_stream);
    if (e != cudaSuccess) return fail_cuda(e, "device->host copy / synchronize");
    drain.armed = false;
    return TORBI_B200_OK;
}

// Streams, the pinned ready words and at least `total` bytes of device memory for this device's arena.
int arena_reserve(Arena *arena, int device, size_t total) {
    cudaError_t e;
    if (!arena->stream) {
        e = cudaStreamCreateWithFlags(&arena->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&arena->copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&arena->reset_event, cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaHostAlloc(reinterpret_cast<void **>(&arena->ready_host),
                              sizeof(int32_t) * kMaxChunks, cudaHostAllocDefault);
        if (e != cudaSuccess) return fail_cuda(e, "create streams");
    }
    if (arena->bytes < total) {
        if (arena->base) cudaFree(arena->base);
        arena->base = nullptr;
        arena->bytes = 0;
        e = cudaMalloc(&arena->base, total);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(TORBI_B200_ERR_OUT_OF_MEMORY, "device %d: cannot allocate %zu bytes: %s",
                        device, total, cudaGetErrorString(e));
        }
        arena->bytes = total;
    }
    return TORBI_B200_OK;
}

// Packed layout from host memory: one copy in of the (rows, S) observations (no padding crosses the
// bus), decode, one copy out of the (rows,) indices.
int decode_packed_on_device(const void *observation, const int64_t *item_row, const int32_t *batch_frames,
                            const float *transition, const float *initial, int64_t B, int64_t rows, int64_t S,
                            uint32_t flags, int32_t *indices, float *posterior, int device) {
    DeviceGuard guard(device);
    DeviceInfo info;
    int rc = device_info(device, &info);
    if (rc) return rc;
    Arena *arena = arena_for(device);
    std::lock_guard<std::mutex> lock(arena->busy);
    Plan plan;
    rc = make_plan(batch_frames, B, 1, S, flags, info.sm_count, &plan, item_row, rows);
    if (rc) return rc;
    const size_t elem = (flags & (TORBI_B200_OBS_F16 | TORBI_B200_OBS_BF16)) ? 2 : 4;
    const size_t obs_bytes = align_up(elem * rows * S, 256);
    const size_t tr_bytes = align_up(sizeof(float) * S * S, 256);
    const size_t init_bytes = align_up(sizeof(float) * S, 256);
    const size_t idx_bytes = align_up(sizeof(int32_t) * rows, 256);
    const size_t post_bytes = posterior ? align_up(sizeof(float) * B * S, 256) : 0;
    const size_t ws_bytes = plan.workspace_bytes;
    rc = arena_reserve(arena, device, obs_bytes + tr_bytes + init_bytes + idx_bytes + post_bytes + ws_bytes + 256);
    if (rc) return rc;
    char *p = static_cast<char *>(arena->base);
    char *d_obs = p; p += obs_bytes;
    float *d_tr = reinterpret_cast<float *>(p); p += tr_bytes;
    float *d_init = reinterpret_cast<float *>(p); p += init_bytes;
    int32_t *d_idx = reinterpret_cast<int32_t *>(p); p += idx_bytes;
    float *d_post = posterior ? reinterpret_cast<float *>(p) : nullptr; p += post_bytes;
    void *d_ws = p;
    cudaStream_t stream = arena->stream;
    struct Drain {
        cudaStream_t stream;
        bool armed = true;
        ~Drain() {
            if (!armed) return;
            cudaStreamSynchronize(stream);
            cudaGetLastError();
        }
    } drain{stream};
    cudaError_t e = cudaMemcpyAsync(d_tr, transition, sizeof(float) * S * S, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_init, initial, sizeof(float) * S, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_obs, observation, elem * rows * S, cudaMemcpyHostToDevice, stream);
    // rows no item covers keep whatever the caller's buffer holds: bring it along
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, indices, sizeof(int32_t) * rows, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return fail_cuda(e, "host->device copy");
    DecodeOptions opt;
    opt.plan = &plan;
    rc = decode_device(d_obs, batch_frames, d_tr, d_init, B, 1, S, flags, d_idx, d_post, d_ws, ws_bytes, stream, opt);
    if (rc) return rc;
    e = cudaMemcpyAsync(indices, d_idx, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && posterior)
        e = cudaMemcpyAsync(posterior, d_post, sizeof(float) * B * S, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail_cuda(e, "device->host copy / synchronize");
    drain.armed = false;
    return TORBI_B200_OK;
}

int validate_host_args(const void *observation, const int32_t *batch_frames,
                       const float *transition, const float *initial, int64_t B, int64_t T,
                       int64_t S, const int32_t *indices) {
    int rc = validate_shape(B, T, S);
    if (rc) return rc;
    if (!observation || !batch_frames || !transition || !initial || !indices)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    for (int64_t b = 0; b < B; ++b)
        if (batch_frames[b] < 1 || batch_frames[b] > T)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "batch_frames[%lld] = %d outside [1, %lld]",
                        (long long)b, batch_frames[b], (long long)T);
    return TORBI_B200_OK;
}

}  // namespace

extern "C" {

int torbi_b200_decode_host(const void *observation, const int32_t *batch_frames,
                           const float *transition, const float *initial, int64_t B, int64_t T,
                           int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                           int device) {
    g_error[0] = 0;
    int rc = validate_host_args(observation, batch_frames, transition, initial, B, T, S, indices);
    if (rc) return rc;
    return guarded([&] {
        std::vector<int32_t> items(B);
        std::iota(items.begin(), items.end(), 0);
        return decode_items_on_device(observation, batch_frames, transition, initial, items.data(), B,
                                      T, S, flags, indices, posterior, device);
    });
}

int torbi_b200_decode_host_multi(const void *observation, const int32_t *batch_frames,
                                 const float *transition, const float *initial, int64_t B,
                                 int64_t T, int64_t S, uint32_t flags, int32_t *indices,
                                 float *posterior, const int *devices, int n_devices) {
    g_error[0] = 0;
    int rc = validate_host_args(observation, batch_frames, transition, initial, B, T, S, indices);
    if (rc) return rc;
    if (!devices || n_devices < 1)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "need at least one device");
    return guarded([&] {
    const int n = (int)std::min<int64_t>(n_devices, B);

    // longest-processing-time-first: cost of an item is proportional to its
    // frame count (S^2 cells per frame, same S for all items)
    std::vector<int32_t> by_length(B);
    std::iota(by_length.begin(), by_length.end(), 0);
    std::stable_sort(by_length.begin(), by_length.end(),
                     [&](int32_t x, int32_t y) { return batch_frames[x] > batch_frames[y]; });
    std::vector<std::vector<int32_t>> shard(n);
    std::vector<int64_t> load(n, 0);
    for (int32_t item : by_length) {
        int k = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        shard[k].push_back(item);
        load[k] += batch_frames[item];
    }
    for (auto &s : shard) std::sort(s.begin(), s.end());  // contiguous runs copy together

    std::vector<int> status(n, TORBI_B200_OK);
    std::vector<std::string> message(n);
    std::vector<std::thread> workers;
    for (int k = 0; k < n; ++k) {
        workers.emplace_back([&, k] {
            status[k] = guarded([&] {
                return decode_items_on_device(observation, batch_frames, transition, initial,
                                              shard[k].data(), (int64_t)shard[k].size(), T, S,
                                              flags, indices, posterior, devices[k]);
            });
            if (status[k]) message[k] = torbi_b200_last_error();
        });
    }
    for (auto &w : workers) w.join();
    for (int k = 0; k < n; ++k)
        if (status[k]) return fail(status[k], "device %d: %s", devices[k], message[k].c_str());
    return TORBI_B200_OK;
    });
}

int torbi_b200_decode_host_packed(const void *observation, const int64_t *item_row, const int32_t *batch_frames,
                                  const float *transition, const float *initial, int64_t B, int64_t rows,
                                  int64_t S, uint32_t flags, int32_t *indices, float *posterior, int device) {
    g_error[0] = 0;
    if (!observation || !item_row || !batch_frames || !transition || !initial || !indices)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    return guarded([&] {
        return decode_packed_on_device(observation, item_row, batch_frames, transition, initial, B, rows, S, flags,
                                       indices, posterior, device);
    });
}

int torbi_b200_entropy(const void *observation, int64_t rows, int64_t S, uint32_t flags, float *entropy,
                       void *stream) {
    g_error[0] = 0;
    if (!observation || !entropy) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    if (rows < 1 || S < 1 || S > 0x7fffffff)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "rows and states must be >= 1 (got %lld, %lld)", (long long)rows, (long long)S);
    if ((flags & TORBI_B200_OBS_F16) && (flags & TORBI_B200_OBS_BF16))
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "TORBI_B200_OBS_F16 and TORBI_B200_OBS_BF16 are exclusive");
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    DeviceInfo info;
    int rc = device_info(device, &info);
    if (rc) return rc;
    e = launch_entropy(observation, (flags & TORBI_B200_OBS_F16) ? 1 : (flags & TORBI_B200_OBS_BF16) ? 2 : 0, rows, S,
                       entropy, info.sm_count, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "entropy");
    g_launches += 1;
    return TORBI_B200_OK;
}

int torbi_b200_selftest_conversion(uint64_t *mismatches) {
    g_error[0] = 0;
    if (!mismatches) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    DeviceInfo info;
    int rc = device_info(device, &info);
    if (rc) return rc;
    unsigned long long *d = nullptr, h = 0;
    e = cudaMalloc(&d, sizeof(h));
    if (e == cudaSuccess) e = cudaMemset(d, 0, sizeof(h));
    if (e == cudaSuccess) e = launch_selftest_conversion(d, info.sm_count, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(e, "conversion self-test");
    *mismatches = h;
    return TORBI_B200_OK;
}

void torbi_b200_release(void) {
    std::lock_guard<std::mutex> lock(g_arena_mutex);
    for (auto &kv : g_arenas) {
        Arena *a = kv.second;
        std::lock_guard<std::mutex> busy(a->busy);
        DeviceGuard guard(kv.first);
        if (a->base) cudaFree(a->base);
        if (a->stream) cudaStreamDestroy(a->stream);
        if (a->copy_stream) cudaStreamDestroy(a->copy_stream);
        if (a->reset_event) cudaEventDestroy(a->reset_event);
        if (a->ready_host) cudaFreeHost(a->ready_host);
        a->base = nullptr;
        a->bytes = 0;
        a->stream = nullptr;
        a->copy_stream = nullptr;
        a->reset_event = nullptr;
        a->ready_host = nullptr;
    }
}

}  // extern "C"
