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
    return carve_workspace(nullptr, B, T, S, flags).bytes;
}

}  // extern "C"

namespace {

enum Choice { kGeneric, kTiled, kWide, kSplit };

// Kernel selection: estimated time of the batch-tiled and the wide decomposition for this
// length profile (longest-first `order`), unless a flag forces one.  Also fixes a.tile_rows.
int choose_kernel(ForwardArgs &a, const int32_t *batch_frames, const std::vector<int32_t> &order,
                  uint32_t flags, Choice *out) {
    const int64_t B = a.B, S = a.S;
    std::vector<int32_t> sorted(B);
    for (int64_t i = 0; i < B; ++i) sorted[i] = batch_frames[order[i]];
    double tiled_seconds = 0.0;
    a.tile_rows = tiled_choose_rows(sorted.data(), B, S, a.sm_count, &tiled_seconds);
    if (flags & TORBI_B200_TILE_ROWS_MASK)
        a.tile_rows = 16 << (((flags & TORBI_B200_TILE_ROWS_MASK) >> 12) - 1);
    const bool can_tile = forward_tiled_supported(a), can_wide = forward_wide_supported(a);
    // the split-K small-tile kernel exists in score-history mode only, dense transitions only
    const bool can_split = forward_split_supported(a) && workspace_has_history(flags) &&
                           !(flags & (TORBI_B200_BANDED | TORBI_B200_TILED_WS));
    Choice choice = kGeneric;
    double best_seconds = 1e300;
    if (can_tile) { choice = kTiled; best_seconds = tiled_seconds; }
    if (can_split) {
        const double seconds = split_estimate_seconds(sorted.data(), B, S, a.sm_count);
        if (seconds < best_seconds) { choice = kSplit; best_seconds = seconds; }
    }
    if (can_wide && wide_estimate_seconds(sorted.data(), B, S, a.sm_count) < best_seconds) choice = kWide;
    if (flags & TORBI_B200_FORCE_SPLIT) {
        if (!can_split)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "split kernel does not support this shape / these flags");
        choice = kSplit;
    }
    if (flags & (TORBI_B200_FORCE_TILED | TORBI_B200_TILE_ROWS_MASK | TORBI_B200_BANDED)) {
        if (!can_tile) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "tiled kernel does not support this shape");
        choice = kTiled;
    }
    if (flags & TORBI_B200_FORCE_WIDE) {
        if (!can_wide) return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "wide kernel does not support this shape");
        choice = kWide;
    }
    if (flags & TORBI_B200_FORCE_GENERIC) choice = kGeneric;
    *out = choice;
    return TORBI_B200_OK;
}

const char *kernel_name(const ForwardArgs &a, Choice choice, uint32_t flags) {
    if (choice == kTiled && (flags & TORBI_B200_TILED_WS) && !(flags & TORBI_B200_BANDED) &&
        forward_tiled_ws_supported(a))
        return "forward_tiled_ws";
    if (choice == kTiled) {
        if (flags & TORBI_B200_BANDED)
            return a.tile_rows == 64 ? "forward_tiled_banded" : a.tile_rows == 32 ? "forward_tiled_banded_32" : "forward_tiled_banded_16";
        return a.tile_rows == 64 ? "forward_tiled" : a.tile_rows == 32 ? "forward_tiled_32" : "forward_tiled_16";
    }
    if (choice == kSplit) return "forward_split";
    return choice == kWide ? "forward_wide" : "forward_generic";
}

// torbi_b200_decode plus `frames_ready`: a device word the forward kernel polls
// when the observations are still being copied in (host entry); nullptr otherwise.
// `*streamed` reports whether the chosen kernel honours it.
int decode_device(const float *observation, const int32_t *batch_frames,
                  const float *transition, const float *initial, int64_t B, int64_t T,
                  int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                  void *workspace, size_t workspace_bytes, void *stream_,
                  const int32_t *frames_ready, bool *streamed) {
    g_error[0] = 0;
    int rc = validate_shape(B, T, S);
    if (rc) return rc;
    if (!observation || !batch_frames || !transition || !initial || !indices || !workspace)
        return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "NULL pointer argument");
    int64_t max_frames = 0;
    for (int64_t b = 0; b < B; ++b) {
        // the reference trusts these and reads out of bounds for 0 (viterbi.cpp:152)
        if (batch_frames[b] < 1 || batch_frames[b] > T)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "batch_frames[%lld] = %d outside [1, %lld]",
                        (long long)b, batch_frames[b], (long long)T);
        max_frames = std::max<int64_t>(max_frames, batch_frames[b]);
    }
    if (reinterpret_cast<uintptr_t>(workspace) % 256)
        return fail(TORBI_B200_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    Workspace ws = carve_workspace(workspace, B, T, S, flags);
    if (workspace_bytes < ws.bytes)
        return fail(TORBI_B200_ERR_WORKSPACE, "workspace has %zu bytes, %zu needed", workspace_bytes, ws.bytes);

    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    DeviceInfo info;
    rc = device_info(device, &info);
    if (rc) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    // lengths and the longest-first processing order go to the device
    std::vector<int32_t> order(pad_batch(B), -1);
    std::iota(order.begin(), order.begin() + B, 0);
    std::stable_sort(order.begin(), order.begin() + B,
                     [&](int32_t x, int32_t y) { return batch_frames[x] > batch_frames[y]; });
    e = cudaMemcpyAsync(ws.frames, batch_frames, sizeof(int32_t) * B, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ws.order, order.data(), sizeof(int32_t) * order.size(),
                            cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.sync, 0, sizeof(uint32_t) * kSyncWords, stream);
    if (e != cudaSuccess) return fail_cuda(e, "stage batch_frames");
    // `order` is pageable host memory: the async copy has been staged by the
    // time cudaMemcpyAsync returns, so the vector may go out of scope.

    ForwardArgs a{};
    a.observation = observation;
    a.transition = transition;
    a.initial = initial;
    a.posterior = posterior;
    a.B = B; a.T = T; a.S = S;
    a.max_frames = max_frames;
    a.flags = flags;
    a.sm_count = info.sm_count;
    a.ws = ws;
    a.frames_ready = frames_ready;
    const bool timed = flags & TORBI_B200_TIME_KERNELS;
    g_timing.valid = false;
    if (timed) {
        e = timing_prepare(device);
        if (e == cudaSuccess) e = cudaEventRecord(g_timing.ev[0], stream);
        if (e != cudaSuccess) return fail_cuda(e, "timing events");
    }

    Choice choice = kGeneric;
    rc = choose_kernel(a, batch_frames, order, flags, &choice);
    if (rc) return rc;
    const bool tiled = choice != kGeneric;               // kernels that honour frames_ready
    if (streamed) *streamed = tiled;
    g_last_kernel = kernel_name(a, choice, flags);
    bool history = false;
    if (choice == kGeneric) {
        if (frames_ready)
            return fail(TORBI_B200_ERR_INVALID_ARGUMENT, "internal: generic kernel cannot stream");
        e = launch_forward_generic(a, stream);
    } else if (choice == kWide) {
        e = launch_forward_wide(a, stream);
    } else if (choice == kSplit) {
        history = true;
        e = launch_forward_split(a, stream);
    } else if (!std::strcmp(g_last_kernel, "forward_tiled_ws")) {
        e = launch_forward_tiled_ws(a, stream);
    } else if (workspace_has_history(flags)) {
        // score-history mode (default): pure max-plus forward, argmax recomputed along the path
        history = true;
        e = launch_forward_tiled_hist(a, stream);
    } else {
        e = launch_forward_tiled(a, stream);
    }
    if (e != cudaSuccess) return fail_cuda(e, g_last_kernel);
    if (timed) cudaEventRecord(g_timing.ev[1], stream);
    g_last_backtrace = history ? "backtrace_recompute" : "backtrace";
    e = history ? launch_backtrace_recompute(ws, indices, B, T, S, stream)
                : launch_backtrace(ws, indices, B, T, S, stream);
    if (e != cudaSuccess) return fail_cuda(e, g_last_backtrace);
    if (timed) {
        cudaEventRecord(g_timing.ev[2], stream);
        g_timing.valid = true;
    }
    // kernels enqueued: (pad_transition + [analyze_band] +) forward + backtrace
    g_launches += tiled ? (choice == kTiled && (flags & TORBI_B200_BANDED) ? 4 : 3) : 2;
    return TORBI_B200_OK;
}

}  // namespace

extern "C" {

int torbi_b200_decode(const float *observation, const int32_t *batch_frames,
                      const float *transition, const float *initial, int64_t B, int64_t T,
                      int64_t S, uint32_t flags, int32_t *indices, float *posterior,
                      void *workspace, size_t workspace_bytes, void *stream) {
    return guarded([&] {
        return decode_device(observation, batch_frames, transition, initial, B, T, S, flags, indices,
                             posterior, workspace, workspace_bytes, stream, nullptr, nullptr);
    });
}

const char *torbi_b200_plan(const int32_t *batch_frames, int64_t B, int64_t T, int64_t S,
                            uint32_t flags, int sm_count) {
    g_error[0] = 0;
    if (validate_shape(B, T, S) != TORBI_B200_OK || !batch_frames || sm_count < 1) return nullptr;
    try {
    std::vector<int32_t> order(pad_batch(B), -1);
    std::iota(order.begin(), order.begin() + B, 0);
    std::stable_sort(order.begin(), order.begin() + B,
                     [&](int32_t x, int32_t y) { return batch_frames[x] > batch_frames[y]; });
    ForwardArgs a{};
    a.B = B; a.T = T; a.S = S;
    a.flags = flags;
    a.sm_count = sm_count;
    a.ws = carve_workspace(reinterpret_cast<void *>(uintptr_t(256)), B, T, S, flags);   // layout only, never dereferenced
    a.observation = reinterpret_cast<const float *>(uintptr_t(256));
    Choice choice = kGeneric;
    if (choose_kernel(a, batch_frames, order, flags, &choice) != TORBI_B200_OK) return nullptr;
    return kernel_name(a, choice, flags);
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

// Decode `count` items (given by index into the caller's batch) on one device.
int decode_items_on_device(const float *observation, const int32_t *batch_frames,
                           const float *transition, const float *initial,
                           const int32_t *items, int64_t count, int64_t T, int64_t S,
                           uint32_t flags, int32_t *indices, float *posterior, int device) {
    DeviceGuard guard(device);
    DeviceInfo info;
    int rc = device_info(device, &info);
    if (rc) return rc;
    Arena *arena = arena_for(device);
    std::lock_guard<std::mutex> lock(arena->busy);

    const size_t obs_bytes = align_up(sizeof(float) * count * T * S, 256);
    const size_t tr_bytes = align_up(sizeof(float) * S * S, 256);
    const size_t init_bytes = align_up(sizeof(float) * S, 256);
    const size_t idx_bytes = align_up(sizeof(int32_t) * count * T, 256);
    const size_t post_bytes = posterior ? align_up(sizeof(float) * count * S, 256) : 0;
    const size_t ws_bytes = carve_workspace(nullptr, count, T, S, flags).bytes;
    const size_t total = obs_bytes + tr_bytes + init_bytes + idx_bytes + post_bytes + ws_bytes + 256;

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
    char *p = static_cast<char *>(arena->base);
    float *d_obs = reinterpret_cast<float *>(p); p += obs_bytes;
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
    std::vector<int32_t> frames(count);
    for (int64_t i = 0; i < count; ++i) frames[i] = batch_frames[items[i]];

    // Large pinned batches are streamed in along T on a second stream while the
    // forward kernel already runs: step t only needs frame t (the kernel polls
    // `d_ready`).  Small, pageable or generic-kernel batches are copied up front.
    const size_t row_bytes = sizeof(float) * T * S;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, observation) == cudaSuccess &&
                        attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const bool can_tile = padded_dests(S) / kTileDests <= info.sm_count && !(flags & TORBI_B200_FORCE_GENERIC);
    const bool stream_in = pinned && can_tile && T >= 64 && row_bytes < (size_t(1) << 31) &&
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
            const float *src = observation + (int64_t)items[i] * T * S + t0 * S;
            float *dst = d_obs + i * T * S + t0 * S;
            if (chunks == 1)
                e = cudaMemcpyAsync(dst, src, row_bytes * run, cudaMemcpyHostToDevice, stream);
            else
                e = cudaMemcpy2DAsync(dst, row_bytes, src, row_bytes, sizeof(float) * (t1 - t0) * S,
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
    rc = decode_device(d_obs, frames.data(), d_tr, d_init, count, T, S, flags, d_idx, d_post, d_ws,
                       ws_bytes, stream, stream_in ? d_ready : nullptr, &streamed);
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
    return TORBI_B200_OK;
}

int validate_host_args(const float *observation, const int32_t *batch_frames,
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

int torbi_b200_decode_host(const float *observation, const int32_t *batch_frames,
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

int torbi_b200_decode_host_multi(const float *observation, const int32_t *batch_frames,
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
