// Ragged backtrace: restates viterbi_backtrace_trellis_cpu
// (/root/reference/torbi/csrc/viterbi.cpp:140-160) together with the
// `argmax(1).unsqueeze(1).repeat(1, T)` pre-fill of viterbi.cpp:218-221:
//   indices[b, t] = final_state[b]                      for t >= frames_b - 1
//   idx = final_state[b]; for t = frames_b-1 .. 1: idx = bp[b,t,idx]; indices[b,t-1] = idx
// One CTA per sequence: thread 0 chases the dependent uint16 loads while the
// other threads write the tail.  Per-item batch_frames, 64-bit offsets.
#include "common.cuh"
#include "kernels.cuh"

namespace torbi_b200 {

constexpr int kBacktraceThreads = 128;

__global__ void __launch_bounds__(kBacktraceThreads)
backtrace_kernel(const uint16_t *__restrict__ backptr, const int32_t *__restrict__ frames_d,
                 const int32_t *__restrict__ final_state, int32_t *__restrict__ indices,
                 int64_t T, int64_t S) {
    const int64_t b = blockIdx.x;
    const int frames = frames_d[b];
    const int last = final_state[b];
    int32_t *out = indices + b * T;
    if (threadIdx.x == 0) {
        const uint16_t *bp = backptr + b * T * S;
        int idx = last;
        for (int64_t t = frames - 1; t >= 1; --t) {
            idx = bp[t * S + idx];
            out[t - 1] = idx;
        }
    } else {
        for (int64_t t = frames - 1 + (threadIdx.x - 1); t < T; t += kBacktraceThreads - 1)
            out[t] = last;
    }
}

cudaError_t launch_backtrace(const Workspace &ws, int32_t *indices, int64_t B, int64_t T,
                             int64_t S, cudaStream_t stream) {
    backtrace_kernel<<<(unsigned)B, kBacktraceThreads, 0, stream>>>(
        ws.backptr, ws.frames, ws.final_state, indices, T, S);
    return cudaGetLastError();
}

}  // namespace torbi_b200
