// Microbenchmark: issue ceiling of the max-plus inner loop on sm_100a.
//
// Every CTA keeps a posterior tile Ps[TM][KC] and a transition tile As[TN][KC]
// in shared memory and re-runs the K loop over them (no global traffic), so
// the number printed is the ALU/LDS ceiling the real kernel cannot exceed:
//     cells / clk / SM      (1 cell = 1 fp32 add + 1 fp32 max)
// for scalar FADD+FMNMX versus packed add.rn.f32x2 + 3-input max, several
// register tiles and lane layouts.  Build/run: see tools/run_microbench.sh.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

#ifndef KUNROLL
#define KUNROLL 2
#endif
constexpr int kKUnroll = KUNROLL;
constexpr int KC = 32;        // source states per chunk
constexpr int NCH = 4;        // chunks resident in smem (a chunk's offset varies with ch: nothing is loop invariant)
constexpr int KP = NCH * KC + 4;  // padded row stride (floats): rows land on distinct banks

typedef unsigned long long u64;

// two fp32 adds in one instruction (Blackwell FADD2); operands are register pairs
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

// MODE 0: scalar add + max          MODE 1: add.f32x2 + max3
// LB x LJ lanes (LB*LJ == 32); thread tile M (b rows) x N (j rows)
template <int MODE, int M, int N, int LB, int LJ, int WARPS_B, int WARPS_J>
__global__ void __launch_bounds__(WARPS_B *WARPS_J * 32, 1)
inner_kernel(float *out, int chunks, int reps, long long *cycles) {
    constexpr int TM = WARPS_B * LB * M, TN = WARPS_J * LJ * N;
    extern __shared__ __align__(16) float smem[];
    float *Ps = smem;                 // [TM][KP]
    float *As = smem + TM * KP;       // [TN][KP]
    for (int i = threadIdx.x; i < (TM + TN) * KP; i += blockDim.x)
        smem[i] = -(float)((i * 2654435761u) % 1000u) * 0.01f;
    __syncthreads();

    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int wb = warp / WARPS_J, wj = warp % WARPS_J;
    const int lane_b = lane / LJ, lane_j = lane % LJ;
    const float *p_base = Ps + (wb * LB * M + lane_b) * KP;   // rows lane_b + LB*i
    const float *a_base = As + (wj * LJ * N + lane_j) * KP;   // rows lane_j + LJ*c

    float cur[M][N];
    int bestc[M][N];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int c = 0; c < N; ++c) { cur[i][c] = -INFINITY; bestc[i][c] = 0; }

    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll 2
        for (int ch = 0; ch < chunks; ++ch) {
            const float *pc = p_base + (ch % NCH) * KC;
            const float *ac = a_base + (ch % NCH) * KC;
            float acc[M][N];
            // first group of 4 source states peeled: acc = max3(cur, ., .) writes a
            // fresh register, so no copy of `cur` is needed for the chunk-end compare
            {
                ulonglong2 p[M], a[N];
#pragma unroll
                for (int i = 0; i < M; ++i)
                    p[i] = *reinterpret_cast<const ulonglong2 *>(pc + i * LB * KP);
#pragma unroll
                for (int c = 0; c < N; ++c)
                    a[c] = *reinterpret_cast<const ulonglong2 *>(ac + c * LJ * KP);
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        if (MODE == 0) {
                            acc[i][c] = fmaxf(cur[i][c], lo(p[i].x) + lo(a[c].x));
                            acc[i][c] = fmaxf(acc[i][c], hi(p[i].x) + hi(a[c].x));
                            acc[i][c] = fmaxf(acc[i][c], lo(p[i].y) + lo(a[c].y));
                            acc[i][c] = fmaxf(acc[i][c], hi(p[i].y) + hi(a[c].y));
                        } else if (MODE == 1) {
                            const u64 v0 = add2(p[i].x, a[c].x);
                            const u64 v1 = add2(p[i].y, a[c].y);
                            acc[i][c] = fmaxf(fmaxf(cur[i][c], lo(v0)), hi(v0));
                            acc[i][c] = fmaxf(fmaxf(acc[i][c], lo(v1)), hi(v1));
                        }
                    }
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        u64 v[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) v[c] = add2(p[i].x, a[c].x);   // N adds share p[i].x
#pragma unroll
                        for (int c = 0; c < N; ++c) acc[i][c] = fmaxf(fmaxf(cur[i][c], lo(v[c])), hi(v[c]));
#pragma unroll
                        for (int c = 0; c < N; ++c) v[c] = add2(p[i].y, a[c].y);
#pragma unroll
                        for (int c = 0; c < N; ++c) acc[i][c] = fmaxf(fmaxf(acc[i][c], lo(v[c])), hi(v[c]));
                    }
                }
            }
#pragma unroll kKUnroll
            for (int k = 4; k < KC; k += 4) {
                ulonglong2 p[M], a[N];
#pragma unroll
                for (int i = 0; i < M; ++i)
                    p[i] = *reinterpret_cast<const ulonglong2 *>(pc + i * LB * KP + k);
#pragma unroll
                for (int c = 0; c < N; ++c)
                    a[c] = *reinterpret_cast<const ulonglong2 *>(ac + c * LJ * KP + k);
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        if (MODE == 0) {
                            acc[i][c] = fmaxf(acc[i][c], lo(p[i].x) + lo(a[c].x));
                            acc[i][c] = fmaxf(acc[i][c], hi(p[i].x) + hi(a[c].x));
                            acc[i][c] = fmaxf(acc[i][c], lo(p[i].y) + lo(a[c].y));
                            acc[i][c] = fmaxf(acc[i][c], hi(p[i].y) + hi(a[c].y));
                        } else if (MODE == 1) {
                            const u64 v0 = add2(p[i].x, a[c].x);
                            const u64 v1 = add2(p[i].y, a[c].y);
                            acc[i][c] = fmaxf(fmaxf(acc[i][c], lo(v0)), hi(v0));
                            acc[i][c] = fmaxf(fmaxf(acc[i][c], lo(v1)), hi(v1));
                        }
                    }
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        u64 v[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) v[c] = add2(p[i].x, a[c].x);
#pragma unroll
                        for (int c = 0; c < N; ++c) acc[i][c] = fmaxf(fmaxf(acc[i][c], lo(v[c])), hi(v[c]));
#pragma unroll
                        for (int c = 0; c < N; ++c) v[c] = add2(p[i].y, a[c].y);
#pragma unroll
                        for (int c = 0; c < N; ++c) acc[i][c] = fmaxf(fmaxf(acc[i][c], lo(v[c])), hi(v[c]));
                    }
                }
            }
            // chunk end: strict improvement records the chunk id
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    if (acc[i][c] > cur[i][c]) bestc[i][c] = ch;
                    cur[i][c] = acc[i][c];
                }
        }
    }
    long long t1 = clock64();
    float s = 0.f; int sc = 0;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int c = 0; c < N; ++c) { s += cur[i][c]; sc += bestc[i][c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + sc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int M, int N, int LB, int LJ, int WARPS_B, int WARPS_J>
void run(const char *name, int sms) {
    constexpr int TM = WARPS_B * LB * M, TN = WARPS_J * LJ * N;
    constexpr int threads = WARPS_B * WARPS_J * 32;
    const int chunks = 45, reps = 40;
    size_t smem = sizeof(float) * (TM + TN) * KP;
    auto kernel = inner_kernel<MODE, M, N, LB, LJ, WARPS_B, WARPS_J>;
    CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float *out; long long *cycles;
    CHECK(cudaMalloc(&out, sizeof(float) * sms * threads));
    CHECK(cudaMalloc(&cycles, sizeof(long long) * sms));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    kernel<<<sms, threads, smem>>>(out, chunks, 2, cycles);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    kernel<<<sms, threads, smem>>>(out, chunks, reps, cycles);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[512]; CHECK(cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += h[i]; cyc /= sms;
    double cells = (double)TM * TN * KC * chunks * reps;
    printf("%-34s tile %3dx%-3d thr %4d smem %6zu  cells/clk/SM %7.2f  (%.3f ms, %.0f clk, %.2f GHz eff)\n",
           name, TM, TN, threads, smem, cells / cyc, ms, cyc, cyc / (ms * 1e6));
    CHECK(cudaFree(out)); CHECK(cudaFree(cycles));
}

int main() {
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    printf("SMs %d; roofline as defined in SURVEY 8d = 64 cells/clk/SM\n", sms);
    //        MODE M  N  LB LJ WB WJ
    run<0, 4, 4, 8, 4, 2, 4>("scalar 4x4 lanes8x4 warps2x4", sms);
    run<1, 4, 4, 8, 4, 2, 4>("packed 4x4 lanes8x4 warps2x4", sms);
    run<0, 4, 5, 8, 4, 2, 4>("scalar 4x5 lanes8x4 warps2x4", sms);
    run<1, 4, 5, 8, 4, 2, 4>("packed 4x5 lanes8x4 warps2x4", sms);
    run<2, 4, 5, 8, 4, 2, 4>("reuse  4x5 lanes8x4 warps2x4", sms);
    run<2, 4, 4, 8, 4, 2, 4>("reuse  4x4 lanes8x4 warps2x4", sms);
    run<2, 8, 4, 8, 4, 2, 4>("reuse  8x4 lanes8x4 warps2x4", sms);
    run<2, 4, 8, 8, 4, 2, 4>("reuse  4x8 lanes8x4 warps2x4", sms);
    run<2, 5, 4, 4, 8, 4, 2>("reuse  5x4 lanes4x8 warps4x2", sms);
    run<1, 4, 5, 4, 8, 4, 2>("packed 4x5 lanes4x8 warps4x2", sms);
    run<1, 8, 4, 4, 8, 2, 4>("packed 8x4 lanes4x8 warps2x4", sms);
    run<1, 8, 4, 8, 4, 2, 4>("packed 8x4 lanes8x4 warps2x4", sms);
    run<0, 8, 4, 8, 4, 2, 4>("scalar 8x4 lanes8x4 warps2x4", sms);
    run<1, 8, 5, 8, 4, 1, 8>("packed 8x5 lanes8x4 warps1x8", sms);
    run<1, 8, 8, 4, 8, 2, 4>("packed 8x8 lanes4x8 warps2x4", sms);
    run<1, 4, 4, 8, 4, 4, 4>("packed 4x4 lanes8x4 warps4x4", sms);
    run<1, 4, 5, 8, 4, 4, 4>("packed 4x5 lanes8x4 warps4x4", sms);
    run<1, 2, 5, 8, 4, 4, 4>("packed 2x5 lanes8x4 warps4x4", sms);
    run<1, 4, 8, 8, 4, 2, 4>("packed 4x8 lanes8x4 warps2x4", sms);
    run<1, 4, 8, 16, 2, 2, 4>("packed 4x8 lanes16x2 warps2x4", sms);
    run<1, 4, 4, 16, 2, 2, 4>("packed 4x4 lanes16x2 warps2x4", sms);
    run<1, 4, 4, 32, 1, 2, 4>("packed 4x4 lanes32x1 warps2x4", sms);
    return 0;
}
