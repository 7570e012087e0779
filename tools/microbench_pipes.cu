// Microbenchmark: sustained issue rate of the instructions the max-plus loop can
// be built from, in warp-instructions / clk / SM sub-partition (1.0 = one warp
// instruction every clock on each of the 4 schedulers).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;
constexpr int CH = 8;       // independent chains per thread
constexpr int UNROLL = 16;  // ops per chain per loop trip

enum Op { FMNMX, FMNMX3, IMIN_U32, VIMIN3_U32, FADD, FADD2, FSETP_SEL, FADD_FMNMX, FADD2_FMNMX3,
          FADD2_2IMIN, FADD2_VIMIN3, FADD2_2FMNMX, FFMA, FADD2_VIMIN3_PLUS_FMNMX,
          TWO_FADD_FMNMX3, MIX_FADD2_2FADD_2FMNMX3, FADD2_FMNMX3_SAMEOP, FADD2_SELF_FMNMX3, FADD2x2_FMNMX3x2 };

template <int OP>
__global__ void __launch_bounds__(512, 1) pipe_kernel(float *out, int trips, long long *cycles, float seed) {
    float x[CH], y[CH], z[CH];
    u64 w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        x[i] = seed * (threadIdx.x + i);
        y[i] = seed * (i + 3) - 1.f;
        z[i] = seed * (i + 7) - 2.f;
        w[i] = ((u64)__float_as_uint(y[i]) << 32) | __float_as_uint(z[i]);
    }
    const u64 cst = ((u64)__float_as_uint(seed) << 32) | __float_as_uint(-seed);
    long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                unsigned xi = __float_as_uint(x[i]);
                if (OP == FMNMX) { if (u & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i])); else asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(z[i])); }
                if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(z[i]));
                if (OP == IMIN_U32) { if (u & 1) asm volatile("min.u32 %0, %0, %1;" : "+r"(xi) : "r"(__float_as_uint(y[i]))); else asm volatile("max.u32 %0, %0, %1;" : "+r"(xi) : "r"(__float_as_uint(z[i]))); x[i] = __uint_as_float(xi); }
                if (OP == VIMIN3_U32) { asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(xi) : "r"(__float_as_uint(y[i])), "r"(__float_as_uint(z[i]))); x[i] = __uint_as_float(xi); }
                if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i]));
                if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(z[i]));
                if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % CH]));
                if (OP == FSETP_SEL) {
                    asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, %0; selp.f32 %0, %2, %0, p; }" : "+f"(x[i]) : "f"(y[i]), "f"(z[i]));
                }
                if (OP == FADD_FMNMX) {
                    asm volatile("add.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(seed));
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i]));
                }
                if (OP == TWO_FADD_FMNMX3) {
                    asm volatile("add.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(seed));
                    asm volatile("add.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(seed));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(z[i]));
                }
                if (OP == MIX_FADD2_2FADD_2FMNMX3) {   // 4 cells: one packed add, two scalar adds, two max3
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(cst));
                    asm volatile("add.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(seed));
                    asm volatile("add.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(seed));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(__uint_as_float((unsigned)w[i])), "f"(__uint_as_float((unsigned)(w[i] >> 32))));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(z[i]));
                }
                if (OP == FADD2_FMNMX3_SAMEOP) {       // max3 reads only two distinct registers
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(cst));
                    asm volatile("max.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(__uint_as_float((unsigned)w[i])));
                }
                if (OP == FADD2_SELF_FMNMX3) {         // packed add reads one pair twice
                    asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(w[i]));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(__uint_as_float((unsigned)w[i])), "f"(__uint_as_float((unsigned)(w[i] >> 32))));
                }
                if (OP == FADD2x2_FMNMX3x2) {          // the real loop's shape: 2 adds feed 2 max3 on ONE accumulator
                    u64 v0 = w[i], v1 = w[(i + 1) % CH];
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(v0) : "l"(w[i]), "l"(cst));
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(v1) : "l"(w[(i + 1) % CH]), "l"(cst));
                    w[i] = v0;
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(__uint_as_float((unsigned)v0)), "f"(__uint_as_float((unsigned)(v0 >> 32))));
                    asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(__uint_as_float((unsigned)v1)), "f"(__uint_as_float((unsigned)(v1 >> 32))));
                }
                if (OP == FADD2_FMNMX3 || OP == FADD2_2IMIN || OP == FADD2_VIMIN3 || OP == FADD2_2FMNMX || OP == FADD2_VIMIN3_PLUS_FMNMX) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(cst));
                    const u64 v = w[i];
                    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
                    if (OP == FADD2_FMNMX3)
                        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(__uint_as_float(lo)), "f"(__uint_as_float(hi)));
                    if (OP == FADD2_2FMNMX) {
                        asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(__uint_as_float(lo)));
                        asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(__uint_as_float(hi)));
                    }
                    if (OP == FADD2_2IMIN) {
                        asm volatile("min.u32 %0, %0, %1;" : "+r"(xi) : "r"(lo));
                        asm volatile("min.u32 %0, %0, %1;" : "+r"(xi) : "r"(hi));
                        x[i] = __uint_as_float(xi);
                    }
                    if (OP == FADD2_VIMIN3) {
                        asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(xi) : "r"(lo), "r"(hi));
                        x[i] = __uint_as_float(xi);
                    }
                    if (OP == FADD2_VIMIN3_PLUS_FMNMX) {
                        // alternate the two max flavours: half the cells on each
                        if (u & 1) {
                            asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(xi) : "r"(lo), "r"(hi));
                            x[i] = __uint_as_float(xi);
                        } else {
                            asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(__uint_as_float(lo)), "f"(__uint_as_float(hi)));
                        }
                    }
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + y[i] + __uint_as_float((unsigned)w[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// LDS.128 with `distinct` different 16-byte addresses per warp (consecutive lanes share)
__global__ void __launch_bounds__(512, 1) lds_kernel(float *out, int trips, long long *cycles, int distinct, int stride_floats) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x % 32;
    const int group = lane / (32 / distinct);
    const float *p = sm + group * stride_floats + (threadIdx.x / 32) * 4;
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"((unsigned)__cvta_generic_to_shared(p + ((u * 36) & 1023))));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

void run_lds(int distinct, int stride_floats, int sms) {
    const int threads = 512, trips = 2000;
    float *out; long long *cycles;
    CHECK(cudaMalloc(&out, sizeof(float) * sms * threads));
    CHECK(cudaMalloc(&cycles, sizeof(long long) * sms));
    lds_kernel<<<sms, threads, 8192 * 4>>>(out, 10, cycles, distinct, stride_floats);
    CHECK(cudaDeviceSynchronize());
    lds_kernel<<<sms, threads, 8192 * 4>>>(out, trips, cycles, distinct, stride_floats);
    CHECK(cudaDeviceSynchronize());
    long long h[512]; CHECK(cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += h[i]; cyc /= sms;
    double lds = (double)trips * 16 * (threads / 32);
    printf("LDS.128 %2d distinct addr/warp, stride %3d floats: %5.2f clk per LDS.128 per SM\n", distinct, stride_floats, cyc / lds);
    CHECK(cudaFree(out)); CHECK(cudaFree(cycles));
}

template <int OP>
void run(const char *name, int instr_per_op, int cells_per_op, int sms) {
    const int threads = 512, trips = 2000;
    float *out; long long *cycles;
    CHECK(cudaMalloc(&out, sizeof(float) * sms * threads));
    CHECK(cudaMalloc(&cycles, sizeof(long long) * sms));
    pipe_kernel<OP><<<sms, threads>>>(out, 10, cycles, 0.5f);
    CHECK(cudaDeviceSynchronize());
    pipe_kernel<OP><<<sms, threads>>>(out, trips, cycles, 0.5f);
    CHECK(cudaDeviceSynchronize());
    long long h[512]; CHECK(cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += h[i]; cyc /= sms;
    double ops_per_warp = (double)trips * UNROLL * CH;
    double warps_per_smsp = threads / 32 / 4;
    double rate = ops_per_warp * warps_per_smsp * instr_per_op / cyc;
    printf("%-28s %5.3f warp-instr/clk/SMSP", name, rate);
    if (cells_per_op) printf("   -> %6.1f cells/clk/SM", ops_per_warp * warps_per_smsp * cells_per_op * 32 * 4 / cyc);
    printf("\n");
    CHECK(cudaFree(out)); CHECK(cudaFree(cycles));
}

int main() {
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    run<FMNMX>("max.f32 (FMNMX)", 1, 0, sms);
    run<FMNMX3>("max.f32 3-input (FMNMX3)", 1, 0, sms);
    run<IMIN_U32>("min.u32 (IMNMX)", 1, 0, sms);
    run<VIMIN3_U32>("min.u32 3-input (VIMNMX3)", 1, 0, sms);
    run<FADD>("add.f32 (FADD)", 1, 0, sms);
    run<FFMA>("fma.f32 (FFMA)", 1, 0, sms);
    run<FADD2>("add.f32x2 (FADD2)", 1, 0, sms);
    run<FSETP_SEL>("setp+selp", 2, 0, sms);
    run<FADD_FMNMX>("FADD + FMNMX", 2, 1, sms);
    run<FADD2_2FMNMX>("FADD2 + 2 FMNMX", 3, 2, sms);
    run<FADD2_FMNMX3>("FADD2 + FMNMX3", 2, 2, sms);
    run<FADD2_2IMIN>("FADD2 + 2 IMNMX.u32", 3, 2, sms);
    run<FADD2_VIMIN3>("FADD2 + VIMNMX3.u32", 2, 2, sms);
    run<FADD2_VIMIN3_PLUS_FMNMX>("FADD2 + alt(VIMNMX3,FMNMX3)", 2, 2, sms);
    run<TWO_FADD_FMNMX3>("2 FADD + FMNMX3", 3, 2, sms);
    run<MIX_FADD2_2FADD_2FMNMX3>("FADD2 + 2 FADD + 2 FMNMX3", 5, 4, sms);
    run<FADD2_FMNMX3_SAMEOP>("FADD2 + FMNMX3(x,lo,lo)", 2, 2, sms);
    run<FADD2_SELF_FMNMX3>("FADD2(w,w) + FMNMX3", 2, 2, sms);
    run<FADD2x2_FMNMX3x2>("2 FADD2 + 2 FMNMX3 (1 acc)", 4, 4, sms);
    for (int d : {1, 2, 4, 8, 16, 32}) run_lds(d, 36, sms);
    run_lds(8, 32, sms);
    run_lds(8, 4, sms);
    return 0;
}
