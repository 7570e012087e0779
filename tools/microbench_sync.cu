// Microbenchmarks that fix the non-FP32 rooflines of the small-batch kernels (round 2):
//   1. grid-wide counter barrier (what forward_wide.cu pays once per timestep): us per barrier for
//      148 CTAs, single counter vs two-level (cluster barrier + one arrival per cluster);
//   2. hardware cluster barrier and a DSMEM push + remote mbarrier arrive round trip (cluster of 2);
//   3. L2 -> SM streaming ceiling: every CTA re-reads ITS OWN slice of an L2-resident buffer each
//      "step" (the access pattern of the streamed part of the wide kernel's transition slice and
//      of the score rows of the split / cluster kernels), LDG.128 and cp.async variants.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_sync tools/microbench_sync.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

namespace cg = cooperative_groups;

#define CHECK(x)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_release(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- 1a. single counter: __syncthreads, fence + atomicAdd, spin, __syncthreads (forward_wide.cu) ----
__global__ void barrier_single(unsigned *counter, int iters, long long *cycles) {
    unsigned target = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            while (ld_acquire(counter) < target) {}
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

// ---- 1b. red.release instead of fence + atomic ----
__global__ void barrier_red(unsigned *counter, int iters, long long *cycles) {
    unsigned target = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release(counter, 1u);
            while (ld_acquire(counter) < target) {}
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

// ---- 1c. two-level: cluster barrier, then one arrival per cluster, spin by rank 0, cluster barrier ----
__global__ void barrier_two_level(unsigned *counter, int iters, long long *cycles) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned nclusters = gridDim.x / cluster.num_blocks();
    unsigned target = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        target += nclusters;
        cluster.sync();
        if (cluster.block_rank() == 0 && threadIdx.x == 0) {
            red_release(counter, 1u);
            while (ld_acquire(counter) < target) {}
        }
        cluster.sync();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

// ---- 1d. flags instead of a counter: every CTA publishes its step number in its own 32-byte sector
//          and warp 0 of every CTA polls all of them with one coalesced load per 32 CTAs ----
__global__ void barrier_flags(unsigned *flags, int iters, long long *cycles) {
    const long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 8), "r"((unsigned)i) : "memory");
        }
        if (threadIdx.x < 32) {
            for (unsigned base = 0; base < gridDim.x; base += 32) {
                const unsigned c = base + threadIdx.x;
                bool ok;
                do {
                    ok = c >= gridDim.x || ld_acquire(flags + c * 8) >= (unsigned)i;
                } while (!__all_sync(0xffffffffu, ok));
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

// ---- 2. cluster barrier alone; DSMEM push + remote mbarrier arrive ping-pong ----
__global__ void cluster_barrier_only(int iters, long long *cycles) {
    cg::cluster_group cluster = cg::this_cluster();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) cluster.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned map_to_rank(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// every thread of CTA r stores one float4 into the peer's inbox and arrives (release.cluster) on the
// peer's mbarrier; then waits (acquire.cluster) on its own: one "exchange" of 256 x 16 B each way
__global__ void __cluster_dims__(2, 1, 1) dsmem_exchange(int iters, long long *cycles, float *sink) {
    __shared__ __align__(16) float4 inbox[2][256];
    __shared__ __align__(8) unsigned long long bar[2];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned peer = cluster.block_rank() ^ 1u;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 2; ++k)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + k)), "r"(blockDim.x));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const int k = i & 1;
        const unsigned parity = (i >> 1) & 1;
        const unsigned dst = map_to_rank(smem_u32(&inbox[k][threadIdx.x]), peer);
        const unsigned rbar = map_to_rank(smem_u32(bar + k), peer);
        const float v = acc + i;
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra DONE_%=;\n\t"
            "bra WAIT_%=;\n\t"
            "DONE_%=:\n\t}" ::"r"(smem_u32(bar + k)), "r"(parity) : "memory");
        acc += inbox[k][threadIdx.x].x;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
    if (acc == 12345.f) *sink = acc;
    cluster.sync();
}

// ---- 3. L2 -> SM streaming: CTA c re-reads slice c (slice_bytes) of the buffer `steps` times ----
template <bool kAsync>
__global__ void __launch_bounds__(512, 1) l2_stream(const float4 *buffer, size_t slice_vec4, int steps, float *sink,
                                                    long long *cycles) {
    extern __shared__ __align__(16) float4 stage[];          // kAsync: 512 threads x 4 x 16 B per pass
    const float4 *slice = buffer + (size_t)blockIdx.x * slice_vec4;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        if constexpr (kAsync) {
            for (size_t i = threadIdx.x; i < slice_vec4; i += 4 * blockDim.x) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const size_t j = i + (size_t)u * blockDim.x;
                    if (j < slice_vec4)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                                         (unsigned)__cvta_generic_to_shared(stage + u * blockDim.x + threadIdx.x)),
                                     "l"(slice + j));
                }
                asm volatile("cp.async.commit_group;");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                acc += stage[threadIdx.x].x;
            }
        } else {
            for (size_t i = threadIdx.x; i < slice_vec4; i += 8 * blockDim.x) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const size_t j = i + (size_t)u * blockDim.x;
                    v[u] = j < slice_vec4 ? __ldcg(slice + j) : make_float4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = clock64() - t0;
    if (acc == 12345.f) *sink = acc;
}

int main() {
    int device = 0, sms = 0, khz = 0;
    CHECK(cudaGetDevice(&device));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    printf("device %d: %d SMs, max clock %.0f MHz\n", device, sms, khz / 1000.0);
    unsigned *counter;
    long long *cycles;
    float *sink;
    CHECK(cudaMalloc(&counter, 4096 * 8 * sizeof(unsigned)));
    CHECK(cudaMalloc(&cycles, sizeof(long long)));
    CHECK(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const int iters = 20000;

    auto report = [&](const char *name, float ms, int n) {
        long long c = 0;
        CHECK(cudaMemcpy(&c, cycles, sizeof(c), cudaMemcpyDeviceToHost));
        printf("%-58s %8.3f us each  (%7.0f cycles)\n", name, 1e3 * ms / n, (double)c / n);
    };

    for (int threads : {256, 512}) {
        for (int which = 0; which < 4; ++which) {
            CHECK(cudaMemset(counter, 0, 4096 * 8 * sizeof(unsigned)));
            void *args[] = {(void *)&counter, (void *)&iters, (void *)&cycles};
            const void *fn = which == 0 ? (const void *)barrier_single : which == 1 ? (const void *)barrier_red
                           : which == 2 ? (const void *)barrier_two_level : (const void *)barrier_flags;
            const char *names[] = {"grid barrier: fence + atomicAdd + spin", "grid barrier: red.release + spin",
                                   "grid barrier: cluster(2) sync + red + spin + cluster sync",
                                   "grid barrier: per-CTA flags, warp-wide poll"};
            CHECK(cudaEventRecord(e0));
            if (which == 2) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(sms);
                cfg.blockDim = dim3(threads);
                cudaLaunchAttribute attr[2];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 2;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                attr[1].id = cudaLaunchAttributeCooperative;
                attr[1].val.cooperative = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 2;
                cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
                if (e != cudaSuccess) {
                    printf("cluster + cooperative launch: %s; retrying without the cooperative attribute\n", cudaGetErrorString(e));
                    cudaGetLastError();
                    cfg.numAttrs = 1;
                    CHECK(cudaLaunchKernelExC(&cfg, fn, args));
                } else {
                    printf("cluster(2) + cooperative launch of %d CTAs accepted\n", sms);
                }
            } else {
                CHECK(cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(threads), args, 0, nullptr));
            }
            CHECK(cudaEventRecord(e1));
            CHECK(cudaEventSynchronize(e1));
            float ms;
            CHECK(cudaEventElapsedTime(&ms, e0, e1));
            char label[128];
            snprintf(label, sizeof(label), "%s, %d CTAs x %d thr", names[which], sms, threads);
            report(label, ms, iters);
        }
    }
    {
        int nclusters = 0;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(sms);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CHECK(cudaFuncSetAttribute(cluster_barrier_only, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CHECK(cudaOccupancyMaxActiveClusters(&nclusters, cluster_barrier_only, &cfg));
        printf("cudaOccupancyMaxActiveClusters(cluster 2, 256 thr, 200 KB smem) = %d clusters\n", nclusters);
        for (int cs : {4, 8}) {
            attr[0].val.clusterDim.x = cs;
            cfg.gridDim = dim3(sms / cs * cs);
            cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, cluster_barrier_only, &cfg);
            printf("cudaOccupancyMaxActiveClusters(cluster %d, 256 thr, 200 KB smem) = %d clusters (%s)\n", cs, nclusters,
                   cudaGetErrorString(e));
        }
        attr[0].val.clusterDim.x = 2;
        cfg.gridDim = dim3(sms);
        cfg.dynamicSmemBytes = 0;
        void *args[] = {(void *)&iters, (void *)&cycles};
        CHECK(cudaEventRecord(e0));
        CHECK(cudaLaunchKernelExC(&cfg, (const void *)cluster_barrier_only, args));
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        report("cluster(2) barrier (arrive.release + wait.acquire), 256 thr", ms, iters);

        void *args2[] = {(void *)&iters, (void *)&cycles, (void *)&sink};
        CHECK(cudaEventRecord(e0));
        dsmem_exchange<<<sms, 256>>>(iters, cycles, sink);
        CHECK(cudaGetLastError());
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        (void)args2;
        report("DSMEM push 256 x 16 B + remote mbarrier arrive + wait", ms, iters);
    }

    // L2 -> SM streaming
    {
        const size_t total = size_t(96) << 20;
        float4 *buffer;
        CHECK(cudaMalloc(&buffer, total));
        CHECK(cudaMemset(buffer, 0, total));
        CHECK(cudaFuncSetAttribute(l2_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
        const int steps = 400;
        for (size_t slice_kb : {64, 128, 192, 278, 448}) {
            for (int async = 0; async < 2; ++async) {
                const size_t slice_vec4 = slice_kb * 1024 / 16;
                // warm the slices into L2
                if (async) l2_stream<true><<<sms, 512, 32 * 1024>>>(buffer, slice_vec4, 2, sink, cycles);
                else l2_stream<false><<<sms, 512>>>(buffer, slice_vec4, 2, sink, cycles);
                CHECK(cudaEventRecord(e0));
                if (async) l2_stream<true><<<sms, 512, 32 * 1024>>>(buffer, slice_vec4, steps, sink, cycles);
                else l2_stream<false><<<sms, 512>>>(buffer, slice_vec4, steps, sink, cycles);
                CHECK(cudaGetLastError());
                CHECK(cudaEventRecord(e1));
                CHECK(cudaEventSynchronize(e1));
                float ms;
                CHECK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = (double)slice_kb * 1024 * sms * steps;
                printf("L2 stream, %s, %3zu KB per CTA x %d CTAs (%5.1f MB set): %7.0f GB/s, %6.2f us per pass\n",
                       async ? "cp.async.cg" : "LDG.128 .cg", slice_kb, sms, slice_kb * sms / 1024.0, bytes / ms / 1e6,
                       1e3 * ms / steps);
            }
        }
        CHECK(cudaFree(buffer));
    }
    return 0;
}
