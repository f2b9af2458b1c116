// Microbenchmark: random 32-byte-sector read throughput on B200 — the physical
// ceiling for the hash-probe kernel (one bucket = one sector per probe).
// Variants: load shape (2xLDG.128 / LDG.256 / 4 lanes x LDG.64), loads in flight
// per thread (ILP), table size, L2 fetch granularity, cache hints.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o sector_gups sector_gups.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint4 ld128(const void* p) {
    uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ uint4 ld128_plain(const void* p) {
    uint4 v; asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
struct u64x4 { uint64_t a, b, c, d; };
__device__ __forceinline__ u64x4 ld256(const void* p) {
    u64x4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p)); return v;
}

// mode 0: 2 x LDG.128 per thread per probe; mode 1: LDG.256; mode 2: plain ld (L1 allocate) 2x128
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) gups(const uint8_t* __restrict__ tab, uint32_t nbuckets, uint64_t nprobe_per_thread, uint64_t* sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    uint64_t ctr = tid * 0x9E3779B97F4A7C15ULL;
    for (uint64_t it = 0; it < nprobe_per_thread; it += ILP) {
        uint32_t b[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) { ctr += 0x9E3779B97F4A7C15ULL; b[j] = __umulhi((uint32_t)(mix(ctr) >> 32), nbuckets); }
        if (MODE == 1) {
            u64x4 v[ILP];
#pragma unroll
            for (int j = 0; j < ILP; j++) v[j] = ld256(tab + (uint64_t)b[j] * 32);
#pragma unroll
            for (int j = 0; j < ILP; j++) acc += (v[j].a == ctr) + (v[j].b == ctr) + (v[j].c == ctr) + (v[j].d == ctr);
        } else {
            uint4 v0[ILP], v1[ILP];
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                const uint8_t* p = tab + (uint64_t)b[j] * 32;
                if (MODE == 0) { v0[j] = ld128(p); v1[j] = ld128(p + 16); } else { v0[j] = ld128_plain(p); v1[j] = ld128_plain(p + 16); }
            }
#pragma unroll
            for (int j = 0; j < ILP; j++) acc += (v0[j].x == (uint32_t)ctr) + (v0[j].z == (uint32_t)ctr) + (v1[j].x == (uint32_t)ctr) + (v1[j].z == (uint32_t)ctr) + v0[j].y + v1[j].w;
        }
    }
    if (acc == 0x1234567887654321ULL) sink[0] = acc;
}

// mode 3: 2 lanes per bucket, each LDG.128 (half sector each)
template <int ILP>
__global__ void __launch_bounds__(256) gups_pair(const uint8_t* __restrict__ tab, uint32_t nbuckets, uint64_t nprobe_per_pair, uint64_t* sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t pair = tid >> 1; uint32_t half = tid & 1;
    uint64_t acc = 0;
    uint64_t ctr = pair * 0x9E3779B97F4A7C15ULL;
    for (uint64_t it = 0; it < nprobe_per_pair; it += ILP) {
        uint4 v[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) { ctr += 0x9E3779B97F4A7C15ULL; uint32_t b = __umulhi((uint32_t)(mix(ctr) >> 32), nbuckets); v[j] = ld128(tab + (uint64_t)b * 32 + half * 16); }
#pragma unroll
        for (int j = 0; j < ILP; j++) acc += (v[j].x == (uint32_t)ctr) + (v[j].z == (uint32_t)ctr) + v[j].y;
    }
    if (acc == 0x1234567887654321ULL) sink[0] = acc;
}

template <typename F> float timeit(F f, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", gran);
    uint64_t* sink; CK(cudaMalloc(&sink, 8));
    double sizes_gb[] = {1.0, 8.0, 32.0};
    for (int gi = 0; gi < 2; gi++) {
        size_t g = gi == 0 ? gran : 32;
        if (gi == 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); printf("set granularity 32: %s\n", cudaGetErrorString(e)); cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity); printf("now %zu\n", gran); }
        for (double sgb : sizes_gb) {
            size_t bytes = (size_t)(sgb * (1ull << 30));
            uint8_t* tab; CK(cudaMalloc(&tab, bytes)); CK(cudaMemset(tab, 0xAB, bytes));
            uint32_t nb = (uint32_t)(bytes / 32);
            const uint64_t total = 1ull << 31;   // probes per run
            for (int occ = 0; occ < 3; occ++) {
                int blocks = 148 * (occ == 0 ? 2 : occ == 1 ? 4 : 8);
                uint64_t threads = (uint64_t)blocks * 256;
                uint64_t per = total / threads / 16 * 16;
                double tot = (double)per * threads;
#define RUN(name, call) { float ms = timeit([&] { call; }, 3); printf("gran=%zu size=%.0fGB blocks/SM=%d %-18s %.3f ms  %.1f Gprobe/s  %.0f GB/s(32B)\n", g, sgb, blocks / 148, name, ms, tot / ms * 1e-6, tot * 32 / ms * 1e-6); }
                RUN("2xLDG128 ilp1", (gups<0, 1><<<blocks, 256>>>(tab, nb, per, sink)));
                RUN("2xLDG128 ilp4", (gups<0, 4><<<blocks, 256>>>(tab, nb, per, sink)));
                RUN("2xLDG128 ilp8", (gups<0, 8><<<blocks, 256>>>(tab, nb, per, sink)));
                RUN("LDG256 ilp1", (gups<1, 1><<<blocks, 256>>>(tab, nb, per, sink)));
                RUN("LDG256 ilp4", (gups<1, 4><<<blocks, 256>>>(tab, nb, per, sink)));
                RUN("LDG256 ilp8", (gups<1, 8><<<blocks, 256>>>(tab, nb, per, sink)));
                RUN("plain2x128 ilp4", (gups<2, 4><<<blocks, 256>>>(tab, nb, per, sink)));
                { uint64_t per2 = per; double tot2 = (double)per2 * threads / 2; float ms = timeit([&] { gups_pair<4><<<blocks, 256>>>(tab, nb, per2, sink); }, 3);
                  printf("gran=%zu size=%.0fGB blocks/SM=%d %-18s %.3f ms  %.1f Gprobe/s  %.0f GB/s(32B)\n", g, sgb, blocks / 148, "pair LDG128 ilp4", ms, tot2 / ms * 1e-6, tot2 * 32 / ms * 1e-6); }
            }
            CK(cudaFree(tab));
        }
    }
    return 0;
}
