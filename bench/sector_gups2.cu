// Microbenchmark v2: what bounds a random 32 B bucket probe on B200?
//  A  true-random sector reads vs table size (L2-resident .. 32 GB), load flavour
//  B  windowed-random: a block's probes fall in a window of W bytes that slides over the table
//  C  TMA-staged: cp.async.bulk a tile of the table into shared memory, probe it there
//  D  random 64 B / 128 B granules
// Each thread owns an independent xorshift stream (v1 had cross-thread address reuse).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
__device__ __forceinline__ uint64_t xs(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
struct u64x4 { uint64_t a, b, c, d; };
__device__ __forceinline__ u64x4 ld256_na(const void* p) {
    u64x4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p)); return v;
}
__device__ __forceinline__ u64x4 ld256(const void* p) {
    u64x4 v; asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p)); return v;
}
__device__ __forceinline__ u64x4 ld256_cg(const void* p) {
    u64x4 v; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p)); return v;
}
__device__ __forceinline__ uint4 ld128_na(const void* p) {
    uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}

// A/D: MODE 0 = LDG.256 nc no_allocate, 1 = LDG.256 plain, 2 = LDG.256 .cg ; GRAN = bytes read per probe (32/64/128, as GRAN/32 LDG.256)
template <int MODE, int ILP, int GRAN>
__global__ void __launch_bounds__(256) rnd(const uint8_t* __restrict__ tab, uint32_t ngran, uint32_t iters, uint64_t* sink) {
    uint64_t s = mix(blockIdx.x * 256ull + threadIdx.x + 1);
    uint64_t acc = 0;
    for (uint32_t it = 0; it < iters; it += ILP) {
        u64x4 v[ILP][GRAN / 32];
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            const uint8_t* p = tab + (uint64_t)__umulhi((uint32_t)(xs(s) >> 32), ngran) * GRAN;
#pragma unroll
            for (int q = 0; q < GRAN / 32; q++) v[j][q] = MODE == 0 ? ld256_na(p + 32 * q) : MODE == 1 ? ld256(p + 32 * q) : ld256_cg(p + 32 * q);
        }
#pragma unroll
        for (int j = 0; j < ILP; j++)
#pragma unroll
            for (int q = 0; q < GRAN / 32; q++) acc += (v[j][q].a == s) + (v[j][q].b == s) + (v[j][q].c == s) + (v[j][q].d == s);
    }
    if (acc == 0x1234567887654321ULL) sink[0] = acc;
}

// B: windowed random. Block b works on windows b, b+grid, ...; in each window every thread does `per` probes.
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) windowed(const uint8_t* __restrict__ tab, uint64_t nwin, uint32_t win_sectors, uint32_t per, uint64_t* sink) {
    uint64_t s = mix(blockIdx.x * 256ull + threadIdx.x + 1);
    uint64_t acc = 0;
    for (uint64_t w = blockIdx.x; w < nwin; w += gridDim.x) {
        const uint8_t* base = tab + w * (uint64_t)win_sectors * 32;
        for (uint32_t it = 0; it < per; it += ILP) {
            u64x4 v[ILP];
#pragma unroll
            for (int j = 0; j < ILP; j++) { const uint8_t* p = base + (uint64_t)__umulhi((uint32_t)(xs(s) >> 32), win_sectors) * 32; v[j] = MODE == 0 ? ld256_na(p) : ld256(p); }
#pragma unroll
            for (int j = 0; j < ILP; j++) acc += (v[j].a == s) + (v[j].b == s) + (v[j].c == s) + (v[j].d == s);
        }
    }
    if (acc == 0x1234567887654321ULL) sink[0] = acc;
}

// C: TMA bulk-staged tiles. 2-stage ring of TILE-byte tiles in smem; one thread issues cp.async.bulk; all threads probe.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
template <int THREADS>
__global__ void __launch_bounds__(THREADS) staged(const uint8_t* __restrict__ tab, uint64_t ntiles, uint32_t tile_bytes, uint32_t per, uint64_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    uint8_t* buf[2] = {smem, smem + tile_bytes};
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    uint64_t s = mix(blockIdx.x * (uint64_t)THREADS + threadIdx.x + 1);
    uint64_t acc = 0;
    uint32_t tile_sectors = tile_bytes / 32;
    uint64_t t = blockIdx.x; int st = 0; uint32_t ph[2] = {0, 0};
    if (threadIdx.x == 0 && t < ntiles) { mbar_expect_tx(&bar[0], tile_bytes); bulk_g2s(buf[0], tab + t * tile_bytes, tile_bytes, &bar[0]); }
    for (; t < ntiles; t += gridDim.x, st ^= 1) {
        uint64_t nt = t + gridDim.x;
        if (threadIdx.x == 0 && nt < ntiles) { mbar_expect_tx(&bar[st ^ 1], tile_bytes); bulk_g2s(buf[st ^ 1], tab + nt * tile_bytes, tile_bytes, &bar[st ^ 1]); }
        mbar_wait(&bar[st], ph[st]); ph[st] ^= 1;
        const uint8_t* base = buf[st];
        for (uint32_t it = 0; it < per; it++) {
            const uint4* p = (const uint4*)(base + (uint64_t)__umulhi((uint32_t)(xs(s) >> 32), tile_sectors) * 32);
            uint4 a = p[0], b = p[1];
            acc += (a.x == (uint32_t)s) + (a.z == (uint32_t)s) + (b.x == (uint32_t)s) + (b.z == (uint32_t)s) + (a.y & b.w & 1);
        }
        __syncthreads();   // everyone done with buf[st] before it is refilled two iterations later
    }
    if (acc == 0x1234567887654321ULL) sink[0] = acc;
}

template <typename F> float timeit(F f, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d\n", prop.name, prop.multiProcessorCount);
    uint64_t* sink; CK(cudaMalloc(&sink, 8));
    size_t maxbytes = 32ull << 30;
    uint8_t* tab; CK(cudaMalloc(&tab, maxbytes)); CK(cudaMemset(tab, 0xAB, maxbytes));
    const int blocks = 148 * 8;
    const uint64_t threads = (uint64_t)blocks * 256;
    // ---- A: true random vs size
    double sizes_mb[] = {32, 64, 256, 1024, 8192, 32768};
    for (double smb : sizes_mb) {
        size_t bytes = (size_t)(smb * (1 << 20));
        uint32_t iters = smb <= 256 ? 4096 : 1024;
        double tot = (double)iters * threads;
#define RUNA(name, MODE, ILP, GRAN) { float ms = timeit([&] { rnd<MODE, ILP, GRAN><<<blocks, 256>>>(tab, (uint32_t)(bytes / GRAN), iters, sink); }, 3); \
        printf("A size=%.0fMB %-22s %8.3f ms %7.1f Gprobe/s %7.0f GB/s\n", smb, name, ms, tot / ms * 1e-6, tot * GRAN / ms * 1e-6); }
        RUNA("na256 ilp1 g32", 0, 1, 32); RUNA("na256 ilp4 g32", 0, 4, 32); RUNA("plain256 ilp1 g32", 1, 1, 32); RUNA("plain256 ilp4 g32", 1, 4, 32);
        RUNA("cg256 ilp4 g32", 2, 4, 32);
        RUNA("na256 ilp2 g64", 0, 2, 64); RUNA("na256 ilp2 g128", 0, 2, 128); RUNA("plain ilp2 g128", 1, 2, 128);
    }
    // occupancy sweep at 8 GB
    for (int bps : {1, 2, 4, 8}) {
        int b = 148 * bps; uint32_t iters = 8192 / bps; double tot = (double)iters * b * 256;
        float ms = timeit([&] { rnd<0, 4, 32><<<b, 256>>>(tab, (uint32_t)((8ull << 30) / 32), iters, sink); }, 3);
        printf("A2 8GB blocks/SM=%d na256 ilp4: %.3f ms %.1f Gprobe/s\n", bps, ms, tot / ms * 1e-6);
        ms = timeit([&] { rnd<0, 8, 32><<<b, 256>>>(tab, (uint32_t)((8ull << 30) / 32), iters, sink); }, 3);
        printf("A2 8GB blocks/SM=%d na256 ilp8: %.3f ms %.1f Gprobe/s\n", bps, ms, tot / ms * 1e-6);
    }
    // ---- B: windowed random over 16 GB, probes per window = window sectors * density
    for (uint32_t wkb : {16, 64, 256, 1024, 4096}) {
        uint32_t wsec = wkb * 1024 / 32; uint64_t nwin = (16ull << 30) / (wkb * 1024ull);
        for (double dens : {1.0, 2.0}) {
            uint32_t per = (uint32_t)(wsec * dens / 256); if (per < 4) per = 4; per = per / 4 * 4;
            double tot = (double)per * 256 * nwin;
            float ms = timeit([&] { windowed<0, 4><<<blocks, 256>>>(tab, nwin, wsec, per, sink); }, 2);
            printf("B win=%uKB dens=%.1f na : %8.3f ms %7.1f Gprobe/s  (tab stream %.0f GB/s)\n", wkb, (double)per * 256 / wsec, ms, tot / ms * 1e-6, 16.0 * 1.0737 / ms * 1e3);
            ms = timeit([&] { windowed<1, 4><<<blocks, 256>>>(tab, nwin, wsec, per, sink); }, 2);
            printf("B win=%uKB dens=%.1f plain: %8.3f ms %7.1f Gprobe/s\n", wkb, (double)per * 256 / wsec, ms, tot / ms * 1e-6);
        }
    }
    // ---- C: TMA-staged tiles over 16 GB
    for (uint32_t tkb : {16, 32, 64, 96}) {
        uint32_t tb = tkb * 1024; uint64_t ntiles = (16ull << 30) / tb;
        for (double dens : {0.5, 1.0, 2.0}) {
            for (int thr : {256, 512}) {
                uint32_t per = (uint32_t)(tb / 32 * dens / thr); if (per < 1) per = 1;
                double tot = (double)per * thr * ntiles;
                int smem = 2 * tb; int grid;
                float ms;
                if (thr == 256) { CK(cudaFuncSetAttribute(staged<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); int occ; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, staged<256>, 256, smem)); grid = 148 * occ;
                    ms = timeit([&] { staged<256><<<grid, 256, smem>>>(tab, ntiles, tb, per, sink); }, 2); }
                else { CK(cudaFuncSetAttribute(staged<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); int occ; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, staged<512>, 512, smem)); grid = 148 * occ;
                    ms = timeit([&] { staged<512><<<grid, 512, smem>>>(tab, ntiles, tb, per, sink); }, 2); }
                printf("C tile=%uKB dens=%.2f thr=%d grid=%d: %8.3f ms %7.1f Gprobe/s  stream %.0f GB/s\n", tkb, (double)per * thr / (tb / 32), thr, grid, ms, tot / ms * 1e-6, 16.0 * 1.0737 / ms * 1e3);
            }
        }
    }
    return 0;
}
