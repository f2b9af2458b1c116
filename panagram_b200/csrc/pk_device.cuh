// Device helpers shared by the kernels of libpkanchor.so.
#ifndef PK_DEVICE_CUH
#define PK_DEVICE_CUH
#include <cstdint>

#include "pk_internal.h"

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t pk_hash32(uint64_t x) {
    x ^= x >> 32;
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    return (uint32_t)(x >> 32);
}

// reverse complement of a right-aligned 2k-bit k-mer (A=0 C=1 G=2 T=3: complement = 3-c = ~c)
__device__ __forceinline__ uint64_t pk_revcomp(uint64_t fwd, uint32_t k) {
    uint64_t x = __brevll(~fwd);
    x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
    return x >> (64 - 2 * k);
}

// canonical k-mer of the window starting at base p; false if the window holds a non-ACGT byte.
// words: 32 bases per uint64, first base in the top two bits. mask64: 1 bit per base, LSB first.
__device__ __forceinline__ bool pk_window(const uint64_t *__restrict__ words, const uint64_t *__restrict__ mask64,
                                          uint64_t p, uint32_t k, uint64_t &canon) {
    const uint64_t m0 = mask64[p >> 6], m1 = mask64[(p >> 6) + 1];
    const uint32_t t = (uint32_t)p & 63;
    const uint64_t win = (m0 >> t) | ((m1 << 1) << (63 - t));
    const uint64_t kmask = k == 64 ? ~0ull : ((1ull << k) - 1);
    if (win & kmask) return false;
    const uint64_t w0 = words[p >> 5], w1 = words[(p >> 5) + 1];
    const uint32_t s = 2 * ((uint32_t)p & 31);
    const uint64_t x = (w0 << s) | ((w1 >> 1) >> (63 - s));
    const uint64_t fwd = x >> (64 - 2 * k);
    const uint64_t rc = pk_revcomp(fwd, k);
    canon = fwd < rc ? fwd : rc;     // kmer < kmer_rev ? kmer : kmer_rev  (kmc_file.cpp:998-1001)
    return true;
}

struct u64x4 { unsigned long long a, b, c, d; };
// one 32-byte bucket in one instruction (LDG.E.256, sm_100+)
__device__ __forceinline__ u64x4 pk_ld_bucket(const unsigned long long *p) {
    u64x4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}


// same bucket through the default (L1-allocating) path: used where neighbouring probes share sectors
__device__ __forceinline__ u64x4 pk_ld_bucket_ca(const unsigned long long *p) {
    u64x4 v;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}
#endif
