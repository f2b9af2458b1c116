// Device helpers shared by the kernels of libpkanchor.so.
#ifndef PK_DEVICE_CUH
#define PK_DEVICE_CUH
#include <cstdint>

#include "pk_internal.h"

// ------------------------------------------------------------------ sequence window
// reverse complement of a right-aligned 2k-bit k-mer (A=0 C=1 G=2 T=3: complement = 3-c = ~c)
__device__ __forceinline__ uint64_t pk_revcomp(uint64_t fwd, uint32_t k) {
    uint64_t x = __brevll(~fwd);
    x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
    return x >> (64 - 2 * k);
}

// canonical k-mer of the (valid) window starting at base p. words: 32 bases per uint64, first base in
// the top two bits.
__device__ __forceinline__ uint64_t pk_canon_at(const uint64_t *__restrict__ words, uint64_t p, uint32_t k) {
    const uint64_t w0 = words[p >> 5], w1 = words[(p >> 5) + 1];
    const uint32_t s = 2 * ((uint32_t)p & 31);
    const uint64_t x = (w0 << s) | ((w1 >> 1) >> (63 - s));
    const uint64_t fwd = x >> (64 - 2 * k);
    const uint64_t rc = pk_revcomp(fwd, k);
    return fwd < rc ? fwd : rc;      // kmer < kmer_rev ? kmer : kmer_rev  (kmc_file.cpp:998-1001)
}

// canonical k-mer of the window starting at base p; false if the window holds a non-ACGT byte.
// mask64: 1 bit per base, LSB first.
__device__ __forceinline__ bool pk_window(const uint64_t *__restrict__ words, const uint64_t *__restrict__ mask64,
                                          uint64_t p, uint32_t k, uint64_t &canon) {
    const uint64_t m0 = mask64[p >> 6], m1 = mask64[(p >> 6) + 1];
    const uint32_t t = (uint32_t)p & 63;
    const uint64_t win = (m0 >> t) | ((m1 << 1) << (63 - t));
    const uint64_t kmask = k == 64 ? ~0ull : ((1ull << k) - 1);
    if (win & kmask) return false;
    canon = pk_canon_at(words, p, k);
    return true;
}

// ------------------------------------------------------------------ hashing
__device__ __forceinline__ uint32_t pk_hash64(uint64_t x) {
    x ^= x >> 32;
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    return (uint32_t)(x >> 32);
}
__device__ __forceinline__ uint32_t pk_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// The 32-bit hash every table is indexed by (bucket = mulhi(hash, n_buckets)).
//  S32: the slot stores only the low 28 key bits (+ 4 displacement bits), so the remaining eb = 2k-28 high
//       key bits must be recoverable from the home bucket: they are XOR-ed with a hash of the low bits
//       (uniform, and a bijection for fixed low bits) and placed in the TOP eb bits of the hash; with
//       n_buckets >= 2^eb two k-mers with equal low bits never share a home bucket.
//  S64: the slot stores the whole k-mer, any good hash would do; the hash has the same structure with a 52-bit
//       remainder (ebu = max(0, 2k-52) top bits carry the high key bits), because the GROUP tables (below) store
//       52 key bits per slot and are indexed by the same hash as the per-genome tables.
#define PK_U_REM_BITS 52
#define PK_U_REM_MASK 0x000FFFFFFFFFFFFFull
__device__ __forceinline__ uint32_t pk_key_hash(uint64_t canon, const PkKeySpec ks) {
    if (ks.fmt == PK_FMT_S64) {
        const uint32_t m = pk_hash64(canon & PK_U_REM_MASK);
        const uint32_t ebu = 2 * ks.k > PK_U_REM_BITS ? 2 * ks.k - PK_U_REM_BITS : 0;
        if (ebu == 0) return m;
        const uint32_t hi = (uint32_t)(canon >> PK_U_REM_BITS);
        return ((hi ^ ((m * 0x9E3779B1u) >> (32 - ebu))) << (32 - ebu)) | (m >> ebu);
    }
    const uint32_t lo = (uint32_t)canon & PK_S32_REM_MASK, m = pk_mix32(lo);
    if (ks.eb == 0) return m;
    const uint32_t hi = (uint32_t)(canon >> PK_S32_REM_BITS);
    return ((hi ^ ((m * 0x9E3779B1u) >> (32 - ks.eb))) << (32 - ks.eb)) | (m >> ks.eb);
}

// ------------------------------------------------------------------ buckets (32 B = one sector)
struct u64x4 { unsigned long long a, b, c, d; };
// one 32-byte bucket in one instruction (LDG.E.256, sm_100+), streaming (no L1 allocation)
__device__ __forceinline__ u64x4 pk_ld_bucket(const void *p) {
    u64x4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}
// same through the default (L1-allocating) path: used where neighbouring probes share sectors
__device__ __forceinline__ u64x4 pk_ld_bucket_ca(const void *p) {
    u64x4 v;
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    return v;
}

// what a lookup compares against in a bucket at displacement r from the key's home bucket
template <int FMT> __device__ __forceinline__ uint64_t pk_target(uint64_t canon, uint32_t r) {
    return FMT == PK_FMT_S64 ? canon : (uint64_t)((((uint32_t)canon & PK_S32_REM_MASK) << PK_S32_DISP_BITS) | r);
}
template <int FMT> __device__ __forceinline__ bool pk_bucket_hit(const u64x4 &v, uint64_t target) {
    if (FMT == PK_FMT_S64) return v.a == target || v.b == target || v.c == target || v.d == target;
    const uint32_t t = (uint32_t)target;
    return (uint32_t)v.a == t || (uint32_t)(v.a >> 32) == t || (uint32_t)v.b == t || (uint32_t)(v.b >> 32) == t ||
           (uint32_t)v.c == t || (uint32_t)(v.c >> 32) == t || (uint32_t)v.d == t || (uint32_t)(v.d >> 32) == t;
}
// slots fill in order (an insert claims the first free slot it meets), so a bucket is full iff its last slot is taken
template <int FMT> __device__ __forceinline__ bool pk_bucket_full(const u64x4 &v) {
    return FMT == PK_FMT_S64 ? v.d != PK_EMPTY : (uint32_t)(v.d >> 32) != PK_EMPTY32;
}
// how far a key may sit from its home bucket
template <int FMT> __device__ __forceinline__ uint32_t pk_max_disp(uint32_t n_buckets) {
    return FMT == PK_FMT_S64 ? n_buckets - 1 : (n_buckets - 1 < PK_S32_MAX_DISP ? n_buckets - 1 : PK_S32_MAX_DISP);
}

// the stash: keys of S32 tables that found no room within PK_S32_MAX_DISP buckets of home
__device__ __forceinline__ unsigned long long pk_stash_entry(uint32_t g, uint64_t canon) { return ((unsigned long long)g << 48) | canon; }
__device__ __forceinline__ bool pk_stash_contains(const PkKeySpec ks, uint32_t g, uint64_t canon) {
    if (*(volatile unsigned int *)ks.stash_n == 0) return false;
    const unsigned long long e = pk_stash_entry(g, canon);
    uint32_t i = pk_hash64(e) & (PK_STASH_SLOTS - 1);
    for (uint32_t t = 0; t < PK_STASH_SLOTS; t++) {
        const unsigned long long cur = ks.stash[i];
        if (cur == e) return true;
        if (cur == PK_EMPTY) return false;
        i = (i + 1) & (PK_STASH_SLOTS - 1);
    }
    return false;
}
// 0 = already there, 1 = inserted, 3 = stash full
__device__ __forceinline__ int pk_stash_insert(const PkKeySpec ks, uint32_t g, uint64_t canon) {
    const unsigned long long e = pk_stash_entry(g, canon);
    uint32_t i = pk_hash64(e) & (PK_STASH_SLOTS - 1);
    for (uint32_t t = 0; t < PK_STASH_SLOTS / 2; t++) {
        const unsigned long long cur = *(volatile unsigned long long *)(ks.stash + i);
        if (cur == e) return 0;
        if (cur == PK_EMPTY) {
            const unsigned long long old = atomicCAS(ks.stash + i, PK_EMPTY, e);
            if (old == PK_EMPTY) { atomicAdd(ks.stash_n, 1u); return 1; }
            if (old == e) return 0;
        }
        i = (i + 1) & (PK_STASH_SLOTS - 1);
    }
    return 3;
}

// full lookup (home bucket + walk-on), used off the hot path. g = local genome index (stash key).
template <int FMT> __device__ __forceinline__ bool pk_lookup(const PkTable t, uint64_t canon, uint32_t h, uint32_t g, const PkKeySpec ks) {
    uint32_t b = __umulhi(h, t.n_buckets);
    const uint32_t maxd = pk_max_disp<FMT>(t.n_buckets);
    for (uint32_t r = 0;; r++) {
        const u64x4 v = pk_ld_bucket_ca((const char *)t.slots + 32ull * b);
        if (pk_bucket_hit<FMT>(v, pk_target<FMT>(canon, r))) return true;
        if (!pk_bucket_full<FMT>(v)) return false;
        if (r == maxd) return FMT == PK_FMT_S32 ? pk_stash_contains(ks, g, canon) : false;
        b = b + 1 == t.n_buckets ? 0 : b + 1;
    }
}
// ------------------------------------------------------------------ group ("union") tables
// One table per group of <= 8 genomes, derived from the per-genome tables at finalize: slot = 64 bits =
// [membership mask 8][key remainder 52][displacement 4], 4 slots per 32 B bucket, EMPTY = ~0 (displacement 15
// never occurs). ONE probe answers 8 genomes — the structure of the reference's own merged "bitvec" database
// (counter = 32-genome bit vector, index.py:407-426), as a bucketed hash table. Indexed by the same hash as the
// per-genome tables; for 2k <= 52 the remainder is the whole k-mer, above that the top 2k-52 key bits are implied
// by the home bucket exactly as in the S32 format. Keys that find 15 full buckets in a row go to the engine-wide
// stash under their genome, like S32 keys.
#define PK_U_KEY_MASK 0x00FFFFFFFFFFFFFFull
#define PK_U_MASK_SHIFT 56
#define PK_U_GROUP 8u
__device__ __forceinline__ uint64_t pk_u_key(uint64_t canon, uint32_t r) { return ((canon & PK_U_REM_MASK) << PK_S32_DISP_BITS) | r; }
__device__ __forceinline__ uint32_t pk_u_max_disp(uint32_t n_buckets) { return n_buckets - 1 < PK_S32_MAX_DISP ? n_buckets - 1 : PK_S32_MAX_DISP; }
// membership mask of `key56` in a bucket (0 when absent)
__device__ __forceinline__ uint32_t pk_u_bucket_mask(const u64x4 &v, uint64_t key56) {
    uint32_t m = 0;
    if ((v.a & PK_U_KEY_MASK) == key56) m |= (uint32_t)(v.a >> PK_U_MASK_SHIFT);
    if ((v.b & PK_U_KEY_MASK) == key56) m |= (uint32_t)(v.b >> PK_U_MASK_SHIFT);
    if ((v.c & PK_U_KEY_MASK) == key56) m |= (uint32_t)(v.c >> PK_U_MASK_SHIFT);
    if ((v.d & PK_U_KEY_MASK) == key56) m |= (uint32_t)(v.d >> PK_U_MASK_SHIFT);
    return m;
}
// full lookup in a group table; g0 = local index of the group's first genome, ng = genomes in the group
__device__ __forceinline__ uint32_t pk_u_lookup(const PkTable t, uint64_t canon, uint32_t h, uint32_t g0, uint32_t ng, const PkKeySpec ks) {
    uint32_t b = __umulhi(h, t.n_buckets);
    const uint32_t maxd = pk_u_max_disp(t.n_buckets);
    for (uint32_t r = 0;; r++) {
        const u64x4 v = pk_ld_bucket_ca((const char *)t.slots + 32ull * b);
        const uint32_t m = pk_u_bucket_mask(v, pk_u_key(canon, r));
        if (m) return m;
        if (v.d == PK_EMPTY) return 0;
        if (r == maxd) {
            uint32_t sm = 0;
            for (uint32_t g = 0; g < ng; g++) sm |= (uint32_t)pk_stash_contains(ks, g0 + g, canon) << g;
            return sm;
        }
        b = b + 1 == t.n_buckets ? 0 : b + 1;
    }
}
// ---- G32: the compact group-table format for short k-mers --------------------------------------------------
// slot = 32 bits = [membership mask 8][key remainder 20][displacement 4], 8 slots per 32 B bucket, EMPTY = ~0.
// The other eb = 2k - 20 key bits are implied by the home bucket, as in the S32 per-genome format: the group
// tables are then indexed by their OWN hash (pk_g32_hash: top eb bits = hi ^ f(lo20)), which needs n_buckets >=
// 2^eb — true for any genome-scale table at the default k = 21 (eb = 22: >= 4 Mi buckets = 128 MB). Half the bytes
// per key of the 64-bit slots: the probe kernel, which streams its table once per launch, reads half as much.
// PkTable.fmt says which format a group table has (all group tables of an engine share it); PkKeySpec.ghash != 0
// tells the kernels of a launch that positions are hashed with pk_g32_hash.
#define PK_G32_REM_BITS 20
#define PK_G32_REM_MASK 0x000FFFFFu
#define PK_G32_KEY_MASK 0x00FFFFFFu
#define PK_G32_MASK_SHIFT 24
#define PK_TFMT_G64 0u
#define PK_TFMT_G32 1u
__device__ __forceinline__ uint32_t pk_g32_hash(uint64_t canon, uint32_t k) {
    const uint32_t lo = (uint32_t)canon & PK_G32_REM_MASK, m = pk_mix32(lo);
    const uint32_t eb = 2 * k > PK_G32_REM_BITS ? 2 * k - PK_G32_REM_BITS : 0;
    if (eb == 0) return m;
    const uint32_t hi = (uint32_t)(canon >> PK_G32_REM_BITS);
    return ((hi ^ ((m * 0x9E3779B1u) >> (32 - eb))) << (32 - eb)) | (eb < 32 ? m >> eb : 0u);
}
// the hash the positions of a probe launch are partitioned and probed by
__device__ __forceinline__ uint32_t pk_probe_hash(uint64_t canon, const PkKeySpec ks) {
    return ks.ghash ? pk_g32_hash(canon, ks.k) : pk_key_hash(canon, ks);
}
__device__ __forceinline__ uint32_t pk_g32_key(uint64_t canon, uint32_t r) { return (((uint32_t)canon & PK_G32_REM_MASK) << PK_S32_DISP_BITS) | r; }
__device__ __forceinline__ uint32_t pk_g32_bucket_mask(const u64x4 &v, uint32_t key24) {
    const uint32_t s[8] = {(uint32_t)v.a, (uint32_t)(v.a >> 32), (uint32_t)v.b, (uint32_t)(v.b >> 32),
                           (uint32_t)v.c, (uint32_t)(v.c >> 32), (uint32_t)v.d, (uint32_t)(v.d >> 32)};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if ((s[i] & PK_G32_KEY_MASK) == key24) m |= s[i] >> PK_G32_MASK_SHIFT;
    return m;
}
__device__ __forceinline__ uint32_t pk_g32_lookup(const PkTable t, uint64_t canon, uint32_t h, uint32_t g0, uint32_t ng, const PkKeySpec ks) {
    uint32_t b = __umulhi(h, t.n_buckets);
    const uint32_t maxd = pk_u_max_disp(t.n_buckets);
    for (uint32_t r = 0;; r++) {
        const u64x4 v = pk_ld_bucket_ca((const char *)t.slots + 32ull * b);
        const uint32_t m = pk_g32_bucket_mask(v, pk_g32_key(canon, r));
        if (m) return m;
        if ((uint32_t)(v.d >> 32) == PK_EMPTY32) return 0;
        if (r == maxd) {
            uint32_t sm = 0;
            for (uint32_t g = 0; g < ng; g++) sm |= (uint32_t)pk_stash_contains(ks, g0 + g, canon) << g;
            return sm;
        }
        b = b + 1 == t.n_buckets ? 0 : b + 1;
    }
}
// as pk_u_insert
__device__ __forceinline__ int pk_g32_insert(const PkTable t, uint64_t canon, uint32_t h, uint32_t bit) {
    uint32_t b = __umulhi(h, t.n_buckets);
    const uint32_t maxd = pk_u_max_disp(t.n_buckets);
    const uint32_t mbit = 1u << (PK_G32_MASK_SHIFT + bit);
    for (uint32_t r = 0; r <= maxd; r++) {
        const uint32_t key24 = pk_g32_key(canon, r);
        uint32_t *slot = (uint32_t *)t.slots + 8ull * b;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t cur = *(volatile uint32_t *)(slot + i);
            if (cur == PK_EMPTY32) {
                cur = atomicCAS(slot + i, PK_EMPTY32, key24 | mbit);
                if (cur == PK_EMPTY32) return r == 0 ? 2 : 3;
            }
            if ((cur & PK_G32_KEY_MASK) == key24) return (atomicOr(slot + i, mbit) & mbit) ? 0 : 1;
        }
        b = b + 1 == t.n_buckets ? 0 : b + 1;
    }
    return 4;
}
// Slot q (bucket q / slots-per-bucket) of a group table back to (canonical k-mer, its hash, membership mask): the
// remainder comes from the slot, the key bits the slot does not store from the HOME bucket (bucket - displacement),
// by inverting the table's hash — (X << sh) | low for the one X that maps to the home bucket (n_buckets >= 2^eb
// makes it unique). false for an empty slot.
__device__ __forceinline__ bool pk_group_slot_decode(const PkTable t, const PkKeySpec ks, uint64_t q, uint64_t &canon, uint32_t &h, uint32_t &mask) {
    uint32_t b, disp, eb, m;
    uint64_t rem;
    if (t.fmt == PK_TFMT_G32) {
        const uint32_t v = ((const uint32_t *)t.slots)[q];
        if (v == PK_EMPTY32) return false;
        b = (uint32_t)(q >> 3); disp = v & 15u; mask = v >> PK_G32_MASK_SHIFT;
        rem = (v >> PK_S32_DISP_BITS) & PK_G32_REM_MASK;
        eb = 2 * ks.k > PK_G32_REM_BITS ? 2 * ks.k - PK_G32_REM_BITS : 0;
        m = pk_mix32((uint32_t)rem);
        if (eb == 0) { canon = rem; h = m; return true; }
    } else {
        const unsigned long long v = t.slots[q];
        if (v == PK_EMPTY) return false;
        b = (uint32_t)(q >> 2); disp = (uint32_t)v & 15u; mask = (uint32_t)(v >> PK_U_MASK_SHIFT);
        rem = (v >> PK_S32_DISP_BITS) & PK_U_REM_MASK;
        if (2 * ks.k <= PK_U_REM_BITS) { canon = rem; h = pk_key_hash(canon, ks); return true; }
        // 2k > 52: only the S64 per-genome format reaches here (S32 covers k <= 24), hash = pk_key_hash's S64 branch
        eb = 2 * ks.k - PK_U_REM_BITS;
        m = pk_hash64(rem);
    }
    const uint32_t home = b >= disp ? b - disp : b + t.n_buckets - disp;
    const uint32_t sh = 32 - eb, low = eb < 32 ? m >> eb : 0u;
    const uint32_t hmin = (uint32_t)((((uint64_t)home << 32) + t.n_buckets - 1) / t.n_buckets);    // smallest hash of the home bucket
    uint32_t X = hmin >> sh;
    h = (X << sh) | low;
    if (__umulhi(h, t.n_buckets) != home) { X = (X + 1) & (eb < 32 ? (1u << eb) - 1 : 0xffffffffu); h = (X << sh) | low; }
    const uint32_t hi = X ^ ((m * 0x9E3779B1u) >> sh);
    canon = ((uint64_t)hi << (t.fmt == PK_TFMT_G32 ? PK_G32_REM_BITS : PK_U_REM_BITS)) | rem;
    return true;
}

// format-dispatching forms (t.fmt)
__device__ __forceinline__ uint32_t pk_group_lookup(const PkTable t, uint64_t canon, uint32_t h, uint32_t g0, uint32_t ng, const PkKeySpec ks);
__device__ __forceinline__ int pk_group_insert(const PkTable t, uint64_t canon, uint32_t h, uint32_t bit);

// set genome bit `bit` (0..7) of `canon` in group table t. 0 = bit was already set, 1 = set in an existing slot,
// 2 = new slot in the home bucket, 3 = new slot in a later bucket, 4 = no room within 15 buckets (caller stashes)
__device__ __forceinline__ int pk_u_insert(const PkTable t, uint64_t canon, uint32_t h, uint32_t bit) {
    uint32_t b = __umulhi(h, t.n_buckets);
    const uint32_t maxd = pk_u_max_disp(t.n_buckets);
    const unsigned long long mbit = 1ull << (PK_U_MASK_SHIFT + bit);
    for (uint32_t r = 0; r <= maxd; r++) {
        const unsigned long long key56 = pk_u_key(canon, r);
        unsigned long long *slot = t.slots + 4ull * b;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            unsigned long long cur = *(volatile unsigned long long *)(slot + i);
            if (cur == PK_EMPTY) {
                cur = atomicCAS(slot + i, PK_EMPTY, key56 | mbit);
                if (cur == PK_EMPTY) return r == 0 ? 2 : 3;
            }
            if ((cur & PK_U_KEY_MASK) == key56) return (atomicOr(slot + i, mbit) & mbit) ? 0 : 1;
        }
        b = b + 1 == t.n_buckets ? 0 : b + 1;
    }
    return 4;
}
__device__ __forceinline__ uint32_t pk_group_lookup(const PkTable t, uint64_t canon, uint32_t h, uint32_t g0, uint32_t ng, const PkKeySpec ks) {
    return t.fmt == PK_TFMT_G32 ? pk_g32_lookup(t, canon, h, g0, ng, ks) : pk_u_lookup(t, canon, h, g0, ng, ks);
}
__device__ __forceinline__ int pk_group_insert(const PkTable t, uint64_t canon, uint32_t h, uint32_t bit) {
    return t.fmt == PK_TFMT_G32 ? pk_g32_insert(t, canon, h, bit) : pk_u_insert(t, canon, h, bit);
}
#endif
