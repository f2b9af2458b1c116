// sm_100a kernels of the anchoring path. Integer/byte work bounded by HBM; no tensor cores.
//
//   pack      ASCII -> 2-bit words + invalid mask         (replaces CKmerAPI::num_codes mapping, kmer_api.h:264-275)
//   insert    canonical k-mers -> per-genome bucket table  (replaces the sorted suffix arrays of a KMC DB)
//   decode    KMC suffix records -> k-mers -> insert       (replaces CKMCFile record decoding, kmc_file.cpp:421-490)
//   probe     sliding canonical k-mer + membership probe   (replaces GetCountersForRead_kmc1_both_strands +
//                                                           count_for_kmer_kmc1 + BinarySearch, kmc_file.cpp:905-1027,1321-1399)
//   reduce    popcount histograms, column sums, low-res    (replaces cpp/anchor.cpp:160,169-190; index.py:1048-1051)
//   interleave all-gathered column planes -> rows          (byte interleave of cpp/anchor.cpp:154-165 across ranks)
#include <cuda_runtime.h>

#include <algorithm>

#include "pk_internal.h"
#include "pk_device.cuh"
#include "pk_gather.cuh"

// ------------------------------------------------------------------ fill / pack
__global__ void __launch_bounds__(256) fill_empty_kernel(ulonglong2 *slots2, uint64_t n2) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n2; i += stride)
        slots2[i] = make_ulonglong2(PK_EMPTY, PK_EMPTY);
}
void pk_launch_fill_empty(unsigned long long *slots, uint64_t n_slots, pk_stream_t s) {
    const uint64_t n2 = n_slots / 2;   // n_slots is a multiple of 4
    const unsigned grid = (unsigned)(n2 / 256 + 1 < 148u * 16 ? n2 / 256 + 1 : 148u * 16);
    fill_empty_kernel<<<grid, 256, 0, s>>>((ulonglong2 *)slots, n2);
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, uint64_t len, uint64_t n_words,
                                                   uint64_t *__restrict__ words, uint32_t *__restrict__ mask) {
    const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint64_t base = w * 32;
    uint64_t bits = 0;
    uint32_t inv = 0;
    uint32_t chunk[8];
    if (base + 32 <= len && (((uintptr_t)ascii) & 15) == 0) {
        const uint4 a = *(const uint4 *)(ascii + base), b = *(const uint4 *)(ascii + base + 16);
        chunk[0] = a.x; chunk[1] = a.y; chunk[2] = a.z; chunk[3] = a.w;
        chunk[4] = b.x; chunk[5] = b.y; chunk[6] = b.z; chunk[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint64_t i = base + 4 * q + b;
                v |= (uint32_t)(i < len ? ascii[i] : (uint8_t)'N') << (8 * b);
            }
            chunk[q] = v;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t c = (chunk[q] >> (8 * b)) & 0xff;
            const uint32_t code = ((c >> 1) ^ (c >> 2)) & 3;          // A,a->0 C,c->1 G,g->2 T,t->3
            const uint32_t u = c | 0x20;
            const bool ok = (u == 'a') | (u == 'c') | (u == 'g') | (u == 't');
            const int i = 4 * q + b;
            bits |= (uint64_t)(ok ? code : 0) << (62 - 2 * i);
            inv |= (ok ? 0u : 1u) << i;
        }
    }
    words[w] = bits;
    mask[w] = inv;
}
void pk_launch_pack(const uint8_t *d_ascii, uint64_t len, uint64_t n_words, uint64_t *d_words, uint32_t *d_mask, pk_stream_t s) {
    if (!n_words) return;
    pack_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>(d_ascii, len, n_words, d_words, d_mask);
}

// ------------------------------------------------------------------ insert
// Bucket-granular linear probing: a key lives in the first bucket, starting at its home bucket,
// that had a free slot when it was inserted. Slots only ever go EMPTY -> key and fill in order, so a
// lookup that sees a free last slot in a bucket knows the key is in no later bucket.
// returns 0 = already present, 1 = inserted in home bucket, 2 = inserted in a later bucket, 3 = no room
__device__ __forceinline__ int pk_table_insert(const PkTable t, unsigned long long key, const PkKeySpec ks, uint32_t g) {
    const uint32_t home = __umulhi(pk_key_hash(key, ks), t.n_buckets);
    uint32_t b = home;
    if (ks.fmt == PK_FMT_S64) {
        for (uint32_t tries = 0; tries < t.n_buckets; ++tries) {
            volatile unsigned long long *slot = t.slots + 4ull * b;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned long long cur = slot[i];
                if (cur == key) return 0;
                if (cur == PK_EMPTY) {
                    const unsigned long long old = atomicCAS((unsigned long long *)slot + i, PK_EMPTY, key);
                    if (old == PK_EMPTY) return b == home ? 1 : 2;
                    if (old == key) return 0;
                }
            }
            b = b + 1 == t.n_buckets ? 0 : b + 1;
        }
        return 3;
    }
    const uint32_t maxd = pk_max_disp<PK_FMT_S32>(t.n_buckets);
    for (uint32_t r = 0; r <= maxd; ++r) {
        const uint32_t want = (uint32_t)pk_target<PK_FMT_S32>(key, r);
        volatile uint32_t *slot = (uint32_t *)t.slots + 8ull * b;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t cur = slot[i];
            if (cur == want) return 0;
            if (cur == PK_EMPTY32) {
                const uint32_t old = atomicCAS((uint32_t *)slot + i, PK_EMPTY32, want);
                if (old == PK_EMPTY32) return r == 0 ? 1 : 2;
                if (old == want) return 0;
            }
        }
        b = b + 1 == t.n_buckets ? 0 : b + 1;
    }
    return pk_stash_insert(ks, g, key);     // 15 consecutive full buckets: engine-wide stash
}

__device__ __forceinline__ void pk_flush_counts(unsigned long long *counters, uint32_t ins, uint32_t ovf, uint32_t fail) {
    ins = __reduce_add_sync(0xffffffffu, ins);
    ovf = __reduce_add_sync(0xffffffffu, ovf);
    fail = __reduce_add_sync(0xffffffffu, fail);
    if ((threadIdx.x & 31) == 0) {
        if (ins) atomicAdd(counters, (unsigned long long)ins);
        if (ovf) atomicAdd(counters + 1, (unsigned long long)ovf);
        if (fail) atomicAdd(counters + 2, (unsigned long long)fail);
    }
}

__global__ void __launch_bounds__(256) insert_seq_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ mask64,
                                                         uint64_t n, PkKeySpec ks, PkTable t, uint32_t g, unsigned long long *counters) {
    uint32_t ins = 0, ovf = 0, fail = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n; p += stride) {
        uint64_t canon;
        if (!pk_window(words, mask64, p, ks.k, canon)) continue;
        const int r = pk_table_insert(t, canon, ks, g);
        ins += r == 1 || r == 2; ovf += r == 2; fail += r == 3;
    }
    pk_flush_counts(counters, ins, ovf, fail);
}
static unsigned grid_for(uint64_t n, unsigned per_block = 256, unsigned cap = 148u * 32) {
    const uint64_t g = (n + per_block - 1) / per_block;
    return (unsigned)(g < 1 ? 1 : g > cap ? cap : g);
}
void pk_launch_insert_seq(const uint64_t *d_words, const uint32_t *d_mask, uint64_t n, PkKeySpec ks, PkTable t, uint32_t g,
                          unsigned long long *d_counters, pk_stream_t s) {
    if (!n) return;
    insert_seq_kernel<<<grid_for(n), 256, 0, s>>>(d_words, (const uint64_t *)d_mask, n, ks, t, g, d_counters);
}

__global__ void __launch_bounds__(256) insert_keys_kernel(const uint64_t *__restrict__ keys, uint64_t n, PkKeySpec ks, PkTable t,
                                                          uint32_t g, unsigned long long *counters) {
    uint32_t ins = 0, ovf = 0, fail = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const int r = pk_table_insert(t, keys[i], ks, g);
        ins += r == 1 || r == 2; ovf += r == 2; fail += r == 3;
    }
    pk_flush_counts(counters, ins, ovf, fail);
}
void pk_launch_insert_keys(const uint64_t *d_keys, uint64_t n, PkKeySpec ks, PkTable t, uint32_t g, unsigned long long *d_counters,
                           pk_stream_t s) {
    if (!n) return;
    insert_keys_kernel<<<grid_for(n), 256, 0, s>>>(d_keys, n, ks, t, g, d_counters);
}

// KMC record r (global index) belongs to LUT slot j iff lut[j] <= r < lut[j+1]; its prefix is
// j mod 4^lut (KMC2 keeps one LUT per signature bin); k-mer = prefix << 8*suf_size | suffix
// (suffix bytes most significant first), counter little-endian (kmc_file.cpp:421-490).
__global__ void __launch_bounds__(256) decode_insert_kernel(PkDecodeArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const uint64_t r = a.rec0 + i;
        uint64_t lo = 0, hi = a.n_lut_slots;        // lut[lo] <= r < lut[hi]
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (a.d_lut[mid] <= r) lo = mid; else hi = mid;
        }
        const uint64_t prefix = lo % a.single_lut;
        const uint8_t *rec = a.d_recs + i * a.rec_size;
        uint64_t suf = 0;
        for (uint32_t b = 0; b < a.suf_size; b++) suf = (suf << 8) | rec[b];
        const uint32_t sbits = 8 * a.suf_size;
        const uint64_t key = (sbits >= 64 ? 0 : prefix << sbits) | suf;
        uint64_t cnt = 1;
        if (a.counter_size) {
            cnt = 0;
            for (uint32_t b = 0; b < a.counter_size; b++) cnt |= (uint64_t)rec[a.suf_size + b] << (8 * b);
            if (cnt < a.min_count || cnt > a.max_count) continue;   // kmc_file.cpp:487 / :1396
        }
        if (!a.bitvec) {
            const int rr = pk_table_insert(a.d_tables[a.local_genome], key, a.ks, a.local_genome);
            unsigned long long *c = a.d_counters + 3 * a.local_genome;
            if (rr == 1 || rr == 2) atomicAdd(c, 1ull);
            if (rr == 2) atomicAdd(c + 1, 1ull);
            if (rr == 3) atomicAdd(c + 2, 1ull);
        } else {
            uint32_t bits = (uint32_t)cnt;
            while (bits) {
                const uint32_t j = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint32_t g = a.first_genome + j;
                if (g < a.gbegin || g >= a.gend) continue;
                const int rr = pk_table_insert(a.d_tables[g - a.gbegin], key, a.ks, g - a.gbegin);
                unsigned long long *c = a.d_counters + 3 * (g - a.gbegin);
                if (rr == 1 || rr == 2) atomicAdd(c, 1ull);
                if (rr == 2) atomicAdd(c + 1, 1ull);
                if (rr == 3) atomicAdd(c + 2, 1ull);
            }
        }
    }
}
void pk_launch_decode_insert(const PkDecodeArgs &a, pk_stream_t s) {
    if (!a.n) return;
    decode_insert_kernel<<<grid_for(a.n), 256, 0, s>>>(a);
}

// per-bit popcount of the counters of a bitvec database (counter bit j = genome first + j): how many k-mers each of
// its <= 32 genomes has, so that every table is sized for its own genome
__global__ void __launch_bounds__(256) count_bits_kernel(const uint8_t *__restrict__ recs, uint64_t n, uint32_t rec_size, uint32_t suf_size,
                                                         uint32_t counter_size, uint64_t min_count, uint64_t max_count,
                                                         unsigned long long *__restrict__ bits) {
    __shared__ unsigned int sh[32];
    if (threadIdx.x < 32) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x; i0 < n; i0 += stride) {
        const uint64_t i = i0 + threadIdx.x;
        uint64_t cnt = 0;
        if (i < n) {
            const uint8_t *rec = recs + i * rec_size + suf_size;
            for (uint32_t b = 0; b < counter_size; b++) cnt |= (uint64_t)rec[b] << (8 * b);
            if (cnt < min_count || cnt > max_count) cnt = 0;
        }
        const uint32_t c32 = (uint32_t)cnt;
        uint32_t mine = 0;
        for (uint32_t j = 0; j < 32; j++) {
            const uint32_t m = __ballot_sync(0xffffffffu, (c32 >> j) & 1);
            if (lane == j) mine = __popc(m);
        }
        if (mine) atomicAdd(&sh[lane], mine);
    }
    __syncthreads();
    if (threadIdx.x < 32 && sh[threadIdx.x]) atomicAdd(&bits[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
void pk_launch_count_bits(const uint8_t *d_recs, uint64_t n, uint32_t rec_size, uint32_t suf_size, uint32_t counter_size, uint64_t min_count,
                          uint64_t max_count, unsigned long long *d_bits, pk_stream_t s) {
    if (!n) return;
    count_bits_kernel<<<grid_for(n), 256, 0, s>>>(d_recs, n, rec_size, suf_size, counter_size, min_count, max_count, d_bits);
}

// ------------------------------------------------------------------ group tables (derived at finalize)
// Merge one per-genome table into its group table: every stored k-mer is re-derived from its slot (S64: the slot
// is the k-mer; S32: low 28 bits from the slot, the other bits from the HOME bucket = bucket - displacement, by
// inverting pk_key_hash: the hash is (X << (32-eb)) | (mix32(lo) >> eb) for the one X that maps to the home bucket)
// and its genome bit is set in the group table. n_src_buckets < src.n_buckets restricts the merge to a prefix of
// the hash range (used to estimate the number of distinct k-mers of a group from 1/64 of it, with hshift = 6).
// counters: [0] new slots created, [1] bits set, [2] keys sent to the stash, [3] failures (stash full)
__global__ void __launch_bounds__(256) union_merge_kernel(PkTable src, uint32_t n_src_buckets, uint32_t hshift, PkKeySpec ks, PkTable dst,
                                                          uint32_t bit, uint32_t g_local, int use_stash, unsigned long long *counters) {
    const uint32_t spb = ks.fmt == PK_FMT_S32 ? 8 : 4;                 // slots per bucket
    const uint64_t total = (uint64_t)n_src_buckets * spb;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t created = 0, set = 0, stashed = 0, failed = 0;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < total; q += stride) {
        const uint32_t b = (uint32_t)(q / spb), i = (uint32_t)(q % spb);
        uint64_t canon;
        uint32_t h;
        if (ks.fmt == PK_FMT_S32) {
            const uint32_t v = ((const uint32_t *)src.slots)[q];
            if (v == PK_EMPTY32) continue;
            const uint32_t lo = v >> PK_S32_DISP_BITS, disp = v & 15u;
            const uint32_t home = b >= disp ? b - disp : b + src.n_buckets - disp;
            const uint32_t m = pk_mix32(lo);
            if (ks.eb == 0) {
                h = m; canon = lo;
            } else {
                const uint32_t sh = 32 - ks.eb, low = m >> ks.eb;
                // smallest hash of the home bucket: ceil(home * 2^32 / n_buckets)
                const uint32_t hmin = (uint32_t)((((uint64_t)home << 32) + src.n_buckets - 1) / src.n_buckets);
                uint32_t X = hmin >> sh;
                h = (X << sh) | low;
                if (__umulhi(h, src.n_buckets) != home) { X = (X + 1) & ((1u << ks.eb) - 1); h = (X << sh) | low; }
                const uint32_t hi = X ^ ((m * 0x9E3779B1u) >> sh);
                canon = ((uint64_t)hi << PK_S32_REM_BITS) | lo;
            }
        } else {
            canon = src.slots[q];
            if (canon == PK_EMPTY) continue;
            h = pk_key_hash(canon, ks);
        }
        (void)i;
        // hshift: a 2^-hshift prefix of the (per-genome) hash range fills all of dst. A G32 table has its own hash: the
        // prefix of the source buckets is then simply a uniform sample of the keys (the same one in every genome)
        const uint32_t hd = dst.fmt == PK_TFMT_G32 ? pk_g32_hash(canon, ks.k) : h << hshift;
        const int r = pk_group_insert(dst, canon, hd, bit);
        created += r >= 2 && r <= 3; set += r >= 1 && r <= 3;
        if (r == 4) {
            // counted by union_merge_stash_kernel, which runs after all table merges and sees every stash entry
            if (use_stash) { const int sr = pk_stash_insert(ks, g_local, canon); failed += sr == 3; }
            else failed++;
        }
    }
    created = __reduce_add_sync(0xffffffffu, created); set = __reduce_add_sync(0xffffffffu, set);
    stashed = __reduce_add_sync(0xffffffffu, stashed); failed = __reduce_add_sync(0xffffffffu, failed);
    if ((threadIdx.x & 31) == 0) {
        if (created) atomicAdd(counters, (unsigned long long)created);
        if (set) atomicAdd(counters + 1, (unsigned long long)set);
        if (stashed) atomicAdd(counters + 2, (unsigned long long)stashed);
        if (failed) atomicAdd(counters + 3, (unsigned long long)failed);
    }
}
// keys of the group's genomes that live in the engine-wide stash (S32 tables only): same merge
__global__ void __launch_bounds__(256) union_merge_stash_kernel(PkKeySpec ks, PkTable dst, uint32_t g0, uint32_t ng, unsigned long long *counters) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= PK_STASH_SLOTS) return;
    const unsigned long long e = ks.stash[i];
    if (e == PK_EMPTY) return;
    const uint32_t g = (uint32_t)(e >> 48);
    if (g < g0 || g >= g0 + ng) return;
    const uint64_t canon = e & 0x0000FFFFFFFFFFFFull;
    const int r = pk_group_insert(dst, canon, dst.fmt == PK_TFMT_G32 ? pk_g32_hash(canon, ks.k) : pk_key_hash(canon, ks), g - g0);
    if (r == 2 || r == 3) atomicAdd(counters, 1ull);
    if (r >= 1 && r <= 3) atomicAdd(counters + 1, 1ull);
    if (r == 4) atomicAdd(counters + 2, 1ull);                // stays in the stash, found there by pk_u_lookup
}
void pk_launch_union_merge_stash(PkKeySpec ks, PkTable dst, uint32_t g0, uint32_t ng, unsigned long long *d_counters, pk_stream_t s) {
    union_merge_stash_kernel<<<PK_STASH_SLOTS / 256, 256, 0, s>>>(ks, dst, g0, ng, d_counters);
}
void pk_launch_union_merge(PkTable src, uint32_t n_src_buckets, uint32_t hshift, PkKeySpec ks, PkTable dst, uint32_t bit, uint32_t g_local,
                           int use_stash, unsigned long long *d_counters, pk_stream_t s) {
    if (!n_src_buckets) return;
    const uint64_t total = (uint64_t)n_src_buckets * (ks.fmt == PK_FMT_S32 ? 8 : 4);
    union_merge_kernel<<<grid_for(total), 256, 0, s>>>(src, n_src_buckets, hshift, ks, dst, bit, g_local, use_stash, d_counters);
}

// A uniform sample of the k-mers of a group table with their membership masks: every k-mer whose hash is below
// `hmax` (a hash of the k-mer alone, so the SAME k-mers are sampled in every group table of every engine of a
// sharded run: a FracMinHash sketch; take_all lists everything). Feeds the pairwise
// Jaccard / Mash distances of genome_dist.tsv (workflow/Snakefile:124-149 runs `mash sketch -s 10000` +
// `mash triangle` for that). out_keys/out_tags[cap]; tag = (group << 8) | mask; *n_out counts every hit (may exceed cap).
__global__ void __launch_bounds__(256) sample_group_kernel(PkTable t, PkKeySpec ks, uint32_t group, uint32_t hmax, int take_all,
                                                           uint64_t n_slots, unsigned long long *__restrict__ out_keys,
                                                           uint32_t *__restrict__ out_tags, uint64_t cap, unsigned long long *__restrict__ n_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n_slots; q += stride) {
        uint64_t canon; uint32_t h, mask;
        if (!pk_group_slot_decode(t, ks, q, canon, h, mask)) continue;
        if (!take_all && pk_hash64(canon) >= hmax) continue;        // a hash of the k-mer alone: the same sample whatever the slot format
        const unsigned long long at = atomicAdd(n_out, 1ull);
        if (at < cap) { out_keys[at] = canon; out_tags[at] = (group << 8) | mask; }
    }
}
// the engine-wide stash (k-mers that found no room within 15 buckets of home): same sample rule, one bit per entry
__global__ void __launch_bounds__(256) sample_stash_kernel(PkKeySpec ks, int g32, uint32_t hmax, int take_all, unsigned long long *__restrict__ out_keys,
                                                           uint32_t *__restrict__ out_tags, uint64_t cap, unsigned long long *__restrict__ n_out) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= PK_STASH_SLOTS || !ks.stash) return;
    const unsigned long long e = ks.stash[i];
    if (e == PK_EMPTY) return;
    const uint32_t g = (uint32_t)(e >> 48);
    const uint64_t canon = e & 0x0000FFFFFFFFFFFFull;
    (void)g32;
    if (!take_all && pk_hash64(canon) >= hmax) return;
    const unsigned long long at = atomicAdd(n_out, 1ull);
    if (at < cap) { out_keys[at] = canon; out_tags[at] = ((g / PK_U_GROUP) << 8) | (1u << (g % PK_U_GROUP)); }
}
void pk_launch_sample_group(PkTable t, PkKeySpec ks, uint32_t group, uint32_t hmax, int take_all, unsigned long long *d_keys, uint32_t *d_tags,
                            uint64_t cap, unsigned long long *d_n, pk_stream_t s) {
    const uint64_t n_slots = (uint64_t)t.n_buckets * (t.fmt == PK_TFMT_G32 ? 8 : 4);
    sample_group_kernel<<<grid_for(n_slots), 256, 0, s>>>(t, ks, group, hmax, take_all, n_slots, d_keys, d_tags, cap, d_n);
}
void pk_launch_sample_stash(PkKeySpec ks, int g32, uint32_t hmax, int take_all, unsigned long long *d_keys, uint32_t *d_tags, uint64_t cap,
                            unsigned long long *d_n, pk_stream_t s) {
    sample_stash_kernel<<<PK_STASH_SLOTS / 256, 256, 0, s>>>(ks, g32, hmax, take_all, d_keys, d_tags, cap, d_n);
}

// ------------------------------------------------------------------ probe (direct)
// One thread per position; for each local genome one 32 B bucket load (LDG.256), U loads in
// flight per thread. Row bits accumulate in a register and are stored once per 32 genomes.
template <int U, int FMT>
__global__ void __launch_bounds__(256) probe_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ mask64,
                                                    uint64_t p0, uint64_t n, PkKeySpec ks,
                                                    const PkTable *__restrict__ tables, uint32_t n_local,
                                                    uint8_t *__restrict__ rows, uint32_t row_stride, uint32_t col_offset) {
    const uint64_t i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    uint64_t canon = 0;
    const bool valid = pk_window(words, mask64, p0 + i, ks.k, canon);
    const uint32_t h = pk_key_hash(canon, ks);
    const uint64_t target = pk_target<FMT>(canon, 0);
    const uint32_t nbl = (n_local + 7) / 8;
    uint8_t *dst = rows + i * row_stride + col_offset;
    const bool al4 = ((row_stride | col_offset) & 3) == 0;
    for (uint32_t g0 = 0; g0 < n_local; g0 += 32) {
        uint32_t bits = 0;
        const uint32_t ng = n_local - g0 < 32 ? n_local - g0 : 32;
        if (valid) {
            for (uint32_t j0 = 0; j0 < ng; j0 += U) {
                u64x4 v[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (j0 + u < ng) {
                        const PkTable t = tables[g0 + j0 + u];
                        v[u] = pk_ld_bucket((const char *)t.slots + 32ull * __umulhi(h, t.n_buckets));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (j0 + u < ng) {
                        bool hit = pk_bucket_hit<FMT>(v[u], target);
                        if (!hit && pk_bucket_full<FMT>(v[u])) hit = pk_lookup<FMT>(tables[g0 + j0 + u], canon, h, g0 + j0 + u, ks);
                        bits |= (uint32_t)hit << (j0 + u);
                    }
                }
            }
        }
        const uint32_t nb = nbl - g0 / 8 < 4 ? nbl - g0 / 8 : 4;
        if (nb == 4 && al4) {
            *(uint32_t *)(dst + g0 / 8) = bits;
        } else {
            for (uint32_t q = 0; q < nb; q++) dst[g0 / 8 + q] = (uint8_t)(bits >> (8 * q));
        }
    }
}
void pk_launch_probe(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, PkKeySpec ks,
                     const PkTable *d_tables, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride,
                     uint32_t col_offset, pk_stream_t s) {
    if (!n) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    const uint64_t *m64 = (const uint64_t *)d_mask;
    if (ks.fmt == PK_FMT_S32) {
        if (n_local <= 4) probe_kernel<4, PK_FMT_S32><<<grid, 256, 0, s>>>(d_words, m64, p0, n, ks, d_tables, n_local, d_rows, row_stride, col_offset);
        else probe_kernel<8, PK_FMT_S32><<<grid, 256, 0, s>>>(d_words, m64, p0, n, ks, d_tables, n_local, d_rows, row_stride, col_offset);
    } else {
        if (n_local <= 4) probe_kernel<4, PK_FMT_S64><<<grid, 256, 0, s>>>(d_words, m64, p0, n, ks, d_tables, n_local, d_rows, row_stride, col_offset);
        else probe_kernel<8, PK_FMT_S64><<<grid, 256, 0, s>>>(d_words, m64, p0, n, ks, d_tables, n_local, d_rows, row_stride, col_offset);
    }
}

// Direct probe out of the GROUP tables (one bucket read answers 8 genomes): used for small batches whenever the
// group tables exist, and the only direct path once the per-genome tables have been freed (group_only).
// Row byte u = membership mask of group u.
__global__ void __launch_bounds__(256) probe_group_kernel(const uint64_t *__restrict__ words, const uint64_t *__restrict__ mask64,
                                                          uint64_t p0, uint64_t n, PkKeySpec ks, const PkTable *__restrict__ utables,
                                                          uint32_t n_local, uint8_t *__restrict__ rows, uint32_t row_stride, uint32_t col_offset) {
    const uint64_t i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    uint64_t canon = 0;
    const bool valid = pk_window(words, mask64, p0 + i, ks.k, canon);
    const uint32_t h = pk_probe_hash(canon, ks);
    uint8_t *dst = rows + i * row_stride + col_offset;
    const uint32_t n_groups = (n_local + PK_U_GROUP - 1) / PK_U_GROUP;
    for (uint32_t u = 0; u < n_groups; u++) {
        uint32_t m = 0;
        if (valid) m = pk_group_lookup(utables[u], canon, h, PK_U_GROUP * u, min(PK_U_GROUP, n_local - PK_U_GROUP * u), ks);
        dst[u] = (uint8_t)m;
    }
}
void pk_launch_probe_group(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, PkKeySpec ks,
                           const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, pk_stream_t s) {
    if (!n) return;
    probe_group_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_words, (const uint64_t *)d_mask, p0, n, ks, d_utables, n_local, d_rows,
                                                                   row_stride, col_offset);
}
// (hash, position) items of the partitioned probe answered out of the group tables through L1/L2, rows written
// directly (such a position appears in no un-permute list). Two uses: the SPILL list (counts == NULL: one flat
// list of *flat_total items — items beyond a region's capacity, repeat-rich sequence), and whole REGIONS when a
// partition's table window is too large to be staged in shared memory piece by piece (short batches against
// large tables) — the probes of a region still fall inside one window of each table, so they hit L2.
__global__ void __launch_bounds__(256) items_group_kernel(const uint2 *__restrict__ buf, const uint32_t *__restrict__ counts,
                                                          const unsigned long long *__restrict__ flat_total, uint32_t n_regions, uint64_t cap,
                                                          const uint64_t *__restrict__ words, uint64_t p0, PkKeySpec ks,
                                                          const PkTable *__restrict__ utables, uint32_t n_local, uint8_t *__restrict__ rows,
                                                          uint32_t row_stride, uint32_t col_offset, uint32_t *__restrict__ list4,
                                                          uint32_t *__restrict__ list_cursor, uint32_t cursor_stride, uint32_t out_shift,
                                                          int compact) {
    const uint32_t n_groups = (n_local + PK_U_GROUP - 1) / PK_U_GROUP;
    auto one = [&](uint2 it) {
        if (compact) it.y &= 0x0FFFFFFFu;        // compact item (pk_partition.cu): 28-bit position; the hash is re-derived
        const uint64_t canon = pk_canon_at(words, p0 + it.y, ks.k);
        if (compact) it.x = pk_probe_hash(canon, ks);
        if (list4) {        // one-byte rows, fine un-permute lists: (position in bin) << 8 | bits, one slot per position of the bin
            const uint32_t m = pk_group_lookup(utables[0], canon, it.x, 0, min(PK_U_GROUP, n_local), ks);
            const uint32_t bin = it.y >> out_shift;
            const uint32_t slot = atomicAdd(&list_cursor[bin * cursor_stride], 1u);
            list4[((uint64_t)bin << out_shift) + slot] = ((it.y & ((1u << out_shift) - 1)) << 8) | (m & 0xffu);
            return;
        }
        uint8_t *dst = rows + (uint64_t)it.y * row_stride + col_offset;
        for (uint32_t u = 0; u < n_groups; u++)
            dst[u] = (uint8_t)pk_group_lookup(utables[u], canon, it.x, PK_U_GROUP * u, min(PK_U_GROUP, n_local - PK_U_GROUP * u), ks);
    };
    if (!counts) {
        const unsigned long long n = min(*flat_total, (unsigned long long)cap);
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n; q += stride) one(buf[q]);
    } else {
        for (uint64_t r = blockIdx.x; r < n_regions; r += gridDim.x) {
            const uint32_t cnt = (uint32_t)min((uint64_t)counts[r], cap);
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) one(buf[r * cap + i]);
        }
    }
}
void pk_launch_items_group(const void *d_buf, const uint32_t *d_counts, const unsigned long long *d_flat_total, uint32_t n_regions, uint64_t cap,
                           const uint64_t *d_words, uint64_t p0, PkKeySpec ks, const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows,
                           uint32_t row_stride, uint32_t col_offset, uint32_t *list4, uint32_t *list_cursor, uint32_t cursor_stride,
                           uint32_t out_shift, int compact, pk_stream_t s) {
    const unsigned grid = d_counts ? (n_regions < 148u * 8 ? (n_regions ? n_regions : 1) : 148u * 8) : 148u * 2;
    items_group_kernel<<<grid, 256, 0, s>>>((const uint2 *)d_buf, d_counts, d_flat_total, n_regions, cap, d_words, p0, ks, d_utables, n_local, d_rows,
                                            row_stride, col_offset, list4, list_cursor, cursor_stride, out_shift, compact);
}

// ------------------------------------------------------------------ reduce
// rows: n full rows (n_cols bit columns, ceil(n_cols/8) bytes used of row_stride). For chromosome
// position p = p_first + i:  hist[(p / binlen) * (n_cols+1) + popcount(row)]++ ; col_sums[g] += bit g ;
// rows_low[(p/step - ceil(p_first/step))] = row  when p % step == 0.
#define PK_RED_ITEMS 16
// IDX = uint32_t when every position fits 32 bits (the common case: 32-bit divides), else uint64_t.
template <typename IDX>
__global__ void __launch_bounds__(256) reduce_kernel(const uint8_t *__restrict__ rows, uint32_t row_stride, uint32_t n_cols,
                                                     uint64_t p_first64, uint64_t n64, uint64_t binlen64,
                                                     unsigned long long *__restrict__ hist, unsigned long long *__restrict__ col_sums) {
    extern __shared__ unsigned int sh[];
    const uint32_t n_words = (n_cols + 31) / 32, nbytes = (n_cols + 7) / 8;
    unsigned int *sh_hist = sh;                    // [n_cols + 1]
    unsigned int *sh_col = sh + n_cols + 1;        // [n_words * 32]
    for (uint32_t q = threadIdx.x; q < n_cols + 1 + n_words * 32; q += blockDim.x) sh[q] = 0;
    __syncthreads();
    const IDX p_first = (IDX)p_first64, n = (IDX)n64, binlen = (IDX)binlen64;
    const IDX base = (IDX)blockIdx.x * (IDX)(256 * PK_RED_ITEMS);
    const IDX bin0 = binlen ? (p_first + base) / binlen : 0;
    const uint32_t lane = threadIdx.x & 31;
    const bool vec1 = nbytes == 1 && row_stride == 1;
    for (int it = 0; it < PK_RED_ITEMS; it++) {
        if (base + (IDX)it * 256 + (threadIdx.x & ~31u) >= n) break;   // whole warp out of range (warp-uniform)
        const IDX i = base + (IDX)it * 256 + threadIdx.x;
        const bool active = i < n;
        const uint8_t *row = rows + (uint64_t)i * row_stride;
        uint32_t pc = 0;
        for (uint32_t wd = 0; wd < n_words; wd++) {
            uint32_t w = 0;
            if (active) {
                if (vec1) w = row[0];
                else for (uint32_t b = 0; b < 4 && 4 * wd + b < nbytes; b++) w |= (uint32_t)row[4 * wd + b] << (8 * b);
                if (wd == n_words - 1 && (n_cols & 31)) w &= (1u << (n_cols & 31)) - 1;
            }
            pc += __popc(w);
            if (col_sums) {
                const uint32_t nbits = min(32u, n_cols - 32 * wd);
                uint32_t mine = 0;
                for (uint32_t j = 0; j < nbits; j++) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (w >> j) & 1);
                    if (lane == j) mine = __popc(m);
                }
                if (mine) atomicAdd(&sh_col[wd * 32 + lane], mine);
            }
        }
        if (hist && binlen) {
            const IDX bin = (p_first + i) / binlen;
            const uint32_t key = active ? (uint32_t)((bin - bin0) << 16 | pc) : 0xffffffffu;   // a block spans <= 4096 positions, so bin - bin0 < 2^16
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            if (active && lane == (uint32_t)(__ffs(peers) - 1)) {
                const uint32_t cnt = __popc(peers);
                if (bin == bin0) atomicAdd(&sh_hist[pc], cnt);
                else atomicAdd(&hist[(uint64_t)bin * (n_cols + 1) + pc], (unsigned long long)cnt);
            }
        }
    }
    __syncthreads();
    if (hist && binlen)
        for (uint32_t q = threadIdx.x; q < n_cols + 1; q += blockDim.x)
            if (sh_hist[q]) atomicAdd(&hist[(uint64_t)bin0 * (n_cols + 1) + q], (unsigned long long)sh_hist[q]);
    if (col_sums)
        for (uint32_t q = threadIdx.x; q < n_cols; q += blockDim.x)
            if (sh_col[q]) atomicAdd(&col_sums[q], (unsigned long long)sh_col[q]);
}

// One-byte rows (N <= 8, the 8-genomes-per-GPU layout) with bins of >= R1_BLOCK_ROWS positions: every thread
// takes 16 consecutive rows per step as one 16-byte load; per-genome column sums come from masked popcounts
// of the four words, the popcount histogram from a SWAR byte popcount accumulated in a packed register
// (9 fields of 7 bits, one per popcount value), one packed register per bin the block touches (<= 2).
// 0.17 ms -> see profiles/ for 27 M rows (the generic kernel spends a ballot per column and a match_any per row).
#define R1_STEPS 4
#define R1_BLOCK_ROWS (256 * 16 * R1_STEPS)
__global__ void __launch_bounds__(256) reduce1_kernel(const uint8_t *__restrict__ rows, uint32_t n_cols, uint64_t p_first, uint64_t n,
                                                      uint64_t binlen, unsigned long long *__restrict__ hist,
                                                      unsigned long long *__restrict__ col_sums) {
    __shared__ unsigned int sh_hist[2][9], sh_col[8];
    if (threadIdx.x < 18) (&sh_hist[0][0])[threadIdx.x] = 0;
    if (threadIdx.x < 8) sh_col[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * R1_BLOCK_ROWS;
    const uint64_t bin0 = binlen ? (p_first + base) / binlen : 0;
    // rows [base, base + split) belong to bin0, the rest of the block's rows to bin0 + 1
    const uint64_t split64 = binlen ? (bin0 + 1) * binlen - (p_first + base) : ~0ull;
    const uint32_t split = split64 > R1_BLOCK_ROWS ? R1_BLOCK_ROWS : (uint32_t)split64;
    const uint32_t cmask = n_cols >= 8 ? 0xffu : ((1u << n_cols) - 1);
    unsigned long long pa = 0, pb = 0;
    uint32_t col[8];
#pragma unroll
    for (int j = 0; j < 8; j++) col[j] = 0;
    for (int st = 0; st < R1_STEPS; st++) {
        const uint32_t r0 = (st * 256 + threadIdx.x) * 16;          // first of this thread's 16 rows, relative to base
        if (base + r0 >= n) break;
        uint32_t w[4];
        if (base + r0 + 16 <= n) {
            const uint4 v = *(const uint4 *)(rows + base + r0);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[q] = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint64_t i = base + r0 + 4 * q + b;
                    w[q] |= (uint32_t)(i < n ? rows[i] : 0) << (8 * b);      // rows past the end count as popcount 0: removed below
                }
            }
        }
        const uint32_t m4 = cmask * 0x01010101u;
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] &= m4;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t mj = 0x01010101u << j;
            col[j] += __popc(w[0] & mj) + __popc(w[1] & mj) + __popc(w[2] & mj) + __popc(w[3] & mj);
        }
        const uint32_t valid = (uint32_t)min((uint64_t)16, n - (base + r0));
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t x = w[q];
            x = x - ((x >> 1) & 0x55555555u);
            x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
            x = (x + (x >> 4)) & 0x0f0f0f0fu;                           // per-byte popcounts 0..8
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t idx = 4 * q + b;
                if (idx < valid) {
                    const unsigned long long one = 1ull << (7 * ((x >> (8 * b)) & 0xf));
                    if (r0 + idx < split) pa += one; else pb += one;
                }
            }
        }
    }
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int f = 0; f < 9; f++) {
        const uint32_t a = __reduce_add_sync(0xffffffffu, (uint32_t)(pa >> (7 * f)) & 127u);
        const uint32_t b = __reduce_add_sync(0xffffffffu, (uint32_t)(pb >> (7 * f)) & 127u);
        if (lane == 0) { if (a) atomicAdd(&sh_hist[0][f], a); if (b) atomicAdd(&sh_hist[1][f], b); }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t c = __reduce_add_sync(0xffffffffu, col[j]);
        if (lane == 0 && c) atomicAdd(&sh_col[j], c);
    }
    __syncthreads();
    if (hist && binlen && threadIdx.x < 18) {
        const uint32_t which = threadIdx.x / 9, f = threadIdx.x % 9;
        const unsigned int v = sh_hist[which][f];
        if (v && f <= n_cols) atomicAdd(&hist[(bin0 + which) * (n_cols + 1) + f], (unsigned long long)v);
    }
    if (col_sums && threadIdx.x >= 32 && threadIdx.x < 32 + n_cols && sh_col[threadIdx.x - 32])
        atomicAdd(&col_sums[threadIdx.x - 32], (unsigned long long)sh_col[threadIdx.x - 32]);
}

// Rows of W = 2, 4, 8 or 16 bytes (9..128 genomes; the genome-sharded slices and configs[2..4]) with bins of >= 2^16
// positions. The generic kernel spends a ballot per COLUMN and a match_any per row (4.6 ms for 150 M 8-byte rows);
// here a thread takes 16-byte vectors (16 / W rows), coalesced, and counts columns in bit-sliced registers: the four
// words of a vector are added into 4-bit fields (one register per bit offset 0..3 of a nibble: (w >> o) & 0x11111111),
// every <= 15 adds the nibble fields are widened into byte fields, every <= 17 widenings... the block ends: the
// byte fields are summed over the warp and added to shared, then global, counters. Row popcounts go to a
// [2 bins][N + 1] shared histogram, run-length merged per thread (consecutive rows mostly share their popcount).
template <int W> struct RwCfg {
    static constexpr int R = 16 / W;                          // rows per vector
    static constexpr int WC = W >= 4 ? W / 4 : 1;             // word classes: which 32 columns a word of the vector holds
    static constexpr int FL = 15 / (4 / WC);                  // vectors per nibble-field lifetime
    static constexpr int NF = (65536 / 256 / (FL * R)) < 17 ? (65536 / 256 / (FL * R)) : 17;      // widenings per block
    static constexpr int ITERS = FL * NF;
    static constexpr uint32_t BLOCK_ROWS = 256u * ITERS * R;
};
#define RW_MIN_BINLEN 65536u
template <int W>
__global__ void __launch_bounds__(256) reducew_kernel(const uint8_t *__restrict__ rows, uint32_t n_cols, uint64_t p_first, uint64_t n,
                                                      uint64_t binlen, unsigned long long *__restrict__ hist,
                                                      unsigned long long *__restrict__ col_sums) {
    typedef RwCfg<W> C;
    __shared__ unsigned int sh_all[2 * 129 + 128];
    unsigned int *sh_hist = sh_all, *sh_col = sh_all + 2 * 129;          // [2 bins][129 popcounts], [128 columns]
    for (uint32_t q = threadIdx.x; q < 2 * 129 + 128; q += 256) sh_all[q] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * C::BLOCK_ROWS;
    const uint64_t bin0 = binlen ? (p_first + base) / binlen : 0;
    const uint64_t split64 = binlen ? (bin0 + 1) * binlen - (p_first + base) : ~0ull;
    const uint32_t split = split64 > C::BLOCK_ROWS ? C::BLOCK_ROWS : (uint32_t)split64;       // rows [0, split) -> bin0, the rest -> bin0 + 1
    const uint32_t nrows = (uint32_t)min((uint64_t)C::BLOCK_ROWS, n - base);
    uint32_t cmask[C::WC];                                    // columns that exist, per word class
#pragma unroll
    for (int c = 0; c < C::WC; c++) {
        const uint32_t lo = 32u * c, bitsw = W == 2 ? 16u : 32u;
        const uint32_t m = n_cols >= lo + bitsw ? (bitsw == 32 ? 0xffffffffu : 0xffffu) : (n_cols > lo ? (1u << (n_cols - lo)) - 1u : 0u);
        cmask[c] = W == 2 ? m * 0x00010001u : m;
    }
    uint32_t acc4[C::WC][4], acc8[C::WC][4][2];
#pragma unroll
    for (int c = 0; c < C::WC; c++)
#pragma unroll
        for (int o = 0; o < 4; o++) { acc4[c][o] = 0; acc8[c][o][0] = 0; acc8[c][o][1] = 0; }
    uint32_t run_key = 0xffffffffu, run_cnt = 0;
    const uint4 *vec = (const uint4 *)(rows + base * W);
    for (int f = 0; f < C::NF; f++) {
#pragma unroll
        for (int t = 0; t < C::FL; t++) {
            const uint32_t v = (uint32_t)(f * C::FL + t) * 256u + threadIdx.x;          // vector of the block
            const uint32_t r0 = v * C::R;                                                // its first row
            if (r0 >= nrows) continue;
            uint32_t w[4];
            if (r0 + C::R <= nrows) {
                const uint4 x = vec[v];
                w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
            } else {                                                                      // the slice's last, partial vector
                const uint8_t *pb = rows + (base + r0) * W;
                const uint32_t nb = (nrows - r0) * W;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    w[q] = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) if ((uint32_t)(4 * q + b) < nb) w[q] |= (uint32_t)pb[4 * q + b] << (8 * b);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[q] &= cmask[q % C::WC];
#pragma unroll
                for (int o = 0; o < 4; o++) acc4[q % C::WC][o] += (w[q] >> o) & 0x11111111u;
            }
            // popcount per row -> run-length merged histogram updates
#pragma unroll
            for (int r = 0; r < (W == 2 ? 8 : C::R); r++) {
                uint32_t pc;
                if (W == 2) pc = __popc((w[r >> 1] >> (16 * (r & 1))) & 0xffffu);
                else if (W == 4) pc = __popc(w[r]);
                else if (W == 8) pc = __popc(w[2 * r]) + __popc(w[2 * r + 1]);
                else pc = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
                const uint32_t row = r0 + r;
                if (row < nrows) {
                    const uint32_t key = (row >= split ? 129u : 0u) + pc;
                    if (key == run_key) run_cnt++;
                    else {
                        if (run_cnt) atomicAdd(&sh_hist[run_key], run_cnt);
                        run_key = key; run_cnt = 1;
                    }
                }
            }
        }
        // widen: nibble fields -> byte fields
#pragma unroll
        for (int c = 0; c < C::WC; c++)
#pragma unroll
            for (int o = 0; o < 4; o++) {
                acc8[c][o][0] += acc4[c][o] & 0x0f0f0f0fu;
                acc8[c][o][1] += (acc4[c][o] >> 4) & 0x0f0f0f0fu;
                acc4[c][o] = 0;
            }
    }
    if (run_cnt) atomicAdd(&sh_hist[run_key], run_cnt);
    // byte b of acc8[c][o][h] counts bit 4 * (2 * b + h) + o of word class c
    const uint32_t lane = threadIdx.x & 31;
    if (col_sums) {
#pragma unroll
        for (int c = 0; c < C::WC; c++)
#pragma unroll
            for (int o = 0; o < 4; o++)
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint32_t bit = 4 * (2 * b + h) + o;
                        const uint32_t col = W == 2 ? (bit & 15u) : 32u * c + bit;
                        if (col < n_cols) {                                              // (uniform)
                            const uint32_t cnt = __reduce_add_sync(0xffffffffu, (acc8[c][o][h] >> (8 * b)) & 0xffu);
                            if (lane == 0 && cnt) atomicAdd(&sh_col[col], cnt);
                        }
                    }
    }
    __syncthreads();
    if (hist && binlen)
        for (uint32_t q = threadIdx.x; q < 2 * (n_cols + 1); q += 256) {
            const uint32_t which = q / (n_cols + 1), f = q % (n_cols + 1);
            const unsigned int v = sh_hist[which * 129 + f];
            if (v) atomicAdd(&hist[(bin0 + which) * (n_cols + 1) + f], (unsigned long long)v);
        }
    if (col_sums)
        for (uint32_t q = threadIdx.x; q < n_cols; q += 256)
            if (sh_col[q]) atomicAdd(&col_sums[q], (unsigned long long)sh_col[q]);
}
template <int W>
static void launch_reducew(const uint8_t *d_rows, uint32_t n_cols, uint64_t p_first, uint64_t n, uint64_t binlen,
                           unsigned long long *d_bin_hist, unsigned long long *d_col_sums, pk_stream_t s) {
    const uint32_t br = RwCfg<W>::BLOCK_ROWS;
    reducew_kernel<W><<<(unsigned)((n + br - 1) / br), 256, 0, s>>>(d_rows, n_cols, p_first, n, d_bin_hist ? binlen : 0, d_bin_hist, d_col_sums);
}

// low-res rows: rows_low[l - l0] = rows[l * step - p_first] for l in [l0, l1), l0 = ceil(p_first/step)
__global__ void __launch_bounds__(256) lowres_kernel(const uint8_t *__restrict__ rows, uint32_t row_stride, uint32_t nbytes,
                                                     uint64_t p_first, uint64_t l0, uint64_t n_low, uint32_t step,
                                                     uint8_t *__restrict__ rows_low) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n_low; q += stride) {
        const uint8_t *src = rows + ((l0 + q) * step - p_first) * row_stride;
        uint8_t *dst = rows_low + q * row_stride;
        for (uint32_t b = 0; b < nbytes; b++) dst[b] = src[b];
    }
}

void pk_launch_reduce(const uint8_t *d_rows, uint32_t row_stride, uint32_t n_cols, uint64_t p_first, uint64_t n,
                      uint64_t binlen, unsigned long long *d_bin_hist, unsigned long long *d_col_sums,
                      uint8_t *d_rows_low, uint32_t step, pk_stream_t s) {
    if (!n) return;
    const bool fast1 = n_cols <= 8 && row_stride == 1 && (((uintptr_t)d_rows) & 15) == 0 &&
                       (!d_bin_hist || binlen == 0 || binlen >= R1_BLOCK_ROWS);
    if ((d_bin_hist || d_col_sums) && fast1) {
        reduce1_kernel<<<(unsigned)((n + R1_BLOCK_ROWS - 1) / R1_BLOCK_ROWS), 256, 0, s>>>(d_rows, n_cols, p_first, n, d_bin_hist ? binlen : 0,
                                                                                           d_bin_hist, d_col_sums);
    } else if ((d_bin_hist || d_col_sums) && (row_stride == 2 || row_stride == 4 || row_stride == 8 || row_stride == 16) &&
               n_cols <= 8 * row_stride && n_cols <= 128 && (((uintptr_t)d_rows) & (row_stride - 1)) == 0 && n >= 64 &&
               (!d_bin_hist || binlen == 0 || binlen >= RW_MIN_BINLEN)) {
        // contiguous rows of 2, 4, 8 or 16 bytes: the (< 16 / W) rows before the first 16-byte boundary go through the
        // generic kernel, the rest through the vector kernel
        const uint64_t head = (((16 - (((uintptr_t)d_rows) & 15)) & 15)) / row_stride;
        if (head) {
            const uint32_t n_words = (n_cols + 31) / 32;
            const size_t shmem = (size_t)(n_cols + 1 + n_words * 32) * sizeof(unsigned int);
            if (p_first + n < (1ull << 32) && binlen < (1ull << 32))
                reduce_kernel<uint32_t><<<1, 256, shmem, s>>>(d_rows, row_stride, n_cols, p_first, head, binlen, d_bin_hist, d_col_sums);
            else
                reduce_kernel<uint64_t><<<1, 256, shmem, s>>>(d_rows, row_stride, n_cols, p_first, head, binlen, d_bin_hist, d_col_sums);
        }
        const uint8_t *r2 = d_rows + head * row_stride;
        if (row_stride == 2) launch_reducew<2>(r2, n_cols, p_first + head, n - head, binlen, d_bin_hist, d_col_sums, s);
        else if (row_stride == 4) launch_reducew<4>(r2, n_cols, p_first + head, n - head, binlen, d_bin_hist, d_col_sums, s);
        else if (row_stride == 8) launch_reducew<8>(r2, n_cols, p_first + head, n - head, binlen, d_bin_hist, d_col_sums, s);
        else launch_reducew<16>(r2, n_cols, p_first + head, n - head, binlen, d_bin_hist, d_col_sums, s);
    } else if (d_bin_hist || d_col_sums) {
        const uint32_t n_words = (n_cols + 31) / 32;
        const size_t shmem = (size_t)(n_cols + 1 + n_words * 32) * sizeof(unsigned int);
        const unsigned grid = (unsigned)((n + 256 * PK_RED_ITEMS - 1) / (256 * PK_RED_ITEMS));
        if (p_first + n < (1ull << 32) && binlen < (1ull << 32))
            reduce_kernel<uint32_t><<<grid, 256, shmem, s>>>(d_rows, row_stride, n_cols, p_first, n, binlen, d_bin_hist, d_col_sums);
        else
            reduce_kernel<uint64_t><<<grid, 256, shmem, s>>>(d_rows, row_stride, n_cols, p_first, n, binlen, d_bin_hist, d_col_sums);
    }
    if (d_rows_low) {
        const uint64_t l0 = (p_first + step - 1) / step, l1 = (p_first + n + step - 1) / step;
        if (l1 > l0)
            lowres_kernel<<<grid_for(l1 - l0), 256, 0, s>>>(d_rows, row_stride, (n_cols + 7) / 8, p_first, l0, l1 - l0, step, d_rows_low);
    }
}

// Pair-count bins (Index.bitmap_to_paircount_bins, panagram/index.py:454-459, the input of the UMAP CSVs): per bin of
// `rows_per_bin` consecutive low-res rows of ONE chromosome, the number of rows in which each genome's bit is set.
// One block per bin; counts[bin][g], uint32.
__global__ void __launch_bounds__(256) paircount_bins_kernel(const uint8_t *__restrict__ rows, uint32_t row_stride, uint32_t n_cols,
                                                             uint64_t n_rows, uint32_t rows_per_bin, uint32_t *__restrict__ counts) {
    extern __shared__ unsigned int sh_cnt[];            // [n_cols]
    for (uint32_t q = threadIdx.x; q < n_cols; q += blockDim.x) sh_cnt[q] = 0;
    __syncthreads();
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_bin;
    const uint64_t r1 = min(n_rows, r0 + rows_per_bin);
    const uint32_t nbytes = (n_cols + 7) / 8;
    for (uint64_t q = r0 * nbytes + threadIdx.x; q < r1 * nbytes; q += blockDim.x) {
        const uint64_t r = q / nbytes;
        const uint32_t b = (uint32_t)(q - r * nbytes);
        uint32_t v = rows[r * row_stride + b];
        while (v) {
            const uint32_t j = __ffs(v) - 1;
            v &= v - 1;
            if (8 * b + j < n_cols) atomicAdd(&sh_cnt[8 * b + j], 1u);
        }
    }
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < n_cols; q += blockDim.x) counts[(uint64_t)blockIdx.x * n_cols + q] = sh_cnt[q];
}
void pk_launch_paircount_bins(const uint8_t *d_rows, uint32_t row_stride, uint32_t n_cols, uint64_t n_rows, uint32_t rows_per_bin,
                              uint32_t *d_counts, pk_stream_t s) {
    if (!n_rows) return;
    const unsigned nbins = (unsigned)((n_rows + rows_per_bin - 1) / rows_per_bin);
    paircount_bins_kernel<<<nbins, 256, n_cols * sizeof(unsigned int), s>>>(d_rows, row_stride, n_cols, n_rows, rows_per_bin, d_counts);
}

// ------------------------------------------------------------------ interleave / unpack
__global__ void __launch_bounds__(256) interleave_kernel(const uint8_t *__restrict__ planes, uint32_t n_ranks, uint64_t n, uint32_t w,
                                                         uint8_t *__restrict__ rows, uint32_t row_stride) {
    const uint64_t total = n * n_ranks;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < total; q += stride) {
        const uint64_t i = q / n_ranks;
        const uint32_t r = (uint32_t)(q % n_ranks);
        const uint8_t *src = planes + ((uint64_t)r * n + i) * w;
        uint8_t *dst = rows + i * row_stride + (uint64_t)r * w;
        for (uint32_t b = 0; b < w; b++) dst[b] = src[b];
    }
}
void pk_launch_interleave(const uint8_t *d_planes, uint32_t n_ranks, uint64_t n, uint32_t w, uint8_t *d_rows,
                          uint32_t row_stride, pk_stream_t s) {
    if (!n || !n_ranks) return;
    interleave_kernel<<<grid_for(n * n_ranks), 256, 0, s>>>(d_planes, n_ranks, n, w, d_rows, row_stride);
}

// Fused exchange: the rows of R genome shards are assembled by reading every rank's plane IN PLACE over
// NVLink (peer-mapped pointers), transposing through shared memory and writing whole interleaved rows:
// rows[i][r*w .. r*w+w) = planes[r][i][0..w). One kernel replaces the NCCL all-gather, the [R][n][w]
// staging buffer and the interleave pass. Loads from a peer are 16-byte, coalesced; stores are 16-byte.
#define GI_MAX_RANKS 16
struct GatherArgs {
    const uint8_t *planes[GI_MAX_RANKS];
    uint32_t n_ranks, w, row_stride, tile_rows;
    uint64_t n;
    uint8_t *rows;
};
__global__ void __launch_bounds__(256) gather_interleave_kernel(const __grid_constant__ GatherArgs a) {
    extern __shared__ __align__(16) uint8_t g_tile[];          // [n_ranks][tile_rows * w]
    const uint32_t plane_bytes = a.tile_rows * a.w;
    for (uint64_t t0 = (uint64_t)blockIdx.x * a.tile_rows; t0 < a.n; t0 += (uint64_t)gridDim.x * a.tile_rows) {
        const uint32_t nrows = (uint32_t)min((uint64_t)a.tile_rows, a.n - t0);
        const uint32_t nbytes = nrows * a.w;
        for (uint32_t r = 0; r < a.n_ranks; r++) {
            const uint8_t *src = a.planes[r] + t0 * a.w;
            uint8_t *dst = g_tile + r * plane_bytes;
            if ((((uintptr_t)src) & 15) == 0) {
                for (uint32_t o = threadIdx.x * 16; o + 16 <= nbytes; o += 256 * 16)
                    *(uint4 *)(dst + o) = *(const uint4 *)(src + o);
                for (uint32_t o = (nbytes & ~15u) + threadIdx.x; o < nbytes; o += 256) dst[o] = src[o];
            } else {
                for (uint32_t o = threadIdx.x; o < nbytes; o += 256) dst[o] = src[o];
            }
        }
        __syncthreads();
        const uint32_t rw = a.n_ranks * a.w;
        uint8_t *out = a.rows + t0 * a.row_stride;
        if (a.w == 1 && (a.n_ranks & 3) == 0 && a.row_stride == rw && (((uintptr_t)out) & 15) == 0) {
            // 1 byte per rank and row (8 genomes per GPU): 4 rows per thread, one 32-bit shared load per rank,
            // byte transpose in registers, rows written as 32-bit words (consecutive lanes -> consecutive rows)
            for (uint32_t r4 = threadIdx.x * 4; r4 < nrows; r4 += 256 * 4) {
                uint32_t x[GI_MAX_RANKS];
#pragma unroll
                for (int r = 0; r < GI_MAX_RANKS; r++)
                    if ((uint32_t)r < a.n_ranks) x[r] = *(const uint32_t *)(g_tile + r * plane_bytes + r4);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (r4 + j < nrows) {
                        uint32_t *orow = (uint32_t *)(out + (uint64_t)(r4 + j) * rw);
#pragma unroll
                        for (int q = 0; q < GI_MAX_RANKS / 4; q++) {
                            if ((uint32_t)(4 * q) < a.n_ranks) {
                                const uint32_t lo = __byte_perm(x[4 * q], x[4 * q + 1], 0x0040 + j * 0x0011);
                                const uint32_t hi = __byte_perm(x[4 * q + 2], x[4 * q + 3], 0x0040 + j * 0x0011);
                                orow[q] = (lo & 0xffffu) | (hi << 16);
                            }
                        }
                    }
                }
            }
        } else if (a.w == 1 && a.n_ranks == 2 && a.row_stride == 2 && (((uintptr_t)out) & 15) == 0) {
            // two ranks of 8 genomes: 4 rows per thread, byte-interleaved in registers, one 8-byte store
            for (uint32_t r4 = threadIdx.x * 4; r4 < nrows; r4 += 256 * 4) {
                const uint32_t x0 = *(const uint32_t *)(g_tile + r4), x1 = *(const uint32_t *)(g_tile + plane_bytes + r4);
                const uint32_t lo = __byte_perm(x0, x1, 0x5140), hi = __byte_perm(x0, x1, 0x7362);
                if (r4 + 4 <= nrows) {
                    *(uint2 *)(out + (uint64_t)r4 * 2) = make_uint2(lo, hi);
                } else {
                    const uint32_t v[2] = {lo, hi};
                    for (uint32_t j = 0; r4 + j < nrows; j++) {
                        out[(uint64_t)(r4 + j) * 2] = (uint8_t)(v[j >> 1] >> (16 * (j & 1)));
                        out[(uint64_t)(r4 + j) * 2 + 1] = (uint8_t)(v[j >> 1] >> (16 * (j & 1) + 8));
                    }
                }
            }
        } else if (a.row_stride == rw && (((uintptr_t)out) & 15) == 0) {
            const uint32_t total = nrows * rw;
            for (uint32_t o = threadIdx.x * 16; o < total; o += 256 * 16) {
                uint8_t tmp[16];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const uint32_t oo = o + q;
                    const uint32_t row = oo / rw, within = oo - row * rw, r = within / a.w, b = within - r * a.w;
                    tmp[q] = oo < total ? g_tile[r * plane_bytes + row * a.w + b] : 0;
                }
                if (o + 16 <= total) *(uint4 *)(out + o) = *(const uint4 *)tmp;
                else for (uint32_t q = 0; o + q < total; q++) out[o + q] = tmp[q];
            }
        } else {
            for (uint32_t e = threadIdx.x; e < nrows * rw; e += 256) {
                const uint32_t row = e / rw, within = e - row * rw, r = within / a.w, b = within - r * a.w;
                out[(uint64_t)row * a.row_stride + within] = g_tile[r * plane_bytes + row * a.w + b];
            }
        }
        __syncthreads();
    }
}
int pk_launch_gather_interleave(const void *const *planes, uint32_t n_ranks, uint64_t n, uint32_t w, uint8_t *d_rows,
                                uint32_t row_stride, pk_stream_t s) {
    if (!n) return 0;
    if (n_ranks > GI_MAX_RANKS) return -1;
    GatherArgs a{};
    for (uint32_t r = 0; r < n_ranks; r++) a.planes[r] = (const uint8_t *)planes[r];
    a.n_ranks = n_ranks; a.w = w; a.row_stride = row_stride; a.n = n; a.rows = d_rows;
    uint32_t tile = 32768 / (n_ranks * w);            // 32 KB of shared memory per block
    tile = tile / 16 * 16;
    if (tile < 16) tile = 16;
    a.tile_rows = tile;
    const size_t shmem = (size_t)n_ranks * tile * w;
    const uint64_t ntiles = (n + tile - 1) / tile;
    const unsigned grid = (unsigned)(ntiles < 148ull * 6 ? ntiles : 148ull * 6);
    gather_interleave_kernel<<<grid, 256, shmem, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Position-split exchange (pk_gather.cuh): this rank's slice of the rows out of all ranks' planes, read in place
// over NVLink. One thread per 16-byte chunk of plane bytes, grid-stride; no shared memory.
template <int RMAX, int W>
__global__ void __launch_bounds__(256) gather_slice_kernel(const __grid_constant__ PkgArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < a.n_chunks; t += stride) pkg_gather_chunk<RMAX, W>(a, t);
}
template <int RMAX>
static void gather_slice_launch(const PkgArgs &a, unsigned grid, pk_stream_t s) {
    switch (a.w) {
        case 1: gather_slice_kernel<RMAX, 1><<<grid, 256, 0, s>>>(a); break;
        case 2: gather_slice_kernel<RMAX, 2><<<grid, 256, 0, s>>>(a); break;
        case 4: gather_slice_kernel<RMAX, 4><<<grid, 256, 0, s>>>(a); break;
        case 8: gather_slice_kernel<RMAX, 8><<<grid, 256, 0, s>>>(a); break;
        case 16: gather_slice_kernel<RMAX, 16><<<grid, 256, 0, s>>>(a); break;
        default: gather_slice_kernel<RMAX, 0><<<grid, 256, 0, s>>>(a); break;
    }
}
template <int R, int W>
__global__ void __launch_bounds__(256) gather_slice_dst_kernel(const __grid_constant__ PkgArgs a, const uint64_t total_rows) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < a.n_chunks; t += stride) pkg_gather_dst_chunk<R, W>(a, t, total_rows);
}
// narrow rows, segments back to back in the output: one 16-byte store per thread (pk_gather.cuh)
int pk_launch_gather_slice_dst(const void *const *planes, uint32_t n_ranks, uint64_t plane_rows, uint32_t w, const void *d_segs,
                               uint32_t n_segs, uint64_t total_rows, uint8_t *d_rows, pk_stream_t s) {
    if (!total_rows) return 0;
    PkgArgs a{};
    for (uint32_t r = 0; r < n_ranks; r++) a.planes[r] = (const uint8_t *)planes[r];
    a.n_ranks = n_ranks; a.w = w; a.plane_rows = plane_rows; a.segs = (const PkgSeg *)d_segs; a.n_segs = n_segs;
    a.row_stride = n_ranks * w; a.row_bytes = n_ranks * w; a.rows = d_rows;
    const uint32_t nr = 16 / (n_ranks * w);
    a.n_chunks = (total_rows + nr - 1) / nr;
    const uint64_t nb = (a.n_chunks + 255) / 256;
    const unsigned grid = (unsigned)(nb < 148ull * 16 ? nb : 148ull * 16);
    const uint32_t key = n_ranks * 100 + w;
    switch (key) {
        case 201: gather_slice_dst_kernel<2, 1><<<grid, 256, 0, s>>>(a, total_rows); break;
        case 202: gather_slice_dst_kernel<2, 2><<<grid, 256, 0, s>>>(a, total_rows); break;
        case 204: gather_slice_dst_kernel<2, 4><<<grid, 256, 0, s>>>(a, total_rows); break;
        case 401: gather_slice_dst_kernel<4, 1><<<grid, 256, 0, s>>>(a, total_rows); break;
        case 402: gather_slice_dst_kernel<4, 2><<<grid, 256, 0, s>>>(a, total_rows); break;
        case 801: gather_slice_dst_kernel<8, 1><<<grid, 256, 0, s>>>(a, total_rows); break;
        default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int pk_launch_gather_slice(const void *const *planes, uint32_t n_ranks, uint64_t plane_rows, uint32_t w, const void *d_segs,
                           uint32_t n_segs, uint64_t n_chunks, uint8_t *d_rows, uint32_t row_stride, uint32_t row_bytes, pk_stream_t s) {
    if (!n_chunks) return 0;
    if (n_ranks > PKG_MAX_RANKS) return -1;
    PkgArgs a{};
    for (uint32_t r = 0; r < n_ranks; r++) a.planes[r] = (const uint8_t *)planes[r];
    a.n_ranks = n_ranks; a.w = w; a.plane_rows = plane_rows; a.segs = (const PkgSeg *)d_segs; a.n_segs = n_segs;
    a.row_stride = row_stride; a.row_bytes = row_bytes; a.rows = d_rows; a.n_chunks = n_chunks;
    const uint64_t nb = (n_chunks + 255) / 256;
    const unsigned grid = (unsigned)(nb < 148ull * 8 ? nb : 148ull * 8);
    if (n_ranks <= 2) gather_slice_launch<2>(a, grid, s);
    else if (n_ranks <= 4) gather_slice_launch<4>(a, grid, s);
    else if (n_ranks <= 8) gather_slice_launch<8>(a, grid, s);
    else gather_slice_launch<16>(a, grid, s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

__global__ void __launch_bounds__(256) rows_to_u32_kernel(const uint8_t *__restrict__ rows, uint32_t row_stride, uint32_t byte_off,
                                                          uint32_t n_bytes, uint32_t bit_mask, uint64_t n, uint32_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t v = 0;
        for (uint32_t b = 0; b < n_bytes; b++) v |= (uint32_t)rows[i * row_stride + byte_off + b] << (8 * b);
        out[i] = v & bit_mask;
    }
}
void pk_launch_rows_to_u32(const uint8_t *d_rows, uint32_t row_stride, uint32_t byte_off, uint32_t n_bytes,
                           uint32_t bit_mask, uint64_t n, uint32_t *d_out, pk_stream_t s) {
    if (!n) return;
    rows_to_u32_kernel<<<grid_for(n), 256, 0, s>>>(d_rows, row_stride, byte_off, n_bytes, bit_mask, n, d_out);
}
