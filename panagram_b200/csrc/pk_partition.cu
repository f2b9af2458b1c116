// Partitioned probe ("sort-probe") — the large-batch hot path.
//
// A membership probe reads one random 32 B bucket. Random sector reads over a multi-GB table
// run at ~40 G/s on B200 (~20% of HBM peak, measured: profiles/sector_gups_r1.md); the same
// reads confined to a window that stays L2-resident run at 150-300 G/s. All genomes' tables
// are indexed by the same hash, so ONE radix partition of the anchor's (hash, position) pairs
// by the top PB hash bits gives, for every table, a contiguous bucket window per partition:
//
//   K1 partition_seq   packed sequence -> (hash32, pos) pairs, split by the top PB1 hash bits
//   K2 partition_fine  each coarse region split again by the next PB2 bits      (PB = PB1+PB2)
//   K3 probe_part      one block per fine partition: canonical k-mers re-derived from the
//                      (L2-resident) packed sequence into shared memory, then genome by genome
//                      every probe of the block lands in that table's 16-64 KB window; the next
//                      genome's window is pulled into L2 by a TMA bulk prefetch while the
//                      current one is probed. Row bits are scattered to rows[pos].
//
// Partitions have fixed capacity (hashing is uniform); items that do not fit (skew from
// repeats) go to a spill list that the same probe kernel drains with unconfined probes.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "pk_device.cuh"
#include "pk_internal.h"

#define PT_THREADS 256
#define PT_IPT 16
#define PT_TILE (PT_THREADS * PT_IPT)
#define PT_WARPS (PT_THREADS / 32)
#define PT_MAXB 512

struct PartSmem {
    uint32_t tot[PT_MAXB];                  // per-digit tile totals (the rank counters)
    uint32_t dstart[PT_MAXB];               // exclusive scan of tot: start of the digit's run in `stage`
    uint32_t gbase[PT_MAXB];                // reserved offset inside the destination region
    uint32_t warp_sums[PT_WARPS];
    uint32_t tile_total;
    uint2 stage[PT_TILE];
    uint16_t dig[PT_TILE];                  // compact K1 items do not carry their digit any more: kept beside them
};

// COMPACT ITEMS (one-byte rows out of 32-bit-slot group tables, k = 19..22, launches of >= 2^9 coarse regions and
// < 2^28 positions). K3 spent most of its L1TEX time re-deriving every k-mer from the packed sequence — two random
// 8-byte loads per item, 270 M sector requests on configs[1] — only to get the 20 key-remainder bits its 32-bit-slot
// probe compares. A G32 probe needs exactly: the hash (bucket index) and those 20 bits; and of the hash, the low
// 32 - eb bits are a function of the remainder (pk_g32_hash), the top 9 bits are the coarse region the item sits in.
// So K1 emits 8-byte items that carry everything K3 needs:
//     x = [ M : eb - 9 hash bits below the coarse digit ][ R20 >> 4 : 16 bits ]      y = [ R20 & 15 : 4 bits ][ pos : 28 bits ]
// K2 takes its digit from the top pb2 bits of M (a plain shift of x), K3 rebuilds h = coarse << 23 | M << (32 - eb) |
// mix32(R20) >> eb and never touches the sequence (only the rare walk-on fallback does).
#define PT_CI_POS_BITS 28
#define PT_CI_POS_MASK 0x0FFFFFFFu


// Scatter a tile of items (h, pos) into fixed-capacity regions by digit = (h >> shift) & (nb-1).
// Region r = region0 + digit holds items dst[r * cap .. r * cap + min(cursor[r], cap)).
// Rank inside (tile, digit): one shared-memory atomicAdd per item (measured 2x faster than warp
// match_any ranking with per-warp counters: K1 1.32 -> 0.98 ms, K2 1.20 -> 0.64 ms). POS_STRIDE > 0:
// the item positions are pos0 + j * POS_STRIDE and are not kept in registers.
template <int POS_STRIDE>
__device__ __forceinline__ void block_partition(PartSmem &S, const uint32_t (&h)[PT_IPT], const uint32_t (&pos)[POS_STRIDE ? 1 : PT_IPT],
                                                uint32_t pos0, uint32_t validmask, uint32_t shift, uint32_t nb,
                                                uint2 *__restrict__ dst, uint64_t region0, uint32_t cap,
                                                uint32_t *__restrict__ cursor, uint2 *__restrict__ spill,
                                                unsigned long long *__restrict__ spill_cursor, uint64_t spill_cap,
                                                uint32_t *__restrict__ err) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (uint32_t d = tid; d < nb; d += PT_THREADS) S.tot[d] = 0;
    __syncthreads();
    uint16_t rank[PT_IPT];
#pragma unroll
    for (int j = 0; j < PT_IPT; j++)
        rank[j] = (validmask >> j) & 1 ? (uint16_t)atomicAdd(&S.tot[(h[j] >> shift) & (nb - 1)], 1u) : 0;
    __syncthreads();
    {   // exclusive scan of tot[0..nb) with 2 digits per thread (nb <= 512 = 2 * PT_THREADS)
        const uint32_t a = 2 * tid < nb ? S.tot[2 * tid] : 0, b = 2 * tid + 1 < nb ? S.tot[2 * tid + 1] : 0;
        uint32_t s = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += y;
        }
        if (lane == 31) S.warp_sums[w] = s;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int ww = 0; ww < PT_WARPS; ww++) woff += ww < (int)w ? S.warp_sums[ww] : 0;
        const uint32_t excl = woff + s - (a + b);
        if (2 * tid < nb) S.dstart[2 * tid] = excl;
        if (2 * tid + 1 < nb) S.dstart[2 * tid + 1] = excl + a;
        if (tid == PT_THREADS - 1) S.tile_total = woff + s;
    }
    // reserve the runs (one global atomic per digit of the tile); the round trip overlaps with the staging
    uint32_t gb[PT_MAXB / PT_THREADS];
#pragma unroll
    for (int q = 0; q < PT_MAXB / PT_THREADS; q++) {
        const uint32_t d = tid + q * PT_THREADS;
        gb[q] = (d < nb && S.tot[d]) ? atomicAdd(&cursor[region0 + d], S.tot[d]) : 0;
    }
    __syncthreads();            // dstart / tile_total visible
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        if ((validmask >> j) & 1) {
            const uint32_t d = (h[j] >> shift) & (nb - 1);
            S.stage[S.dstart[d] + rank[j]] = make_uint2(h[j], POS_STRIDE ? pos0 + j * POS_STRIDE : pos[POS_STRIDE ? 0 : j]);
        }
    }
#pragma unroll
    for (int q = 0; q < PT_MAXB / PT_THREADS; q++) {
        const uint32_t d = tid + q * PT_THREADS;
        if (d < nb) S.gbase[d] = gb[q];
    }
    __syncthreads();
    const uint32_t total = S.tile_total;
    for (uint32_t i = tid; i < total; i += PT_THREADS) {
        const uint2 it = S.stage[i];
        const uint32_t d = (it.x >> shift) & (nb - 1);
        const uint64_t off = (uint64_t)S.gbase[d] + (i - S.dstart[d]);
        if (off < cap) {
            dst[(region0 + d) * cap + off] = it;
        } else {
            const unsigned long long s = atomicAdd(spill_cursor, 1ull);
            if (s < spill_cap) spill[s] = it; else *err = 1;
        }
    }
    __syncthreads();
}

// The K1 form for compact items: per item ONE register holds the finished x word and one more the 9-bit coarse digit,
// the 4 remainder bits of the y word and, once known, the rank inside the (tile, digit) run — 32 registers for the
// 16 items of a thread instead of hash + remainder + rank (40; the rolling K1 spilled with those). Positions are
// pos0 + j; 512 coarse regions.
__device__ __forceinline__ void block_partition_ck1(PartSmem &S, const uint32_t (&xp)[PT_IPT], uint32_t (&aux)[PT_IPT], uint32_t pos0,
                                                    uint32_t validmask, uint2 *__restrict__ dst, uint32_t cap, uint32_t *__restrict__ cursor,
                                                    uint2 *__restrict__ spill, unsigned long long *__restrict__ spill_cursor,
                                                    uint64_t spill_cap, uint32_t *__restrict__ err) {
    constexpr uint32_t nb = 512;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (uint32_t d = tid; d < nb; d += PT_THREADS) S.tot[d] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PT_IPT; j++)
        if ((validmask >> j) & 1) aux[j] |= atomicAdd(&S.tot[aux[j] & 511u], 1u) << 13;
    __syncthreads();
    {
        const uint32_t a = S.tot[2 * tid], b = S.tot[2 * tid + 1];
        uint32_t s = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += y;
        }
        if (lane == 31) S.warp_sums[w] = s;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int ww = 0; ww < PT_WARPS; ww++) woff += ww < (int)w ? S.warp_sums[ww] : 0;
        const uint32_t excl = woff + s - (a + b);
        S.dstart[2 * tid] = excl;
        S.dstart[2 * tid + 1] = excl + a;
        if (tid == PT_THREADS - 1) S.tile_total = woff + s;
    }
    uint32_t gb[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const uint32_t d = tid + q * PT_THREADS;
        gb[q] = S.tot[d] ? atomicAdd(&cursor[d], S.tot[d]) : 0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        if ((validmask >> j) & 1) {
            const uint32_t d = aux[j] & 511u, slot = S.dstart[d] + (aux[j] >> 13);
            S.stage[slot] = make_uint2(xp[j], (((aux[j] >> 9) & 15u) << PT_CI_POS_BITS) | (pos0 + j));
            S.dig[slot] = (uint16_t)d;
        }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) S.gbase[tid + q * PT_THREADS] = gb[q];
    __syncthreads();
    const uint32_t total = S.tile_total;
    for (uint32_t i = tid; i < total; i += PT_THREADS) {
        const uint2 it = S.stage[i];
        const uint32_t d = S.dig[i];
        const uint64_t off = (uint64_t)S.gbase[d] + (i - S.dstart[d]);
        if (off < cap) {
            dst[(uint64_t)d * cap + off] = it;
        } else {
            const unsigned long long s = atomicAdd(spill_cursor, 1ull);
            if (s < spill_cap) spill[s] = it; else *err = 1;
        }
    }
    __syncthreads();
}

struct PartArgs {
    const uint64_t *words;
    const uint64_t *mask64;
    uint64_t p0, off, n;            // batch base position, first position of this launch (relative), count
    PkKeySpec ks;
    uint32_t pb1, pb2;              // coarse / fine bits
    uint32_t cap1, cap2;            // region capacities
    uint2 *buf1, *buf2, *spill;
    uint32_t *cursor1, *cursor2;
    unsigned long long *spill_cursor;
    uint64_t spill_cap;
    uint32_t *err;
    uint8_t *rows;
    uint32_t row_stride, col_offset, nbl;
    uint32_t compact, eb;           // compact items (see PartSmem): on/off, hash bits that carry key bits (2k - 20)
};

// K1: positions [p0 + off, p0 + off + n) -> (hash, i) pairs (i = position - p0), partitioned by the top pb1 bits.
// Invalid windows get their all-zero row here and never enter the pipeline.
__global__ void __launch_bounds__(PT_THREADS, 4) partition_seq_kernel(PartArgs a) {
    __shared__ PartSmem S;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t base = a.off + blockIdx.x * (uint64_t)PT_TILE + (uint64_t)w * (32 * PT_IPT) + lane;
    uint32_t h[PT_IPT], valid = 0;
    const uint32_t nopos[1] = {0};
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        const uint64_t i = base + 32 * j;
        h[j] = 0;
        if (i < a.off + a.n) {
            uint64_t canon;
            if (pk_window(a.words, a.mask64, a.p0 + i, a.ks.k, canon)) {
                h[j] = pk_probe_hash(canon, a.ks);
                valid |= 1u << j;
            } else {
                uint8_t *dst = a.rows + i * a.row_stride + a.col_offset;
                for (uint32_t q = 0; q < a.nbl; q++) dst[q] = 0;
            }
        }
    }
    block_partition<32>(S, h, nopos, (uint32_t)base, valid, 32 - a.pb1, 1u << a.pb1, a.buf1, 0, a.cap1, a.cursor1, a.spill,
                        a.spill_cursor, a.spill_cap, a.err);
}

// K1, rolling form: a thread takes PT_IPT CONSECUTIVE positions, extracts the first k-mer and its reverse complement
// once (three packed words, two mask words) and then slides: one base in at the low end of the forward k-mer, its
// complement in at the high end of the reverse one — the SHL_insert2bits / SHR_insert2bits pair of
// kmer_api.h:54-81 — instead of re-extracting and re-reversing every window. Same items, same order of positions
// inside a thread's run (the rank inside a (tile, digit) run differs, which no consumer depends on).
template <int MINB, int COMPACT>
__global__ void __launch_bounds__(PT_THREADS, MINB) partition_seq_roll_kernel(PartArgs a) {
    __shared__ PartSmem S;
    const uint64_t base = a.off + blockIdx.x * (uint64_t)PT_TILE + (uint64_t)threadIdx.x * PT_IPT;
    const uint64_t end = a.off + a.n;
    uint32_t h[PT_IPT], aux[COMPACT ? PT_IPT : 1], valid = 0;        // compact: h holds the item's x word, aux digit | y bits
    const uint32_t nopos[1] = {0};
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) { h[j] = 0; if (COMPACT) aux[COMPACT ? j : 0] = 0; }
    if (base < end) {
        const uint32_t k = a.ks.k;
        const uint64_t P = a.p0 + base;
        const uint64_t wi = P >> 5;
        const uint32_t s = 2 * ((uint32_t)P & 31);
        const uint64_t w0 = a.words[wi], w1 = a.words[wi + 1], w2 = a.words[wi + 2];
        const uint64_t hi = s ? (w0 << s) | (w1 >> (64 - s)) : w0;          // bases P .. P+31
        const uint64_t lo = s ? (w1 << s) | (w2 >> (64 - s)) : w1;          // bases P+32 .. P+63
        uint64_t fwd = hi >> (64 - 2 * k);
        uint64_t rc = pk_revcomp(fwd, k);
        uint64_t up = k < 32 ? (hi << (2 * k)) | (lo >> (64 - 2 * k)) : lo;   // the bases that slide in, first one on top
        const uint64_t mi = P >> 6;
        const uint32_t t = (uint32_t)P & 63;
        const uint64_t m0 = a.mask64[mi], m1 = a.mask64[mi + 1];
        const uint64_t M = t ? (m0 >> t) | (m1 << (64 - t)) : m0;           // bit i: base P+i is not ACGT
        const uint64_t wmask = (1ull << k) - 1;                             // k <= 32
        const uint64_t kk = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
        const uint32_t topsh = 2 * (k - 1);
        const uint32_t mmask = COMPACT ? (1u << (a.eb - 9)) - 1 : 0, msh = COMPACT ? 32 - a.eb : 0;
#pragma unroll
        for (int j = 0; j < PT_IPT; j++) {
            if (base + j < end) {
                if (((M >> j) & wmask) == 0) {
                    const uint64_t canon = fwd < rc ? fwd : rc;
                    const uint32_t hh = pk_probe_hash(canon, a.ks);
                    if (COMPACT) {
                        const uint32_t r20 = (uint32_t)canon & PK_G32_REM_MASK;
                        h[j] = (((hh >> msh) & mmask) << 16) | (r20 >> 4);
                        aux[COMPACT ? j : 0] = (hh >> 23) | ((r20 & 15u) << 9);
                    } else {
                        h[j] = hh;
                    }
                    valid |= 1u << j;
                } else {
                    uint8_t *dst = a.rows + (base + j) * a.row_stride + a.col_offset;
                    for (uint32_t q = 0; q < a.nbl; q++) dst[q] = 0;
                }
            }
            const uint64_t b = up >> 62;
            up <<= 2;
            fwd = ((fwd << 2) | b) & kk;
            rc = (rc >> 2) | ((3 - b) << topsh);
        }
    }
    if constexpr (COMPACT)
        block_partition_ck1(S, h, aux, (uint32_t)base, valid, a.buf1, a.cap1, a.cursor1, a.spill, a.spill_cursor, a.spill_cap, a.err);
    else
        block_partition<1>(S, h, nopos, (uint32_t)base, valid, 32 - a.pb1, 1u << a.pb1, a.buf1, 0, a.cap1, a.cursor1, a.spill,
                           a.spill_cursor, a.spill_cap, a.err);
}

// K2: coarse region c = blockIdx.y, tile blockIdx.x of it -> fine regions c * 2^pb2 + next pb2 bits.
__global__ void __launch_bounds__(PT_THREADS, 4) partition_fine_kernel(PartArgs a) {
    __shared__ PartSmem S;
    const uint32_t c = blockIdx.y;
    const uint32_t cnt = min(a.cursor1[c], a.cap1);
    const uint64_t tile0 = blockIdx.x * (uint64_t)PT_TILE;
    if (tile0 >= cnt) return;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint2 *src = a.buf1 + (uint64_t)c * a.cap1;
    uint32_t h[PT_IPT], pos[PT_IPT], valid = 0;
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        const uint64_t i = tile0 + (uint64_t)w * (32 * PT_IPT) + 32 * j + lane;
        h[j] = 0; pos[j] = 0;
        if (i < cnt) {
            const uint2 it = src[i];
            h[j] = it.x; pos[j] = it.y;
            valid |= 1u << j;
        }
    }
    // compact items: the fine digit is the top pb2 bits of M, which sits above the 16 remainder bits of x
    const uint32_t shift = a.compact ? 16 + (a.eb - 9) - a.pb2 : 32 - a.pb1 - a.pb2;
    block_partition<0>(S, h, pos, 0, valid, shift, 1u << a.pb2, a.buf2, (uint64_t)c << a.pb2, a.cap2, a.cursor2,
                       a.spill, a.spill_cursor, a.spill_cap, a.err);
}

// ------------------------------------------------------------------ K3: probe one partition per block
#define PP_OBINS 128                      // position bins of the un-permute lists
#define PP_OCS 32                         // the bins' cursors sit 128 bytes apart: 2^18 blocks add to each of them
#define PP_FBINS 512                      // fine mode (one-byte rows): up to 512 bins ...
#define PP_FSLICE_SHIFT 17                // ... un-permuted through 128 KB shared-memory slices of the bitmap
#define PS_BINS 1024                      // counting-sort bins of the bucket-sorted variant

__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One launch covers one group of <= 32 genomes; their table descriptors travel as kernel parameters
// (constant bank: no LDS/LDG per probe).
struct ProbeArgs {
    const uint2 *buf;                 // regions of `cap` (hash, pos) items
    const uint32_t *counts;           // per-region fill (may exceed cap: clamp); NULL => flat list of *flat_total items
    const unsigned long long *flat_total;
    uint32_t cap, n_regions;
    uint32_t pb;                      // total partition bits (regions mode; 0 disables prefetch)
    const uint64_t *words;
    uint64_t p0;
    PkKeySpec ks;
    uint32_t ng, grp;                 // tables in this launch, column group (bytes 4*grp.. of a row)
    uint32_t tbits;                   // genomes per table: 1 (per-genome tables) or 8 (group tables: tabs[u] covers bits 8u..8u+7)
    uint32_t g_first;                 // local index of the launch's first genome (stash keys)
    uint32_t n_genomes;               // genomes covered by the launch (<= 32)
    // window kernel: the loop runs over `ng` WINDOW PIECES; piece g is chunk c_of[g] (of chunk_buckets buckets; 0 = the
    // whole window) of table tabs[t_of[g]]'s window — a window larger than a shared-memory stage is probed piece by piece
    uint32_t chunk_buckets;
    uint8_t t_of[32], c_of[32];
    uint8_t *rows;
    uint32_t row_stride, col_offset, nbl;
    int prefetch;
    // un-permute: instead of scattering row bytes over the whole bitmap (32x write amplification and a
    // read-modify-write per byte in DRAM: measured 4.3 GB extra reads + 4.6 GB writes on configs[1]),
    // results are appended as (pos, bits) to lists binned by pos >> out_shift; unpermute_kernel then
    // scatters each list inside its own L2-resident slice of the bitmap
    uint2 *out_list;                  // [(grp * PP_OBINS + bin) << out_shift]
    uint32_t *out_cursor;             // [n_groups * PP_OBINS]
    uint32_t out_shift;
    uint32_t compact, eb, pb2;        // compact items (PartSmem): on/off, 2k - 20, fine partition bits
    uint32_t n_bins;                  // position bins of this launch (<= PP_OBINS, fine mode <= PP_FBINS)
    uint32_t out_fine;                // one-byte rows: 4-byte items ((position in bin) << 8 | bits), bins indexed without the group factor
    uint32_t rank_atomic;             // window kernel: rank the results inside their position bin by one shared-memory atomicAdd per
                                      // item (1) or by warp match_any + per-warp counters (0)
    PkTable tabs[32];
};

// K3. One block per fine partition (<= T*IPT items). Items go to registers (k-mer re-derived from the
// L2-resident packed sequence); genome by genome every probe of the block lands in that partition's
// window of the table (8 KB at 2^18 partitions of a 2 GB table) while a TMA bulk prefetch pulls the
// window of the genome after next into L2. One load in flight per thread and many warps beat ILP here
// (measured: 8/4/2/1 loads in flight -> 15.9/9.7/8.4/6.9 ms). A full home bucket without the key (the
// key may sit in a later bucket) is rare; those walk-ons are queued in shared memory and resolved by all
// lanes together instead of diverging inside the probe loop (-40 % issued instructions).
// SORT = 1 first counting-sorts the items by the hash bits below the partition bits, so that the lanes of
// a warp probe neighbouring buckets (half the L1 tag lookups; pays off once the tables are small enough
// for the kernel to be L1-bound rather than DRAM-bound).
template <int T, int IPT, int FMT, int SORT, int MINB, int GU>
__global__ void __launch_bounds__(T, MINB) probe_part_kernel(const __grid_constant__ ProbeArgs a) {
    constexpr int CAP = T * IPT;
    __shared__ uint32_t s_bits[CAP];               // results of the deferred walk-ons, by item
    __shared__ unsigned long long q_key[CAP];      // deferred walk-on queue: canonical k-mer,
    __shared__ uint32_t q_meta[CAP], q_h[CAP];     //   (item << 5 | genome), hash
    __shared__ uint32_t q_n, s_ws[T / 32];
    __shared__ uint16_t o_wc[T / 32][PP_OBINS];
    __shared__ uint32_t o_gb[PP_OBINS];
    __shared__ unsigned long long s_canon[SORT ? CAP : 1];
    __shared__ uint32_t s_h[SORT ? CAP : 1], s_pos[SORT ? CAP : 1], s_cnt[SORT ? PS_BINS : 1];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint64_t nreg = a.n_regions;
    unsigned long long flat = 0;
    if (!a.counts) { flat = *a.flat_total; nreg = (flat + a.cap - 1) / a.cap; }
    const uint32_t sshift = a.pb + 10 <= 32 ? 32 - a.pb - 10 : 0;
    for (uint64_t q = blockIdx.x; q < nreg; q += gridDim.x) {
        uint32_t cnt;
        if (a.counts) cnt = min(a.counts[q], a.cap);
        else cnt = (uint32_t)min((unsigned long long)a.cap, flat - q * a.cap);
        if (cnt == 0) continue;
        const bool do_pf = a.prefetch && a.counts && a.pb;
        const uint32_t h_lo = do_pf ? (uint32_t)(q << (32 - a.pb)) : 0;
        const uint32_t h_hi = do_pf ? (uint32_t)(((q + 1) << (32 - a.pb)) - 1) : 0;
        if (do_pf && tid < min(2u, a.ng)) {
            const PkTable t = a.tabs[tid];
            const uint32_t b0 = __umulhi(h_lo, t.n_buckets), b1 = __umulhi(h_hi, t.n_buckets);
            if (b1 - b0 < 8192) l2_prefetch_bulk(t.slots + 4ull * b0, (b1 - b0 + 1) * 32);
        }
        if (SORT) for (uint32_t i = tid; i < PS_BINS; i += T) s_cnt[i] = 0;
        for (uint32_t i = tid; i < (uint32_t)CAP; i += T) s_bits[i] = 0;
        for (uint32_t i = tid; i < (T / 32) * PP_OBINS; i += T) (&o_wc[0][0])[i] = 0;
        if (tid == 0) q_n = 0;
        __syncthreads();
        const uint2 *src = a.buf + q * (uint64_t)a.cap;
        uint64_t canon[IPT];
        uint32_t h[IPT], pos[IPT], rnk[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const uint32_t i = tid + j * T;
            canon[j] = 0; h[j] = 0; pos[j] = 0; rnk[j] = 0;
            if (i < cnt) {
                const uint2 it = src[i];
                canon[j] = pk_canon_at(a.words, a.p0 + it.y, a.ks.k);
                h[j] = it.x; pos[j] = it.y;
                if (SORT) rnk[j] = atomicAdd(&s_cnt[(it.x >> sshift) & (PS_BINS - 1)], 1u);
            }
        }
        if (SORT) {
            __syncthreads();
            {   // exclusive scan of s_cnt[PS_BINS], PS_BINS / T bins per thread
                constexpr int BPT = PS_BINS / T;
                uint32_t c[BPT], sum = 0;
#pragma unroll
                for (int b = 0; b < BPT; b++) { c[b] = s_cnt[tid * BPT + b]; sum += c[b]; }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += y;
                }
                if (lane == 31) s_ws[wid] = inc;
                __syncthreads();
                uint32_t off = inc - sum;
#pragma unroll
                for (int ww = 0; ww < T / 32; ww++) off += ww < (int)wid ? s_ws[ww] : 0;
#pragma unroll
                for (int b = 0; b < BPT; b++) { s_cnt[tid * BPT + b] = off; off += c[b]; }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                if (tid + j * T < cnt) {
                    const uint32_t d = s_cnt[(h[j] >> sshift) & (PS_BINS - 1)] + rnk[j];
                    s_canon[d] = canon[j]; s_h[d] = h[j]; s_pos[d] = pos[j];
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                const uint32_t i = tid + j * T;
                if (i < cnt) { canon[j] = s_canon[i]; h[j] = s_h[i]; pos[j] = s_pos[i]; }
            }
        }
        uint32_t bits[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) bits[j] = 0;
        // ---- probe, genome by genome (GU genomes' loads in flight per item)
        for (uint32_t g0 = 0; g0 < a.ng; g0 += GU) {
            if (do_pf && tid < GU && g0 + GU + 1 + tid < a.ng) {
                const PkTable tn = a.tabs[g0 + GU + 1 + tid];
                const uint32_t b0 = __umulhi(h_lo, tn.n_buckets), b1 = __umulhi(h_hi, tn.n_buckets);
                if (b1 - b0 < 8192) l2_prefetch_bulk(tn.slots + 4ull * b0, (b1 - b0 + 1) * 32);
            }
            if constexpr (GU == 1) {
                const PkTable t = a.tabs[g0];
#pragma unroll
                for (int j = 0; j < IPT; j++) {
                    const uint32_t i = tid + j * T;
                    if (i < cnt) {
                        const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * __umulhi(h[j], t.n_buckets));
                        const bool hit = pk_bucket_hit<FMT>(v, pk_target<FMT>(canon[j], 0));
                        bits[j] |= (uint32_t)hit << g0;
                        if (!hit && pk_bucket_full<FMT>(v)) {          // deferred walk-on
                            const uint32_t slot = atomicAdd(&q_n, 1u);
                            if (slot < (uint32_t)CAP) {
                                q_key[slot] = canon[j]; q_h[slot] = h[j]; q_meta[slot] = (i << 5) | g0;
                            } else if (pk_lookup<FMT>(t, canon[j], h[j], 32 * a.grp + g0, a.ks)) {      // queue full (pathological)
                                bits[j] |= 1u << g0;
                            }
                        }
                    }
                }
            } else {
            u64x4 v[GU][IPT];
#pragma unroll
            for (int u = 0; u < GU; u++) {
                if (g0 + u < a.ng) {
                    const PkTable t = a.tabs[g0 + u];
#pragma unroll
                    for (int j = 0; j < IPT; j++)
                        if (tid + j * T < cnt) v[u][j] = pk_ld_bucket_ca(t.slots + 4ull * __umulhi(h[j], t.n_buckets));
                }
            }
#pragma unroll
            for (int u = 0; u < GU; u++) {
                const uint32_t g = g0 + u;
                if (g < a.ng) {
#pragma unroll
                    for (int j = 0; j < IPT; j++) {
                        const uint32_t i = tid + j * T;
                        if (i < cnt) {
                            const bool hit = pk_bucket_hit<FMT>(v[u][j], pk_target<FMT>(canon[j], 0));
                            bits[j] |= (uint32_t)hit << g;
                            if (!hit && pk_bucket_full<FMT>(v[u][j])) {          // deferred walk-on
                                const uint32_t slot = atomicAdd(&q_n, 1u);
                                if (slot < (uint32_t)CAP) {
                                    q_key[slot] = canon[j]; q_h[slot] = h[j]; q_meta[slot] = (i << 5) | g;
                                } else if (pk_lookup<FMT>(a.tabs[g], canon[j], h[j], 32 * a.grp + g, a.ks)) {      // queue full (pathological)
                                    bits[j] |= 1u << g;
                                }
                            }
                        }
                    }
                }
            }
            }
        }
        __syncthreads();
        {
            const uint32_t nq = min(q_n, (uint32_t)CAP);
            for (uint32_t e = tid; e < nq; e += T) {
                const uint32_t meta = q_meta[e], g = meta & 31;
                if (pk_lookup<FMT>(a.tabs[g], q_key[e], q_h[e], 32 * a.grp + g, a.ks)) atomicOr(&s_bits[meta >> 5], 1u << g);
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < IPT; j++)
            if (tid + j * T < cnt) bits[j] |= s_bits[tid + j * T];
        if (a.out_list) {
            // append (pos, bits) to the list of bin = pos >> out_shift: rank inside the block by warp match,
            // one atomicAdd per (block, bin) reserves the run
            uint16_t rank[IPT];
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                const bool v = tid + j * T < cnt;
                const uint32_t bin = v ? pos[j] >> a.out_shift : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (v && lane == leader) {
                    old = o_wc[wid][bin];
                    o_wc[wid][bin] = (uint16_t)(old + __popc(peers));
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                rank[j] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1)));
                __syncwarp();
            }
            __syncthreads();
            for (uint32_t b = tid; b < PP_OBINS; b += T) {
                uint32_t run = 0;
#pragma unroll
                for (int ww = 0; ww < T / 32; ww++) {
                    const uint32_t tt = o_wc[ww][b];
                    o_wc[ww][b] = (uint16_t)run;
                    run += tt;
                }
                if (run) o_gb[b] = atomicAdd(&a.out_cursor[(a.grp * PP_OBINS + b) * PP_OCS], run);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                if (tid + j * T < cnt) {
                    const uint32_t bin = pos[j] >> a.out_shift;
                    const uint64_t slot = (((uint64_t)a.grp * PP_OBINS + bin) << a.out_shift) + o_gb[bin] + o_wc[wid][bin] + rank[j];
                    a.out_list[slot] = make_uint2(pos[j], bits[j]);
                }
            }
        } else {
            const uint32_t nb = min(4u, a.nbl - 4 * a.grp);
            const bool al4 = nb == 4 && ((a.row_stride | a.col_offset) & 3) == 0;
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                if (tid + j * T < cnt) {
                    uint8_t *dst = a.rows + (uint64_t)pos[j] * a.row_stride + a.col_offset + 4 * a.grp;
                    if (al4) *(uint32_t *)dst = bits[j];
                    else for (uint32_t qb = 0; qb < nb; qb++) dst[qb] = (uint8_t)(bits[j] >> (8 * qb));
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K3w: table windows staged in shared memory by TMA
// The L1TEX tag stage bounds probe_part_kernel (one 32 B sector per lane and probe: 1.29 G tag lookups, 86 %
// busy on configs[1]). Every probe of a block falls inside one contiguous window of each table (the
// partition's hash range), and with ~4 probes per bucket nearly every sector of that window is needed anyway:
// so one elected thread streams the whole window of genome g into shared memory with a TMA bulk copy
// (cp.async.bulk.shared.global, completion on an mbarrier) and the lanes probe shared memory. DRAM sees
// purely sequential 4-8 KB reads and the L1 tag stage sees nothing.
//
// Windows are staged in two groups of `gsz` genomes (2 * gsz stage buffers): while the lanes probe the
// windows of one group, the copies of the next group are in flight; with <= 2 * gsz genomes in the launch
// everything is issued up front and nothing is recycled. The probe loop is straight-line: a full home
// bucket without the key (the key may sit in a later bucket) is pushed to a small queue, and after the
// group's last genome all lanes drain the queue together — still out of the staged windows, one more
// shared-memory read per step. Only a walk-on past the end of the window (or a table whose window does
// not fit a stage) takes the global path.
// First version (inline walk-on loop, 3-4 single-genome stages, 6 blocks/SM): 5.59 ms on configs[1], issue-bound
// (149 thread-instructions per probe: rematerialised window geometry under a 40-register cap, and a
// second trip round the walk-on loop for 40 % of the warp-iterations) — profiles/r1c_*.
#define PW_MAX_GROUP 4
#define PW_MAX_STAGES (2 * PW_MAX_GROUP)
#define PW_QCAP 384
#define PW_MAX_STAGE_BYTES 12288u

__device__ __forceinline__ uint32_t pw_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pw_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pw_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pw_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pw_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// bounded wait: a copy that never lands (which would be a bug) traps instead of hanging the GPU
__device__ __forceinline__ void pw_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 22)) __trap();
    }
}

// one LDS.128 (the compiler otherwise splits the bucket into lazily evaluated 32-bit loads)
__device__ __forceinline__ uint4 pw_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// a bucket read as its two 16-byte halves (in either order); `last` is the half that holds the last slot
template <int FMT> __device__ __forceinline__ bool pw_hit(const uint4 &A, const uint4 &B, uint64_t target) {
    if (FMT == PK_FMT_S32) {
        const uint32_t t = (uint32_t)target;
        return (A.x == t) | (A.y == t) | (A.z == t) | (A.w == t) | (B.x == t) | (B.y == t) | (B.z == t) | (B.w == t);
    }
    const uint32_t lo = (uint32_t)target, hi = (uint32_t)(target >> 32);
    return ((A.x == lo) & (A.y == hi)) | ((A.z == lo) & (A.w == hi)) | ((B.x == lo) & (B.y == hi)) | ((B.z == lo) & (B.w == hi));
}
template <int FMT> __device__ __forceinline__ bool pw_full(const uint4 &last) {
    return FMT == PK_FMT_S32 ? last.w != PK_EMPTY32 : !(last.z == 0xFFFFFFFFu && last.w == 0xFFFFFFFFu);
}
template <int FMT> struct PwKey { typedef uint64_t type; };
template <> struct PwKey<PK_FMT_S32> { typedef uint32_t type; };      // S32 compares 32-bit slots: keep 32 bits per item
// group tables (pk_device.cuh): slot = [mask 8][key 56]; membership mask of key56 in a bucket read as two halves
#define PK_FMT_GROUP 2
#define PK_FMT_GROUP32 3      // group tables with 32-bit slots (G32, pk_device.cuh)
template <> struct PwKey<PK_FMT_GROUP> { typedef uint64_t type; };
template <> struct PwKey<PK_FMT_GROUP32> { typedef uint32_t type; };
// G32: slot = [mask 8][key 24]; membership mask of key24 in a bucket read as two halves
__device__ __forceinline__ uint32_t pw_umask32(const uint4 &A, const uint4 &B, uint32_t key24) {
    uint32_t m = 0;
    if ((A.x & PK_G32_KEY_MASK) == key24) m |= A.x >> PK_G32_MASK_SHIFT;
    if ((A.y & PK_G32_KEY_MASK) == key24) m |= A.y >> PK_G32_MASK_SHIFT;
    if ((A.z & PK_G32_KEY_MASK) == key24) m |= A.z >> PK_G32_MASK_SHIFT;
    if ((A.w & PK_G32_KEY_MASK) == key24) m |= A.w >> PK_G32_MASK_SHIFT;
    if ((B.x & PK_G32_KEY_MASK) == key24) m |= B.x >> PK_G32_MASK_SHIFT;
    if ((B.y & PK_G32_KEY_MASK) == key24) m |= B.y >> PK_G32_MASK_SHIFT;
    if ((B.z & PK_G32_KEY_MASK) == key24) m |= B.z >> PK_G32_MASK_SHIFT;
    if ((B.w & PK_G32_KEY_MASK) == key24) m |= B.w >> PK_G32_MASK_SHIFT;
    return m;
}
__device__ __forceinline__ uint32_t pw_umask(const uint4 &A, const uint4 &B, uint64_t key56) {
    const uint32_t lo = (uint32_t)key56, hi = (uint32_t)(key56 >> 32);
    uint32_t m = 0;
    if (A.x == lo && (A.y & 0x00FFFFFFu) == hi) m |= A.y >> 24;
    if (A.z == lo && (A.w & 0x00FFFFFFu) == hi) m |= A.w >> 24;
    if (B.x == lo && (B.y & 0x00FFFFFFu) == hi) m |= B.y >> 24;
    if (B.z == lo && (B.w & 0x00FFFFFFu) == hi) m |= B.w >> 24;
    return m;
}

template <int T, int IPT, int FMT, int MINB, int CHUNKED>
__global__ void __launch_bounds__(T, MINB) probe_win_kernel(const __grid_constant__ ProbeArgs a, const uint32_t stage_bytes,
                                                            const uint32_t gsz, const uint32_t n_stages) {
    constexpr int CAP = T * IPT;
    typedef typename PwKey<FMT>::type key_t;
    extern __shared__ __align__(128) uint8_t s_win[];          // [2 * gsz][stage_bytes]
    __shared__ __align__(8) unsigned long long s_bar[PW_MAX_STAGES];
    __shared__ uint32_t s_bits[CAP];               // results of the deferred walk-ons, by item
    __shared__ key_t q_key[PW_QCAP];                                       // deferred queue: key (as in the home bucket),
    __shared__ uint32_t q_pos[PW_QCAP], q_h[PW_QCAP], q_meta[PW_QCAP];   //   position, hash, (item << 5 | genome)
    __shared__ uint32_t q_n[2];
    __shared__ uint16_t o_wc[T / 32][PP_OBINS];
    __shared__ uint32_t o_gb[PP_FBINS], o_cnt[PP_FBINS];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr bool GRP = FMT == PK_FMT_GROUP || FMT == PK_FMT_GROUP32;       // group tables: a probe returns an 8-bit membership mask
    constexpr bool G32 = FMT == PK_FMT_GROUP32;
    const uint32_t win0 = pw_smem(s_win), bar0 = pw_smem(s_bar);
    const uint32_t xoff = (lane & 1) * 16;          // odd lanes read the bucket halves in the other order: spreads the banks
    // thread s initialises the mbarrier of stage s and issues the first copy into it straight away (other threads
    // first touch that barrier after a __syncthreads that follows); copies are issued by several threads in
    // parallel: one thread issuing 8 windows back to back held the whole block at the first barrier for > 1 us
    if (tid < n_stages) {
        pw_mbar_init(bar0 + 8 * tid, 1);
        pw_mbar_fence_init();
    }
    uint32_t par = 0;            // bit s: the phase parity the next wait on stage s expects
    for (uint64_t q = blockIdx.x; q < a.n_regions; q += gridDim.x) {
        const uint32_t cnt = min(a.counts[q], a.cap);
        if (cnt == 0) continue;
        const uint32_t h_lo = (uint32_t)(q << (32 - a.pb));
        const uint32_t h_hi = (uint32_t)(((q + 1) << (32 - a.pb)) - 1);
        // window of genome g = buckets [b0, b1] of its table; streamed into stage g % n_stages when it fits
        // piece g -> its table and its bucket range [cs, ce) inside this partition's window of that table
        auto geom = [&](uint32_t g, uint32_t &cs, uint32_t &ce) -> PkTable {
            const PkTable t = a.tabs[GRP ? a.t_of[g] : g];
            const uint32_t b0 = __umulhi(h_lo, t.n_buckets), b1 = __umulhi(h_hi, t.n_buckets);
            if constexpr (!CHUNKED) { cs = b0; ce = b1 + 1; }
            else { cs = min(b0 + a.c_of[g] * a.chunk_buckets, b1 + 1); ce = min(cs + a.chunk_buckets, b1 + 1); }
            return t;
        };
        auto issue = [&](uint32_t g) {
            uint32_t cs, ce;
            const PkTable t = geom(g, cs, ce);
            const uint32_t bytes = (ce - cs) * 32;
            if (bytes && bytes <= stage_bytes) {
                const uint32_t s = g & (n_stages - 1);
                pw_mbar_expect_tx(bar0 + 8 * s, bytes);
                pw_bulk_g2s(win0 + s * stage_bytes, t.slots + 4ull * cs, bytes, bar0 + 8 * s);
            }
        };
        if (tid < n_stages && tid < a.ng) issue(tid);
        const uint2 *src = a.buf + q * (uint64_t)a.cap;
        uint2 it[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) it[j] = tid + j * T < cnt ? src[tid + j * T] : make_uint2(0, 0);
        if (tid == 0) { q_n[0] = 0; q_n[1] = 0; }
        // s_bits (results of deferred walk-ons) is zeroed per item by the thread that defers it, not in bulk; the warp-ranking
        // counters only when that ranking is in use; the bin counters only as far as this launch has bins
        const bool atomic_rank = a.out_fine || a.rank_atomic == 1;
        if (a.out_list && !atomic_rank)
            for (uint32_t i = tid; i < (T / 32) * PP_OBINS / 2; i += T) ((uint32_t *)&o_wc[0][0])[i] = 0;
        for (uint32_t i = tid; i < a.n_bins; i += T) o_cnt[i] = 0;
        uint32_t dmask = 0;      // bit j: item j of this thread was deferred (its s_bits entry is live)
        key_t key[IPT];          // S64: the canonical k-mer; S32: the slot value it has in its home bucket
        uint32_t h[IPT], pos[IPT], bits[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            key[j] = 0; h[j] = it[j].x; pos[j] = it[j].y; bits[j] = 0;
            if (tid + j * T < cnt) {
                if (G32 && a.compact) {
                    // compact item: hash and key remainder travel in the item, the sequence is not read
                    const uint32_t r20 = ((it[j].x & 0xFFFFu) << 4) | (it[j].y >> PT_CI_POS_BITS);
                    pos[j] = it[j].y & PT_CI_POS_MASK;
                    h[j] = ((uint32_t)(q >> a.pb2) << 23) | ((it[j].x >> 16) << (32 - a.eb)) | (pk_mix32(r20) >> a.eb);
                    key[j] = (key_t)(r20 << PK_S32_DISP_BITS);
                } else {
                    const uint64_t canon = pk_canon_at(a.words, a.p0 + it[j].y, a.ks.k);
                    if constexpr (G32) key[j] = pk_g32_key(canon, 0);
                    else if constexpr (GRP) key[j] = pk_u_key(canon, 0);
                    else key[j] = (key_t)pk_target<GRP ? PK_FMT_S64 : FMT>(canon, 0);
                }
            }
        }
        __syncthreads();
        // the whole answer of table g for one k-mer through global memory, as row bits of this launch
        auto full_lookup = [&](const PkTable &t, uint64_t canon, uint32_t hh, uint32_t g) -> uint32_t {
            if constexpr (GRP) {
                const uint32_t ngg = min(PK_U_GROUP, a.n_genomes - PK_U_GROUP * g);
                return pk_group_lookup(t, canon, hh, a.g_first + PK_U_GROUP * g, ngg, a.ks) << (PK_U_GROUP * g);
            } else {
                return (uint32_t)pk_lookup<GRP ? PK_FMT_S64 : FMT>(t, canon, hh, a.g_first + g, a.ks) << g;
            }
        };
        // ---- probe, group by group, out of the staged windows
        uint32_t grp_i = 0;
        for (uint32_t g0 = 0; g0 < a.ng; g0 += gsz, grp_i++) {
            const uint32_t g1 = min(g0 + gsz, a.ng);
            uint32_t *qn = &q_n[grp_i & 1];
            // a full home bucket without the key: the key may sit in a later bucket, decide after the group
            auto defer = [&](int j, uint32_t i, uint32_t g, const PkTable &t) {
                const uint32_t slot = atomicAdd(qn, 1u);
                if (slot < (uint32_t)PW_QCAP) {
                    if (!((dmask >> j) & 1)) { s_bits[i] = 0; dmask |= 1u << j; }
                    q_key[slot] = key[j]; q_pos[slot] = pos[j]; q_h[slot] = h[j]; q_meta[slot] = (i << 5) | g;
                } else {                    // queue full (pathological)
                    bits[j] |= full_lookup(t, pk_canon_at(a.words, a.p0 + pos[j], a.ks.k), h[j], GRP ? a.t_of[g] : g);
                }
            };
            for (uint32_t g = g0; g < g1; g++) {
                uint32_t cs, ce;
                const PkTable t = geom(g, cs, ce);
                const uint32_t s = g & (n_stages - 1);
                const uint32_t gbit = 1u << g, gsh = GRP ? PK_U_GROUP * a.t_of[g] : 0;
                (void)gbit; (void)gsh;
                if (CHUNKED && ce == cs) continue;                    // empty piece (short window): nothing was copied
                if ((ce - cs) * 32 <= stage_bytes) {
                    pw_mbar_wait(bar0 + 8 * s, (par >> s) & 1);
                    par ^= 1u << s;
                    // + 32 * bucket = the bucket's first (even lanes) / second (odd lanes) half; kept opaque so that it
                    // stays in a register instead of being re-derived per item
                    uint32_t wbase = win0 + s * stage_bytes - cs * 32 + xoff, nb = t.n_buckets, nbp = ce - cs;
                    asm volatile("" : "+r"(wbase), "+r"(nb), "+r"(nbp));
#pragma unroll
                    for (int j = 0; j < IPT; j++) {
                        const uint32_t i = tid + j * T;
                        const uint32_t b = __umulhi(h[j], nb);
                        if (i < cnt && (!CHUNKED || b - cs < nbp)) {  // the item's home bucket lies in this piece
                            const uint32_t wa = wbase + b * 32;
                            const uint4 A = pw_lds128(wa), B = pw_lds128(wa ^ 16);
                            if constexpr (G32) {
                                const uint32_t m = pw_umask32(A, B, key[j]);
                                if (m) bits[j] |= m << gsh;
                                else if (pw_full<PK_FMT_S32>(xoff ? A : B)) defer(j, i, g, t);
                            } else if constexpr (GRP) {
                                const uint32_t m = pw_umask(A, B, key[j]);
                                if (m) bits[j] |= m << gsh;
                                else if (pw_full<PK_FMT_S64>(xoff ? A : B)) defer(j, i, g, t);
                            } else {
                                const bool hit = pw_hit<GRP ? PK_FMT_S64 : FMT>(A, B, key[j]);
                                if (hit) bits[j] |= gbit;
                                else if (pw_full<GRP ? PK_FMT_S64 : FMT>(xoff ? A : B)) defer(j, i, g, t);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < IPT; j++) {
                        const uint32_t i = tid + j * T;
                        if (i < cnt) {
                            const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * __umulhi(h[j], t.n_buckets));
                            if constexpr (G32) {
                                const uint32_t m = pk_g32_bucket_mask(v, key[j]);
                                if (m) bits[j] |= m << gsh;
                                else if ((uint32_t)(v.d >> 32) != PK_EMPTY32) defer(j, i, g, t);
                            } else if constexpr (GRP) {
                                const uint32_t m = pk_u_bucket_mask(v, key[j]);
                                if (m) bits[j] |= m << gsh;
                                else if (v.d != PK_EMPTY) defer(j, i, g, t);
                            } else {
                                if (pk_bucket_hit<GRP ? PK_FMT_S64 : FMT>(v, key[j])) bits[j] |= gbit;
                                else if (pk_bucket_full<GRP ? PK_FMT_S64 : FMT>(v)) defer(j, i, g, t);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            {   // drain the queue: walk on through the staged window, bucket by bucket
                const uint32_t nq = min(*qn, (uint32_t)PW_QCAP);
                for (uint32_t e = tid; e < nq; e += T) {
                    const uint32_t meta = q_meta[e], g = meta & 31, hh = q_h[e], u = GRP ? a.t_of[g] : g;
                    const key_t kk = q_key[e];
                    uint32_t cs, ce;
                    const PkTable t = geom(g, cs, ce);
                    const uint32_t nbk = ce - cs, off = __umulhi(hh, t.n_buckets) - cs;
                    const uint32_t maxd = GRP ? pk_u_max_disp(t.n_buckets) : pk_max_disp<GRP ? PK_FMT_S64 : FMT>(t.n_buckets);
                    uint32_t found = 0;     // row bits of this launch the k-mer gets from table g
                    bool decided = false;   // false: leave the window -> global lookup
                    if (nbk * 32 <= stage_bytes) {
                        const uint32_t wb = win0 + (g & (n_stages - 1)) * stage_bytes + xoff;
                        for (uint32_t r = 1; r <= maxd && off + r < nbk; r++) {
                            const uint32_t wa = wb + (off + r) * 32;
                            const uint4 A = pw_lds128(wa), B = pw_lds128(wa ^ 16);
                            // S32 / group: the slot value at displacement r is the home value + r; S64: the k-mer itself
                            if constexpr (G32) {
                                const uint32_t m = pw_umask32(A, B, (uint32_t)kk + r);
                                if (m) { found = m << (PK_U_GROUP * u); decided = true; break; }
                                if (!pw_full<PK_FMT_S32>(xoff ? A : B)) { decided = true; break; }
                            } else if constexpr (GRP) {
                                const uint32_t m = pw_umask(A, B, (uint64_t)kk + r);
                                if (m) { found = m << (PK_U_GROUP * u); decided = true; break; }
                                if (!pw_full<PK_FMT_S64>(xoff ? A : B)) { decided = true; break; }
                            } else {
                                if (pw_hit<GRP ? PK_FMT_S64 : FMT>(A, B, FMT == PK_FMT_S32 ? (uint64_t)kk + r : (uint64_t)kk)) { found = 1u << u; decided = true; break; }
                                if (!pw_full<GRP ? PK_FMT_S64 : FMT>(xoff ? A : B)) { decided = true; break; }
                            }
                        }
                    }
                    if (!decided)       // rare: past the window's end, > 14 full buckets in a row (stash), or no window
                        found = full_lookup(t, pk_canon_at(a.words, a.p0 + q_pos[e], a.ks.k), hh, u);
                    if (found) atomicOr(&s_bits[meta >> 5], found);
                }
            }
            __syncthreads();            // the group's windows and the queue are free again
            if (tid == 0) *qn = 0;
            if (tid < gsz && g0 + n_stages + tid < a.ng) issue(g0 + n_stages + tid);     // refill the group's stages in parallel
        }
        // (the last group's second barrier also orders s_bits)
#pragma unroll
        for (int j = 0; j < IPT; j++)
            if ((dmask >> j) & 1) bits[j] |= s_bits[tid + j * T];
        if (a.out_list && a.rank_atomic == 2) {
            // ABLATION (timing experiments only; rows are wrong): no output at all
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < IPT; j++) acc |= bits[j];
            if (acc == 0xdeadbeefu) a.out_cursor[0] = acc;
        } else if (a.out_list && (a.out_fine || a.rank_atomic == 1)) {
            // rank inside (block, bin) by one shared-memory atomicAdd per item, one global atomicAdd per (block, bin)
            // reserves the run in the bin's list. (Reserving at the top of the block, to overlap the atomics' round trip
            // with the probes, was slower: 2.43 vs 2.19 ms, profiles/r2k_sweep_early_reservation.json.) Fine mode
            // (one-byte rows): (position in bin) << 8 | bits in 4 bytes, <= 512 bins; else (position, bits) in 8 bytes
            uint32_t rk[IPT];
#pragma unroll
            for (int j = 0; j < IPT; j++)
                rk[j] = tid + j * T < cnt ? atomicAdd(&o_cnt[pos[j] >> a.out_shift], 1u) : 0u;
            __syncthreads();
            for (uint32_t b = tid; b < a.n_bins; b += T) {
                const uint32_t c = o_cnt[b];
                if (c) o_gb[b] = atomicAdd(&a.out_cursor[(a.out_fine ? b : a.grp * PP_OBINS + b) * PP_OCS], c);
            }
            __syncthreads();
            uint32_t *list4 = (uint32_t *)a.out_list;
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                if (tid + j * T < cnt) {
                    const uint32_t bin = pos[j] >> a.out_shift;
                    if (a.out_fine)
                        list4[((uint64_t)bin << a.out_shift) + o_gb[bin] + rk[j]] = ((pos[j] & ((1u << a.out_shift) - 1)) << 8) | (bits[j] & 0xffu);
                    else
                        a.out_list[(((uint64_t)a.grp * PP_OBINS + bin) << a.out_shift) + o_gb[bin] + rk[j]] = make_uint2(pos[j], bits[j]);
                }
            }
        } else if (a.out_list) {
            uint16_t rank[IPT];
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                const bool v = tid + j * T < cnt;
                const uint32_t bin = v ? pos[j] >> a.out_shift : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (v && lane == leader) {
                    old = o_wc[wid][bin];
                    o_wc[wid][bin] = (uint16_t)(old + __popc(peers));
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                rank[j] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1)));
                __syncwarp();
            }
            __syncthreads();
            for (uint32_t b = tid; b < PP_OBINS; b += T) {
                uint32_t run = 0;
#pragma unroll
                for (int ww = 0; ww < T / 32; ww++) {
                    const uint32_t tt = o_wc[ww][b];
                    o_wc[ww][b] = (uint16_t)run;
                    run += tt;
                }
                if (run) o_gb[b] = atomicAdd(&a.out_cursor[(a.grp * PP_OBINS + b) * PP_OCS], run);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                if (tid + j * T < cnt) {
                    const uint32_t bin = pos[j] >> a.out_shift;
                    const uint64_t slot = (((uint64_t)a.grp * PP_OBINS + bin) << a.out_shift) + o_gb[bin] + o_wc[wid][bin] + rank[j];
                    a.out_list[slot] = make_uint2(pos[j], bits[j]);
                }
            }
        } else {
            const uint32_t nb = min(4u, a.nbl - 4 * a.grp);
            const bool al4 = nb == 4 && ((a.row_stride | a.col_offset) & 3) == 0;
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                if (tid + j * T < cnt) {
                    uint8_t *dst = a.rows + (uint64_t)pos[j] * a.row_stride + a.col_offset + 4 * a.grp;
                    if (al4) *(uint32_t *)dst = bits[j];
                    else for (uint32_t qb = 0; qb < nb; qb++) dst[qb] = (uint8_t)(bits[j] >> (8 * qb));
                }
            }
        }
        __syncthreads();
    }
}

// K3, lean form for the configuration the metric is quoted on: ONE 32-bit-slot group table (<= 8 genomes, one-byte
// rows), compact items, fine position bins, the partition's whole window in one stage. Same results as
// probe_win_kernel<.., PK_FMT_GROUP32, .., 0> on such a launch; what it drops is that kernel's generality, which the
// ncu source view showed to cost half its instructions and its worst stalls (profiles/r2r: 1.04 G warp instructions,
// the row bits in LOCAL memory because a lambda indexed them, per-group window geometry re-read from the constant
// bank, a deferred-walk-on queue with two more barriers, and the reservation atomics' round trip exposed between two
// barriers):
//   * a key missing from a full home bucket walks on INLINE through the staged window (rare: keys are displaced
//     from their home bucket in < 1 % of the cases), leaving the window or reaching the maximal displacement (the
//     stash) through the global lookup;
//   * a slot matches when (slot ^ key24) has no low 24 bits: one LOP3 that sets a predicate + one predicated OR per
//     slot, the mask shifted out once per bucket;
//   * ranks inside the position bins are taken BEFORE the probe (they depend on the positions only) and with EARLY
//     the per-bin reservations (one global atomicAdd per bin the block touches) are issued before the wait on the
//     window, so their round trip runs under the TMA copy and the probes; their results are first needed after.
template <int T, int IPT, int MINB, int EARLY>
__global__ void __launch_bounds__(T, MINB) probe_g32c_kernel(const __grid_constant__ ProbeArgs a, const uint32_t stage_bytes) {
    static_assert(2 * T >= PP_FBINS, "two rounds of the block cover the fine bins");
    extern __shared__ __align__(128) uint8_t s_win[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t o_cnt[PP_FBINS], o_gb[PP_FBINS];
    const uint32_t tid = threadIdx.x;
    const uint32_t win0 = pw_smem(s_win), bar = pw_smem(&s_bar);
    const uint32_t xoff = (tid & 1) * 16;           // odd lanes read the bucket halves in the other order: spreads the banks
    const PkTable t = a.tabs[0];
    const uint32_t nb = t.n_buckets, sh = a.out_shift, lowmask = (1u << sh) - 1;
    const uint32_t ngg = min(PK_U_GROUP, a.n_genomes);
    uint32_t *list4 = (uint32_t *)a.out_list;
    if (tid == 0) { pw_mbar_init(bar, 1); pw_mbar_fence_init(); }
    uint32_t par = 0;
    for (uint64_t q = blockIdx.x; q < a.n_regions; q += gridDim.x) {
        // the region's fill, the window copy and the first half of the item loads are issued together: none of them
        // depends on another (item slots beyond the fill hold stale items of the same buffer: loaded, never used)
        const uint32_t craw = a.counts[q];
        const uint32_t b0 = __umulhi((uint32_t)(q << (32 - a.pb)), nb);
        const uint32_t nbk = __umulhi((uint32_t)(((q + 1) << (32 - a.pb)) - 1), nb) - b0 + 1;
        const bool staged = nbk * 32 <= stage_bytes;            // the host sizes the stage for every window; kept as a guard
        if (tid == 0 && staged) {
            pw_mbar_expect_tx(bar, nbk * 32);
            pw_bulk_g2s(win0, t.slots + 4ull * b0, nbk * 32, bar);
        }
        const uint2 *src = a.buf + q * (uint64_t)a.cap;
        uint2 it[IPT];
#pragma unroll
        for (int j = 0; j < IPT / 2; j++) it[j] = src[tid + j * T];
        const uint32_t cnt = min(craw, a.cap);
        if (cnt == 0) {                                         // (uniform) nothing to probe: let the copy land, then move on
            __syncthreads();                                    // (first trip: the mbarrier's initialisation becomes visible)
            if (staged) { pw_mbar_wait(bar, par); par ^= 1; }
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int j = IPT / 2; j < IPT; j++) it[j] = tid + j * T < cnt ? src[tid + j * T] : make_uint2(0, 0);
        for (uint32_t i = tid; i < a.n_bins; i += T) o_cnt[i] = 0;
        __syncthreads();                                        // bin counters zeroed, mbarrier initialised
        uint32_t rk[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++)
            rk[j] = tid + j * T < cnt ? atomicAdd(&o_cnt[(it[j].y & PT_CI_POS_MASK) >> sh], 1u) : 0u;
        uint32_t gb0 = 0, gb1 = 0;
        if (EARLY) {
            __syncthreads();
            if (tid < a.n_bins) { const uint32_t c = o_cnt[tid]; if (c) gb0 = atomicAdd(&a.out_cursor[tid * PP_OCS], c); }
            if (tid + T < a.n_bins) { const uint32_t c = o_cnt[tid + T]; if (c) gb1 = atomicAdd(&a.out_cursor[(tid + T) * PP_OCS], c); }
        }
        // the hash of a compact item (PartSmem): top 9 bits = its coarse region, then the stored bits, then mix32(R20) >> eb
        const uint32_t h_top = (uint32_t)(q >> a.pb2) << 23;
        if (staged) { pw_mbar_wait(bar, par); par ^= 1; }
        uint32_t wbase = win0 + xoff - b0 * 32;                 // + 32 * bucket = first (even lanes) / second (odd lanes) half
        asm volatile("" : "+r"(wbase));
        uint32_t bits[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            bits[j] = 0;
            if (tid + j * T < cnt) {
                const uint32_t r20 = ((it[j].x & 0xFFFFu) << 4) | (it[j].y >> PT_CI_POS_BITS);
                const uint32_t h = h_top | ((it[j].x >> 16) << (32 - a.eb)) | (pk_mix32(r20) >> a.eb);
                const uint32_t key = r20 << PK_S32_DISP_BITS;
                const uint32_t b = __umulhi(h, nb);
                uint32_t acc = 0, last;
                if (staged) {
                    const uint32_t wa = wbase + b * 32;
                    const uint4 A = pw_lds128(wa), B = pw_lds128(wa ^ 16);
                    if (((A.x ^ key) & PK_G32_KEY_MASK) == 0) acc |= A.x;
                    if (((A.y ^ key) & PK_G32_KEY_MASK) == 0) acc |= A.y;
                    if (((A.z ^ key) & PK_G32_KEY_MASK) == 0) acc |= A.z;
                    if (((A.w ^ key) & PK_G32_KEY_MASK) == 0) acc |= A.w;
                    if (((B.x ^ key) & PK_G32_KEY_MASK) == 0) acc |= B.x;
                    if (((B.y ^ key) & PK_G32_KEY_MASK) == 0) acc |= B.y;
                    if (((B.z ^ key) & PK_G32_KEY_MASK) == 0) acc |= B.z;
                    if (((B.w ^ key) & PK_G32_KEY_MASK) == 0) acc |= B.w;
                    last = xoff ? A.w : B.w;
                } else {
                    const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * b);
                    acc = pk_g32_bucket_mask(v, key) << PK_G32_MASK_SHIFT;
                    last = (uint32_t)(v.d >> 32);
                }
                if (acc == 0 && last != PK_EMPTY32) {
                    // full home bucket without the key: the slot at displacement r holds key + r
                    const uint32_t maxd = pk_u_max_disp(nb);
                    bool decided = false;
                    if (staged)
                        for (uint32_t r = 1; r <= maxd && b - b0 + r < nbk; r++) {
                            const uint32_t wa = wbase + (b + r) * 32;
                            const uint4 A = pw_lds128(wa), B = pw_lds128(wa ^ 16);
                            const uint32_t m = pw_umask32(A, B, key + r);
                            if (m) { acc = m << PK_G32_MASK_SHIFT; decided = true; break; }
                            if ((xoff ? A.w : B.w) == PK_EMPTY32) { decided = true; break; }
                        }
                    if (!decided)       // past the window's end, or 14 full buckets in a row (stash)
                        acc = pk_group_lookup(t, pk_canon_at(a.words, a.p0 + (it[j].y & PT_CI_POS_MASK), a.ks.k), h, a.g_first, ngg, a.ks)
                              << PK_G32_MASK_SHIFT;
                }
                bits[j] = acc >> PK_G32_MASK_SHIFT;
            }
        }
        if (!EARLY) {
            __syncthreads();
            if (tid < a.n_bins) { const uint32_t c = o_cnt[tid]; if (c) gb0 = atomicAdd(&a.out_cursor[tid * PP_OCS], c); }
            if (tid + T < a.n_bins) { const uint32_t c = o_cnt[tid + T]; if (c) gb1 = atomicAdd(&a.out_cursor[(tid + T) * PP_OCS], c); }
        }
        if (tid < a.n_bins) o_gb[tid] = gb0;
        if (tid + T < a.n_bins) o_gb[tid + T] = gb1;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < IPT; j++)
            if (tid + j * T < cnt) {
                const uint32_t pos = it[j].y & PT_CI_POS_MASK;
                // slot = start of the bin (pos with its low bits cleared) + the block's run in the bin + the rank in the run
                list4[(pos & ~lowmask) + o_gb[pos >> sh] + rk[j]] = ((pos & lowmask) << 8) | (bits[j] & 0xffu);
            }
        __syncthreads();            // the window, the counters and o_gb are free again
    }
}
// K3 straight out of the COARSE regions — no K2, no shared-memory windows (tune "k3_l2"; the default for one
// 32-bit-slot group table with compact items and fine bins). The 2^9 coarse regions are processed in order, a few
// hundred blocks per region, so at any time the blocks in flight probe three or four coarse windows of the table
// (5 MB each on configs[1]): they stay L2-resident and every probe is one LDG.E.256 that hits L2 (or brings its
// sector in: the table is read from DRAM once, and only the sectors some probe needs). Same items, ranks, reservations
// and result lists as probe_g32c_kernel; a block takes T * IPT consecutive items of its region whatever the plan's
// fine capacity is, so larger blocks mean longer runs per position bin (fewer reservations, fuller store sectors).
// Measured on configs[1] (profiles/r2ad_sweep_k3_l2_prefetch.json): K2 0.52 + K3 1.18 -> K3 1.14 ms, stage 2.69 ->
// 2.12 ms; asking the TMA unit to prefetch the window of the region 1..5 ahead into L2 (cp.async.bulk.prefetch.L2)
// changed nothing (1.15-1.17 ms) and was dropped.
// Block shape (profiles/r2ae_sweep_k3_l2_variants.json): occupancy beats run length — 256 threads x 4 items at 8 blocks/SM
// (64 warps, <= 32 registers) 1.04 ms, 384 x 4 at 4 blocks/SM 1.14, 512 x 8 1.30, 1024 x 4 1.59. Taking the slots
// directly with one global atomicAdd per item instead of ranking inside the block: 3.78 ms (135 M atomics on ~500
// addresses; profiles/r2af_sweep_k3_l2_direct_atomics.json) — dropped.
template <int T, int IPT, int MINB>
__global__ void __launch_bounds__(T, MINB) probe_g32l2_kernel(const __grid_constant__ ProbeArgs a) {
    static_assert(2 * T >= PP_FBINS, "two rounds of the block cover the fine bins");
    constexpr uint32_t CAP = T * IPT;
    __shared__ uint32_t o_cnt[PP_FBINS], o_gb[PP_FBINS];
    const uint32_t tid = threadIdx.x, r = blockIdx.y;
    const PkTable t = a.tabs[0];
    const uint32_t nb = t.n_buckets, sh = a.out_shift, lowmask = (1u << sh) - 1;
    const uint32_t ngg = min(PK_U_GROUP, a.n_genomes);
    uint32_t *list4 = (uint32_t *)a.out_list;
    const uint32_t cnt_r = min(a.counts[r], a.cap);
    const uint32_t base = blockIdx.x * CAP;
    if (base >= cnt_r) return;
    const uint32_t cnt = min(CAP, cnt_r - base);
    const uint2 *src = a.buf + (uint64_t)r * a.cap + base;
    uint2 it[IPT];
#pragma unroll
    for (int j = 0; j < IPT; j++) it[j] = tid + j * T < cnt ? src[tid + j * T] : make_uint2(0, 0);
    for (uint32_t i = tid; i < a.n_bins; i += T) o_cnt[i] = 0;
    __syncthreads();
    // Ranks inside the position bins, WARP-AGGREGATED: K1 appends to a coarse region tile after tile, so a region's items
    // are in position order and the ~1000 consecutive items of a block fall into two or three bins — one shared-memory
    // atomicAdd per item meant 32 lanes on one address (ncu: 29 wavefronts per ATOMS, 124 M of the kernel's 138 M
    // shared wavefronts, profiles/r2ag_ncu_k1_k3l2_k4.txt). The lanes that share a bin elect a leader that adds
    // their number once. (The same order makes a block's result runs hundreds of items long: its stores coalesce.)
    const uint32_t lane = tid & 31;
    uint32_t rk[IPT];
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const bool valid = tid + j * T < cnt;
        const uint32_t bin = valid ? (it[j].y & PT_CI_POS_MASK) >> sh : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t first = 0;
        if (valid && lane == leader) first = atomicAdd(&o_cnt[bin], (uint32_t)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        rk[j] = first + __popc(peers & ((1u << lane) - 1));
    }
    // the probes' loads are issued before the ranks are complete: nothing below the barrier depends on them
    const uint32_t h_top = r << (32 - a.pb);
    uint32_t bits[IPT];
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        bits[j] = 0;
        if (tid + j * T < cnt) {
            const uint32_t r20 = ((it[j].x & 0xFFFFu) << 4) | (it[j].y >> PT_CI_POS_BITS);
            const uint32_t h = h_top | ((it[j].x >> 16) << (32 - a.eb)) | (pk_mix32(r20) >> a.eb);
            const uint32_t key = r20 << PK_S32_DISP_BITS;
            const u64x4 v = pk_ld_bucket(t.slots + 4ull * __umulhi(h, nb));
            uint32_t m = pk_g32_bucket_mask(v, key);
            if (m == 0 && (uint32_t)(v.d >> 32) != PK_EMPTY32)       // full home bucket without the key: walk on (rare)
                m = pk_group_lookup(t, pk_canon_at(a.words, a.p0 + (it[j].y & PT_CI_POS_MASK), a.ks.k), h, a.g_first, ngg, a.ks);
            bits[j] = m;
        }
    }
    __syncthreads();
    uint32_t gb0 = 0, gb1 = 0;
    if (tid < a.n_bins) { const uint32_t c = o_cnt[tid]; if (c) gb0 = atomicAdd(&a.out_cursor[tid * PP_OCS], c); }
    if (tid + T < a.n_bins) { const uint32_t c = o_cnt[tid + T]; if (c) gb1 = atomicAdd(&a.out_cursor[(tid + T) * PP_OCS], c); }
    if (tid < a.n_bins) o_gb[tid] = gb0;
    if (tid + T < a.n_bins) o_gb[tid + T] = gb1;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < IPT; j++)
        if (tid + j * T < cnt) {
            const uint32_t pos = it[j].y & PT_CI_POS_MASK;
            list4[(pos & ~lowmask) + o_gb[pos >> sh] + rk[j]] = ((pos & lowmask) << 8) | (bits[j] & 0xffu);
        }
}

struct K3L2Variant { int threads, cap; void (*fn)(ProbeArgs); };
#define K3G(T, IPT, MINB) {T, T * IPT, probe_g32l2_kernel<T, IPT, MINB>}
static const K3L2Variant k3g_variants[] = {
    K3G(384, 4, 4),         // 1 (tune.k3_l2 = index + 1)
    K3G(512, 4, 3),         // 2: 2048 items per block
    K3G(512, 8, 2),         // 3: 4096
    K3G(256, 8, 4),         // 4: 2048, 8 items per thread
    K3G(1024, 4, 1),        // 5: 4096, one block per SM
    K3G(1024, 8, 1),        // 6: 8192
    K3G(256, 4, 8),         // 7: 1024 items, 8 blocks/SM (<= 32 registers) — the default
    K3G(512, 2, 4),         // 8: 1024 items, 2 per thread
    K3G(256, 3, 8),         // 9: 768 items
    K3G(256, 2, 8),         // 10: 512 items
};
int pk_part_n_gvariants(void) { return (int)(sizeof k3g_variants / sizeof k3g_variants[0]); }
struct K3LeanVariant { int threads, cap; void (*fn)(ProbeArgs, uint32_t); };
#define K3L(T, IPT, MINB, EARLY) {T, T * IPT, probe_g32c_kernel<T, IPT, MINB, EARLY>}
static const K3LeanVariant k3l_variants[] = {
    K3L(384, 4, 4, 1),      // 1 (tune.lean = index + 1): capacity 1536, reservations before the probes
    K3L(384, 4, 4, 0),      // 2: ... after the probes
    K3L(384, 4, 5, 1),      // 3: <= 32 registers, 5 blocks/SM
    K3L(512, 3, 3, 1),      // 4
    K3L(512, 3, 4, 1),      // 5: <= 32 registers
    K3L(256, 6, 6, 1),      // 6
};
int pk_part_n_lvariants(void) { return (int)(sizeof k3l_variants / sizeof k3l_variants[0]); }

// variants of K3 selectable at run time (PK_K3_VARIANT) while the design is being tuned
struct K3Variant { int threads, cap; void (*fn[2])(ProbeArgs); };   // fn[fmt]
#define K3V(T, IPT, SORT, MINB, GU) {T, T * IPT, {probe_part_kernel<T, IPT, PK_FMT_S64, SORT, MINB, GU>, probe_part_kernel<T, IPT, PK_FMT_S32, SORT, MINB, GU>}}
static const K3Variant k3_variants[] = {
    K3V(256, 3, 0, 6, 1),   // 0: unsorted, 6 blocks/SM, one genome in flight
    K3V(256, 3, 1, 5, 1),   // 1: bucket-sorted
    K3V(256, 3, 0, 8, 1),   // 2: 8 blocks/SM (32 registers)
    K3V(256, 3, 0, 5, 2),   // 3: two genomes in flight
    K3V(256, 3, 0, 4, 2),   // 4
    K3V(128, 6, 0, 12, 1),  // 5: 128-thread blocks
    K3V(256, 6, 0, 4, 1),   // 6: cap 1536 (2^17 partitions)
    K3V(256, 4, 0, 5, 1),   // 7: cap 1024
    K3V(256, 3, 0, 7, 1),   // 8
};
// -1 = auto: per launch (group of <= 32 genomes), bucket-sorted blocks once the sort is amortised over
// >= 16 genomes (configs[2]: 18.9 vs 19.9 ms), unsorted below (configs[1]: 5.5 vs 6.3 ms). Variants 0 and 1
// share the block capacity, hence the partition plan.
int pk_part_n_variants(void) { return (int)(sizeof k3_variants / sizeof k3_variants[0]); }
static const K3Variant &k3_pick(const PkPartTune &tu, uint32_t n_genomes_in_launch) {
    if (tu.variant >= 0 && tu.variant < pk_part_n_variants()) return k3_variants[tu.variant];
    return k3_variants[n_genomes_in_launch >= 16 ? 1 : 0];
}

// window (TMA-staged) variants of K3; same block capacity as variants 0/1, so the partition plan is shared.
// fn[0] / fn[1]: per-genome S64 / S32 tables, fn[2] / fn[3]: group tables with 64-bit slots, whole windows / windows in pieces,
// fn[4] / fn[5]: group tables with 32-bit slots (G32)
struct K3WinVariant { int threads, cap; void (*fn[6])(ProbeArgs, uint32_t, uint32_t, uint32_t); };
#define K3W(T, IPT, MINB) {T, T * IPT, {probe_win_kernel<T, IPT, PK_FMT_S64, MINB, 0>, probe_win_kernel<T, IPT, PK_FMT_S32, MINB, 0>, \
                                        probe_win_kernel<T, IPT, PK_FMT_GROUP, MINB, 0>, probe_win_kernel<T, IPT, PK_FMT_GROUP, MINB, 1>, \
                                        probe_win_kernel<T, IPT, PK_FMT_GROUP32, MINB, 0>, probe_win_kernel<T, IPT, PK_FMT_GROUP32, MINB, 1>}}
static const K3WinVariant k3w_variants[] = {
    K3W(256, 3, 4),      // 0: <= 64 registers
    K3W(256, 3, 5),      // 1: <= 48 registers
    K3W(256, 3, 3),      // 2: <= 80 registers
    K3W(256, 3, 6),      // 3: <= 40 registers
    // large blocks for one-byte rows out of 32-bit-slot group tables (k3w_big; profiles/r2j..r2q_sweep_*.json: 128-thread
    // blocks, 8 blocks/SM of 256 threads and capacities 1280 / 1344 were all slower)
    K3W(512, 3, 2),      // 4: capacity 1536 (2^17 partitions on configs[1]; pairs with k3_variant 6)
    K3W(512, 3, 3),      // 5
    K3W(384, 4, 4),      // 6: capacity 1536 with 4 items per thread — the default (K3 1.39 vs 1.51 ms for 5)
};
#define PW_MAX_GROUP_STAGE_BYTES 32768u
#define PW_GROUP_STAGE_TARGET 24576u          // pieces of a group-table window are at most this large
#define PW_LEAN_MAX_STAGE 65536u              // the lean kernel's single stage: a whole window of up to this size
int pk_part_n_wvariants(void) { return (int)(sizeof k3w_variants / sizeof k3w_variants[0]); }
// auto: per-genome tables 6 blocks/SM (profiles/r1e_sweep.json: 5.14 vs 5.32 ms), group tables 4 blocks/SM with up to
// 64 registers (profiles/r1l_sweep.json: 2.29 vs 3.62 ms — one 20 KB window per block, the probe loop is short and
// the item registers matter more than occupancy)
// ... group tables with 32-bit slots (10 KB windows): 6 blocks/SM again (profiles/r2d_sweep_k3_variants_g32.json: 1.84 vs 1.95 ms)
static const K3WinVariant &k3w_pick(const PkPartTune &tu, uint32_t n_genomes_in_launch, bool group_tables = false, bool g32 = false) {
    (void)n_genomes_in_launch;
    return k3w_variants[tu.wvariant >= 0 && tu.wvariant < pk_part_n_wvariants() ? tu.wvariant : (group_tables && !g32 ? 0 : 3)];
}
// bytes one stage must hold for every table of the launch, or 0 when some window does not fit a stage
static uint32_t k3w_stage_bytes(const PkTable *tabs, uint32_t ng, uint32_t pb, uint32_t limit) {
    uint64_t mx = 0;
    for (uint32_t g = 0; g < ng; g++) {
        const uint64_t nbk = ((uint64_t)tabs[g].n_buckets >> pb) + 2;       // b1 - b0 + 1 <= ceil(nb / 2^pb) + 1
        mx = nbk > mx ? nbk : mx;
    }
    const uint64_t bytes = (mx * 32 + 127) & ~127ull;
    return bytes <= limit ? (uint32_t)bytes : 0;
}

// K4: scatter the (pos, bits) lists into rows. All blocks of one bin write inside a slice of
// 2^out_shift rows, which stays in L2 until its sectors are complete.
#define UP_TILE 4096
__global__ void __launch_bounds__(256) unpermute_kernel(const uint2 *__restrict__ list, const uint32_t *__restrict__ cursor,
                                                        uint32_t out_shift, uint32_t bin0, uint8_t *__restrict__ rows,
                                                        uint32_t row_stride, uint32_t col_offset, uint32_t nbl) {
    const uint32_t bin = bin0 + blockIdx.y, grp = blockIdx.z;
    const uint32_t cnt = cursor[(grp * PP_OBINS + bin) * PP_OCS];
    const uint32_t t0 = blockIdx.x * UP_TILE;
    if (t0 >= cnt) return;
    const uint2 *src = list + (((uint64_t)grp * PP_OBINS + bin) << out_shift);
    const uint32_t nb = min(4u, nbl - 4 * grp);
    const bool al4 = nb == 4 && ((row_stride | col_offset) & 3) == 0;
    const uint32_t t1 = min(cnt, t0 + UP_TILE);
    for (uint32_t i = t0 + threadIdx.x; i < t1; i += 256) {
        const uint2 e = src[i];
        uint8_t *dst = rows + (uint64_t)e.x * row_stride + col_offset + 4 * grp;
        if (al4) *(uint32_t *)dst = e.y;
        else for (uint32_t qb = 0; qb < nb; qb++) dst[qb] = (uint8_t)(e.y >> (8 * qb));
    }
}

// K4, one-byte rows: the 4-byte (position in bin, bits) lists -> rows through SHARED-MEMORY SLICES of the bitmap. A block
// owns 2^PP_FSLICE_SHIFT consecutive rows (128 KB of shared memory, zero-filled: rows of invalid windows are zero),
// scans its bin's whole list (L2-resident: the 2^(out_shift - 17) blocks of a bin run side by side), keeps the items of
// its slice with byte stores into shared memory, and writes the slice out with 16-byte stores. The scattered byte
// stores — 135 M of them on configs[1], each its own 32-byte sector through L1TEX in unpermute_kernel (1.05 ms) — stay
// on chip; the price is reading a bin's list once per slice.
__global__ void __launch_bounds__(1024, 1) unpermute_slice_kernel(const uint32_t *__restrict__ list, const uint32_t *__restrict__ cursor,
                                                                  uint32_t out_shift, uint32_t bin0, uint64_t n_rows, uint8_t *__restrict__ rows) {
    extern __shared__ __align__(16) uint8_t s_slice[];          // [1 << PP_FSLICE_SHIFT]
    constexpr uint32_t S = 1u << PP_FSLICE_SHIFT;
    const uint32_t sub_shift = out_shift > PP_FSLICE_SHIFT ? out_shift - PP_FSLICE_SHIFT : 0;
    const uint32_t bin = bin0 + (blockIdx.x >> sub_shift), sub = blockIdx.x & ((1u << sub_shift) - 1);
    const uint32_t slice = out_shift > PP_FSLICE_SHIFT ? S : 1u << out_shift;       // rows of this block
    const uint64_t row0 = ((uint64_t)bin << out_shift) + (uint64_t)sub * S;
    if (row0 >= n_rows) return;
    for (uint32_t i = threadIdx.x * 16; i < slice; i += blockDim.x * 16) *(uint4 *)(s_slice + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t cnt = min(cursor[bin * PP_OCS], 1u << out_shift);
    const uint32_t *src = list + ((uint64_t)bin << out_shift);
    const uint32_t n4 = cnt & ~3u;
    // four 16-byte loads in flight per thread: one block per SM has to cover the L2 latency on its own
    for (uint32_t i0 = threadIdx.x * 4; i0 < n4; i0 += blockDim.x * 16) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t i = i0 + u * blockDim.x * 4;
            v[u] = i < n4 ? *(const uint4 *)(src + i) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t it[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t p = it[q] >> 8;          // 0xffffff (padding): never inside a bin of <= 2^24 positions' slice range
                if ((p >> PP_FSLICE_SHIFT) == sub && i0 + u * blockDim.x * 4 < n4) s_slice[p & (S - 1)] = (uint8_t)it[q];
            }
        }
    }
    for (uint32_t i = n4 + threadIdx.x; i < cnt; i += blockDim.x) {
        const uint32_t v = src[i], p = v >> 8;
        if ((p >> PP_FSLICE_SHIFT) == sub) s_slice[p & (S - 1)] = (uint8_t)v;
    }
    __syncthreads();
    const uint32_t nout = (uint32_t)min((uint64_t)slice, n_rows - row0);
    uint8_t *dst = rows + row0;
    if ((((uintptr_t)dst) & 15) == 0) {
        for (uint32_t i = threadIdx.x * 16; i + 16 <= nout; i += blockDim.x * 16) *(uint4 *)(dst + i) = *(const uint4 *)(s_slice + i);
        for (uint32_t i = (nout & ~15u) + threadIdx.x; i < nout; i += blockDim.x) dst[i] = s_slice[i];
    } else {
        for (uint32_t i = threadIdx.x; i < nout; i += blockDim.x) dst[i] = s_slice[i];
    }
}

// ------------------------------------------------------------------ host orchestration
static uint32_t ceil_log2(uint64_t v) { uint32_t b = 0; while ((1ull << b) < v) b++; return b; }

uint32_t pk_part_max_fine_bins(void) { return PP_FBINS; }
void pk_part_plan(uint64_t n, const PkPartTune &tune, PkPartPlan *pl, uint32_t n_local, int fine_out, uint32_t k) {
    const uint32_t eb_g32 = 2 * k > 20 ? 2 * k - 20 : 0;
    // mean fill <= 5/6 of the K3 block capacity (>= 20% head-room for the Poisson spread). One-byte rows out of 32-bit-slot
    // group tables (fine_out == 2): 512-thread blocks of capacity 1536 (profiles/r2j_sweep.json: K3 2.01 vs 2.19 ms)
    const bool big = fine_out == 2 && tune.fine_out && n_local <= 8 && tune.variant < 0 && tune.wvariant < 0;
    const int wbig = tune.wbig >= 4 && tune.wbig < pk_part_n_wvariants() ? tune.wbig : 6;
    const uint32_t cap = big ? (uint32_t)k3w_variants[wbig].cap : (uint32_t)k3_pick(tune, 1).cap;
    pl->wbig = big ? wbig : -1;
    pl->compact = 0;
    const uint64_t fill = (uint64_t)cap * 5 / 6;
    uint32_t pb = ceil_log2((n + fill - 1) / fill);
    if (pb > 18) pb = 18;
    if (pb < 1) pb = 1;
    pl->pb1 = pb > 9 ? 9 : pb;
    pl->pb2 = pb - pl->pb1;
    pl->cap2 = cap;
    pl->cap1 = pl->pb2 ? cap << pl->pb2 : cap;
    pl->n_regions1 = 1u << pl->pb1;
    pl->n_regions2 = pl->pb2 ? 1u << pb : 0;
    pl->buf1_items = (uint64_t)pl->n_regions1 * pl->cap1;
    pl->buf2_items = (uint64_t)pl->n_regions2 * pl->cap2;
    pl->spill_items = n;
    // un-permute lists: <= PP_OBINS bins of 2^out_shift positions, one set per group of 32 genomes
    pl->out_shift = ceil_log2((n + PP_OBINS - 1) / PP_OBINS);
    if (pl->out_shift < 8) pl->out_shift = 8;
    pl->out_bins = (uint32_t)((n + (1ull << pl->out_shift) - 1) >> pl->out_shift);
    pl->out_fine = 0;
    pl->fine_rows = n;
    pl->out_bytes = ((uint64_t)((n_local + 31) / 32) * PP_OBINS << pl->out_shift) * 8;
    if (fine_out && tune.fine_out && n_local <= 8) {
        // one-byte rows: <= 512 bins of >= 2^17 positions (a position in its bin + 8 bits fit 32 bits up to bins of 2^24)
        uint32_t sh = ceil_log2((n + PP_FBINS - 1) / PP_FBINS);
        if (tune.fine_shift > 0 && (uint32_t)tune.fine_shift > sh) sh = (uint32_t)tune.fine_shift;     // fewer, larger bins
        if (sh < PP_FSLICE_SHIFT) sh = PP_FSLICE_SHIFT;
        if (sh <= 24) {
            pl->out_fine = 1;
            pl->out_shift = sh;
            pl->out_bins = (uint32_t)((n + (1ull << sh) - 1) >> sh);
            pl->out_bytes = ((uint64_t)pl->out_bins << sh) * 4;
            // compact items: see PartSmem. 9 coarse bits, the fine digit inside the eb - 9 stored hash bits, 28-bit positions
            pl->compact = (big && tune.compact && pl->pb1 == 9 && eb_g32 >= 18 && eb_g32 <= 25 && n <= (1ull << PT_CI_POS_BITS)) ? 1 : 0;
        }
    }
}
uint32_t pk_part_obins(void) { return PP_OBINS; }
uint32_t pk_part_ocursor_words(void) { return (PP_OBINS > PP_FBINS ? PP_OBINS : PP_FBINS) * PP_OCS; }

// The stages of one partitioned batch. A batch may be fed in pieces (pk_part_append per chromosome, as
// its bytes arrive) and drained in pieces (pk_part_unpermute per range of position bins).
static PartArgs make_part_args(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, PkKeySpec ks, uint32_t n_local,
                               uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl,
                               const PkPartScratch &sc) {
    PartArgs a{};
    a.words = d_words; a.mask64 = (const uint64_t *)d_mask; a.p0 = p0; a.ks = ks;
    a.pb1 = pl.pb1; a.pb2 = pl.pb2; a.cap1 = pl.cap1; a.cap2 = pl.cap2;
    a.buf1 = (uint2 *)sc.buf1; a.buf2 = (uint2 *)sc.buf2; a.spill = (uint2 *)sc.spill;
    a.cursor1 = sc.cursor1; a.cursor2 = sc.cursor2; a.spill_cursor = sc.spill_cursor; a.spill_cap = pl.spill_items;
    a.err = sc.err;
    a.rows = d_rows; a.row_stride = row_stride; a.col_offset = col_offset; a.nbl = (n_local + 7) / 8;
    a.compact = pl.compact && ks.ghash; a.eb = 2 * ks.k > 20 ? 2 * ks.k - 20 : 0;
    return a;
}

void pk_part_begin(uint32_t n_local, const PkPartPlan &pl, const PkPartScratch &sc, pk_stream_t s) {
    cudaMemsetAsync(sc.cursor1, 0, sizeof(uint32_t) * pl.n_regions1, s);
    if (pl.n_regions2) cudaMemsetAsync(sc.cursor2, 0, sizeof(uint32_t) * pl.n_regions2, s);
    cudaMemsetAsync(sc.spill_cursor, 0, sizeof(unsigned long long), s);
    if (sc.out_list) cudaMemsetAsync(sc.out_cursor, 0, sizeof(uint32_t) * (pl.out_fine ? PP_FBINS : ((n_local + 31) / 32) * PP_OBINS) * PP_OCS, s);
}

// K1 over positions [p0 + off, p0 + off + n) of the batch that starts at p0 (pos = off + i)
void pk_part_append(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t off, uint64_t n, PkKeySpec ks,
                    uint32_t n_local, uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl,
                    const PkPartScratch &sc, pk_stream_t s) {
    if (!n) return;
    PartArgs a = make_part_args(d_words, d_mask, p0, ks, n_local, d_rows, row_stride, col_offset, pl, sc);
    a.off = off; a.n = n;
    const unsigned grid = (unsigned)((n + PT_TILE - 1) / PT_TILE);
    if (a.compact || (sc.tune && sc.tune->k1_roll)) {        // compact items: only the rolling K1 emits them
        if (a.compact) partition_seq_roll_kernel<4, 1><<<grid, PT_THREADS, 0, s>>>(a);
        else partition_seq_roll_kernel<4, 0><<<grid, PT_THREADS, 0, s>>>(a);
    }
    else partition_seq_kernel<<<grid, PT_THREADS, 0, s>>>(a);
}

// K2 + K3 (+ spill drain) over everything appended so far
void pk_part_probe(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, PkKeySpec ks, const PkTable *h_tables,
                   const PkTable *h_utables, const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl,
                   const PkPartScratch &sc, int prefetch, pk_stream_t s, cudaEvent_t *evs) {
    const int fi = ks.fmt == PK_FMT_S32 ? 1 : 0;
    const uint32_t n_groups = (n_local + 31) / 32;
    const PkPartTune &tu = *sc.tune;
    int lw_dummy = 0;
    int &last_window = sc.last_window ? *sc.last_window : lw_dummy;
    PartArgs a = make_part_args(d_words, d_mask, p0, ks, n_local, d_rows, row_stride, col_offset, pl, sc);
    // tune "k3_l2" = variant + 1 (0: off): no K2, K3 probes the coarse regions through L2
    const bool l2_path = tu.k3_l2 > 0 && pl.pb2 && a.compact && pl.out_fine && sc.out_list && h_utables && n_local <= PK_U_GROUP &&
                         h_utables[0].fmt == PK_TFMT_G32 && tu.rank_atomic == 1;
    if (pl.pb2 && !l2_path) {
        dim3 grid((pl.cap1 + PT_TILE - 1) / PT_TILE, pl.n_regions1);
        partition_fine_kernel<<<grid, PT_THREADS, 0, s>>>(a);
    }
    ProbeArgs p{};
    p.buf = pl.pb2 ? (const uint2 *)sc.buf2 : (const uint2 *)sc.buf1;
    p.counts = pl.pb2 ? sc.cursor2 : sc.cursor1;
    p.cap = pl.pb2 ? pl.cap2 : pl.cap1;
    p.n_regions = pl.pb2 ? pl.n_regions2 : pl.n_regions1;
    p.pb = pl.pb1 + pl.pb2;
    p.words = d_words; p.p0 = p0; p.ks = ks;
    p.rows = d_rows; p.row_stride = row_stride; p.col_offset = col_offset; p.nbl = a.nbl;
    p.prefetch = prefetch;
    p.out_list = (uint2 *)sc.out_list;
    p.out_cursor = sc.out_cursor;
    p.out_shift = pl.out_shift;
    p.n_bins = pl.out_bins;
    p.compact = a.compact; p.eb = a.eb; p.pb2 = pl.pb2;
    p.rank_atomic = (uint32_t)tu.rank_atomic;
    p.out_fine = 0;
    if (evs) cudaEventRecord(evs[2], s);
    bool regions_done = false;         // the regions have been answered for ALL genomes by the L1/L2 group kernel
    if (l2_path) {
        ProbeArgs q = p;
        q.buf = (const uint2 *)sc.buf1; q.counts = sc.cursor1; q.cap = pl.cap1; q.n_regions = pl.n_regions1; q.pb = pl.pb1; q.pb2 = 0;
        q.grp = 0; q.g_first = 0; q.n_genomes = n_local; q.ng = 1; q.tbits = PK_U_GROUP; q.out_fine = pl.out_fine;
        q.tabs[0] = h_utables[0];
        const K3L2Variant &gv = k3g_variants[tu.k3_l2 <= pk_part_n_gvariants() ? tu.k3_l2 - 1 : 0];
        dim3 grid((pl.cap1 + gv.cap - 1) / gv.cap, pl.n_regions1);
        gv.fn<<<grid, gv.threads, 0, s>>>(q);
        last_window = 5;
        regions_done = true;
    }
    for (uint32_t grp = 0; grp < n_groups && !regions_done; grp++) {       // one launch per group of 32 genomes
        const uint32_t ngen = n_local - 32 * grp < 32 ? n_local - 32 * grp : 32;
        p.grp = grp; p.g_first = 32 * grp; p.n_genomes = ngen;
        last_window = 0;
        const K3WinVariant &wv_group = (tu.wvariant < 0 && pl.wbig >= 0) ? k3w_variants[pl.wbig]
                                                                            : k3w_pick(tu, ngen, true, h_utables && h_utables[4 * grp].fmt == PK_TFMT_G32);
        if (tu.window && h_utables && (uint32_t)wv_group.cap == pl.cap2) {
            // group tables: one probe per 8 genomes; the launch walks the (<= 4) group tables of its 32 genomes. A
            // window larger than PW_GROUP_STAGE_TARGET is cut into equal pieces that pass through the stages one
            // after the other (every item probes the piece its home bucket lies in).
            const uint32_t nt = (ngen + PK_U_GROUP - 1) / PK_U_GROUP;
            uint64_t maxw = 0, win[4];
            for (uint32_t u = 0; u < nt; u++) {
                p.tabs[u] = h_utables[4 * grp + u];
                win[u] = ((uint64_t)p.tabs[u].n_buckets >> p.pb) + 2;          // buckets, upper bound
                maxw = std::max(maxw, win[u]);
            }
            const uint64_t target = PW_GROUP_STAGE_TARGET / 32;
            const uint64_t nch = (maxw + target - 1) / target;
            const uint64_t cb = ((maxw + nch - 1) / nch + 3) & ~3ull;          // buckets per piece, a multiple of 128 bytes
            uint32_t np = 0;
            bool ok = cb * 32 <= PW_MAX_GROUP_STAGE_BYTES;
            for (uint32_t u = 0; u < nt && ok; u++)
                for (uint64_t c = 0; c * cb < win[u]; c++) {
                    if (np == 32) { ok = false; break; }
                    p.t_of[np] = (uint8_t)u; p.c_of[np] = (uint8_t)c; np++;
                }
            if (ok && np) {
                p.out_fine = pl.out_fine;
                const K3WinVariant &wv = wv_group;
                const uint32_t stage_bytes = (uint32_t)cb * 32, n_stages = np == 1 ? 1 : 2;
                p.ng = np; p.tbits = PK_U_GROUP; p.chunk_buckets = nch > 1 ? (uint32_t)cb : 0;
                const int fk = (nch > 1 ? 3 : 2) + (p.tabs[0].fmt == PK_TFMT_G32 ? 2 : 0);
                // ONE 32-bit-slot table, compact items, fine bins: the lean kernel, which stages a partition's whole
                // window at once — up to PW_LEAN_MAX_STAGE bytes (a half-genome batch of the end-to-end call has 2^16
                // partitions and 40 KB windows: 4-5 blocks/SM still fit), where the general kernel would cut it in pieces
                const uint64_t lean_stage = (maxw * 32 + 127) & ~127ull;
                if (p.tabs[0].fmt == PK_TFMT_G32 && nt == 1 && lean_stage <= PW_LEAN_MAX_STAGE && p.compact && p.out_fine && sc.out_list &&
                    p.counts && tu.lean > 0 && tu.lean <= pk_part_n_lvariants() && (uint32_t)k3l_variants[tu.lean - 1].cap == pl.cap2 &&
                    tu.rank_atomic == 1) {
                    const K3LeanVariant &lv = k3l_variants[tu.lean - 1];
                    cudaFuncSetAttribute(lv.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PW_LEAN_MAX_STAGE);
                    lv.fn<<<p.n_regions, lv.threads, (size_t)lean_stage, s>>>(p, (uint32_t)lean_stage);
                    last_window = 4;
                    continue;
                }
                (void)fk;
                cudaFuncSetAttribute(wv.fn[fk], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PW_MAX_STAGES * PW_MAX_STAGE_BYTES));
                wv.fn[fk]<<<p.n_regions, wv.threads, (size_t)n_stages * stage_bytes, s>>>(p, stage_bytes, 1, n_stages);
                last_window = 2;
                continue;
            }
        }
        if (h_utables && d_utables) {
            // group tables, but this launch's windows cannot be staged (more than 32 pieces: a short batch against
            // large tables) or the window kernels are switched off: probe the regions through L1/L2, all genomes at once
            const bool fine = pl.out_fine && sc.out_list;
            pk_launch_items_group(p.buf, p.counts, nullptr, p.n_regions, p.cap, d_words, p0, ks, d_utables, n_local, d_rows, row_stride, col_offset,
                                  fine ? (uint32_t *)sc.out_list : nullptr, sc.out_cursor, PP_OCS, pl.out_shift, (int)a.compact, s);
            last_window = 3;
            regions_done = true;
            continue;
        }
        if (pl.out_fine) { last_window = -1; regions_done = true; continue; }     // unreachable: a fine plan needs group tables (handled above)
        p.chunk_buckets = 0;
        for (uint32_t g = 0; g < 32; g++) { p.t_of[g] = (uint8_t)g; p.c_of[g] = 0; }
        p.ng = ngen; p.tbits = 1;
        for (uint32_t g = 0; g < p.ng; g++) p.tabs[g] = h_tables[32 * grp + g];
        const uint32_t stage_bytes = tu.window && k3_pick(tu, p.ng).cap == k3w_pick(tu, p.ng).cap ? k3w_stage_bytes(p.tabs, p.ng, p.pb, PW_MAX_STAGE_BYTES) : 0;
        if (stage_bytes) {
            const K3WinVariant &wv = k3w_pick(tu, p.ng);
            // two genomes per group while four windows stay under ~24 KB (6 blocks/SM), else one: measured on
            // configs[1] (4.1 KB windows: 5.14 vs 5.81 ms) and on k=31 tables (9.2 KB windows: 9.42 vs 7.86 ms)
            const uint32_t gsz = tu.wgroup ? tu.wgroup : (4 * stage_bytes <= 24576 ? 2 : 1);
            const size_t dyn = (size_t)2 * gsz * stage_bytes;
            cudaFuncSetAttribute(wv.fn[fi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PW_MAX_STAGES * PW_MAX_STAGE_BYTES));
            wv.fn[fi]<<<p.n_regions, wv.threads, dyn, s>>>(p, stage_bytes, gsz, 2 * gsz);
            last_window = 1;
        } else {
            const K3Variant &kv = k3_pick(tu, p.ng);
            kv.fn[fi]<<<p.n_regions, kv.threads, 0, s>>>(p);
        }
    }
    if (evs) cudaEventRecord(evs[3], s);
    if (h_utables && d_utables) {      // drain the spill list (normally empty) out of the group tables
        const bool fine = pl.out_fine && sc.out_list;
        pk_launch_items_group(sc.spill, nullptr, sc.spill_cursor, 0, pl.spill_items, d_words, p0, ks, d_utables, n_local, d_rows, row_stride, col_offset,
                              fine ? (uint32_t *)sc.out_list : nullptr, sc.out_cursor, PP_OCS, pl.out_shift, (int)a.compact, s);
        if (evs) cudaEventRecord(evs[4], s);
        return;
    }
    ProbeArgs sp = p;          // drain the spill list (normally empty: the blocks exit at once)
    sp.buf = (const uint2 *)sc.spill; sp.counts = nullptr; sp.flat_total = sc.spill_cursor; sp.cap = k3_pick(tu, 1).cap; sp.pb = 0;
    for (uint32_t grp = 0; grp < n_groups; grp++) {
        sp.grp = grp; sp.ng = n_local - 32 * grp < 32 ? n_local - 32 * grp : 32;
        sp.tbits = 1; sp.g_first = 32 * grp; sp.n_genomes = sp.ng;
        for (uint32_t g = 0; g < sp.ng; g++) sp.tabs[g] = h_tables[32 * grp + g];
        const K3Variant &kv = k3_pick(tu, sp.ng);
        kv.fn[fi]<<<148 * 2, kv.threads, 0, s>>>(sp);
    }
    if (evs) cudaEventRecord(evs[4], s);
}

// K4 for position bins [bin0, bin1)
void pk_part_unpermute(uint32_t bin0, uint32_t bin1, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride,
                       uint32_t col_offset, const PkPartPlan &pl, const PkPartScratch &sc, pk_stream_t s) {
    if (!sc.out_list || bin1 <= bin0) return;
    if (pl.out_fine) {
        const uint32_t sub_shift = pl.out_shift > PP_FSLICE_SHIFT ? pl.out_shift - PP_FSLICE_SHIFT : 0;
        const size_t smem = (size_t)1 << (pl.out_shift > PP_FSLICE_SHIFT ? PP_FSLICE_SHIFT : pl.out_shift);
        static bool attr_set = false;
        if (!attr_set) { cudaFuncSetAttribute(unpermute_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 << PP_FSLICE_SHIFT); attr_set = true; }
        unpermute_slice_kernel<<<(bin1 - bin0) << sub_shift, 1024, smem, s>>>((const uint32_t *)sc.out_list, sc.out_cursor, pl.out_shift, bin0,
                                                                              pl.fine_rows, d_rows + col_offset);
        return;
    }
    dim3 grid((unsigned)(((1ull << pl.out_shift) + UP_TILE - 1) / UP_TILE), bin1 - bin0, (n_local + 31) / 32);
    unpermute_kernel<<<grid, 256, 0, s>>>((const uint2 *)sc.out_list, sc.out_cursor, pl.out_shift, bin0, d_rows, row_stride,
                                          col_offset, (n_local + 7) / 8);
}

int pk_launch_probe_partitioned(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, PkKeySpec ks,
                                const PkTable *d_tables, const PkTable *h_tables, const PkTable *h_utables, const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows,
                                uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl, const PkPartScratch &sc,
                                int prefetch, pk_stream_t s, cudaEvent_t *evs) {
    if (!n) return 0;
    (void)d_tables;
    pk_part_begin(n_local, pl, sc, s);
    if (evs) cudaEventRecord(evs[0], s);
    pk_part_append(d_words, d_mask, p0, 0, n, ks, n_local, d_rows, row_stride, col_offset, pl, sc, s);
    if (evs) cudaEventRecord(evs[1], s);
    pk_part_probe(d_words, d_mask, p0, ks, h_tables, h_utables, d_utables, n_local, d_rows, row_stride, col_offset, pl, sc, prefetch, s, evs);
    pk_part_unpermute(0, pl.out_bins, n_local, d_rows, row_stride, col_offset, pl, sc, s);
    if (evs) cudaEventRecord(evs[5], s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
