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

#include <cstdlib>

#include "pk_device.cuh"
#include "pk_internal.h"

#define PT_THREADS 256
#define PT_IPT 16
#define PT_TILE (PT_THREADS * PT_IPT)
#define PT_WARPS (PT_THREADS / 32)
#define PT_MAXB 512

struct PartSmem {
    uint16_t warp_cnt[PT_WARPS][PT_MAXB];   // per-warp digit counts, then exclusive prefix over warps
    uint32_t tot[PT_MAXB];                  // per-digit tile totals
    uint32_t dstart[PT_MAXB];               // exclusive scan of tot: start of the digit's run in `stage`
    uint32_t gbase[PT_MAXB];                // reserved offset inside the destination region
    uint32_t warp_sums[PT_WARPS];
    uint32_t tile_total;
    uint2 stage[PT_TILE];
};

// Scatter a tile of items (h, pos) into fixed-capacity regions by digit = (h >> shift) & (nb-1).
// Region r = region0 + digit holds items dst[r * cap .. r * cap + min(cursor[r], cap)).
template <int RANK>
__device__ __forceinline__ void block_partition(PartSmem &S, const uint32_t (&h)[PT_IPT], const uint32_t (&pos)[PT_IPT],
                                                uint32_t validmask, uint32_t shift, uint32_t nb, uint2 *__restrict__ dst,
                                                uint64_t region0, uint32_t cap, uint32_t *__restrict__ cursor,
                                                uint2 *__restrict__ spill, unsigned long long *__restrict__ spill_cursor,
                                                uint64_t spill_cap, uint32_t *__restrict__ err) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t lt = (1u << lane) - 1;
    uint16_t rank[PT_IPT];
    if (RANK == 1) {
        // block-wide counters: one shared-memory atomic per item returns its rank inside (tile, digit)
        for (uint32_t d = tid; d < nb; d += PT_THREADS) S.tot[d] = 0;
        for (uint32_t i = tid; i < PT_WARPS * PT_MAXB; i += PT_THREADS) (&S.warp_cnt[0][0])[i] = 0;   // prefix over warps = 0
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PT_IPT; j++)
            rank[j] = (validmask >> j) & 1 ? (uint16_t)atomicAdd(&S.tot[(h[j] >> shift) & (nb - 1)], 1u) : 0;
        __syncthreads();
    } else {
    for (uint32_t i = tid; i < PT_WARPS * PT_MAXB; i += PT_THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t peers[PT_IPT];
    // all MATCH.ANY first (independent, long latency), then the serial per-warp counter updates
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        const bool v = (validmask >> j) & 1;
        peers[j] = __match_any_sync(0xffffffffu, v ? (h[j] >> shift) & (nb - 1) : 0xffffffffu);
    }
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        const bool v = (validmask >> j) & 1;
        const uint32_t d = (h[j] >> shift) & (nb - 1);
        const uint32_t leader = __ffs(peers[j]) - 1;
        uint32_t old = 0;
        if (v && lane == leader) {
            old = S.warp_cnt[w][d];
            S.warp_cnt[w][d] = (uint16_t)(old + __popc(peers[j]));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = (uint16_t)(old + __popc(peers[j] & lt));
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t d = tid; d < nb; d += PT_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < PT_WARPS; ww++) {
            const uint32_t t = S.warp_cnt[ww][d];
            S.warp_cnt[ww][d] = (uint16_t)run;
            run += t;
        }
        S.tot[d] = run;
    }
    __syncthreads();
    }
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
    for (uint32_t d = tid; d < nb; d += PT_THREADS)
        if (S.tot[d]) S.gbase[d] = atomicAdd(&cursor[region0 + d], S.tot[d]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        if ((validmask >> j) & 1) {
            const uint32_t d = (h[j] >> shift) & (nb - 1);
            S.stage[S.dstart[d] + S.warp_cnt[w][d] + rank[j]] = make_uint2(h[j], pos[j]);
        }
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

struct PartArgs {
    const uint64_t *words;
    const uint64_t *mask64;
    uint64_t p0, n;
    uint32_t k;
    uint32_t pb1, pb2;              // coarse / fine bits
    uint32_t cap1, cap2;            // region capacities
    uint2 *buf1, *buf2, *spill;
    uint32_t *cursor1, *cursor2;
    unsigned long long *spill_cursor;
    uint64_t spill_cap;
    uint32_t *err;
    uint8_t *rows;
    uint32_t row_stride, col_offset, nbl;
};

// K1: positions [p0, p0+n) -> (hash, i) pairs (i = position - p0), partitioned by the top pb1 bits.
// Invalid windows get their all-zero row here and never enter the pipeline.
template <int RANK>
__global__ void __launch_bounds__(PT_THREADS, 3) partition_seq_kernel(PartArgs a) {
    __shared__ PartSmem S;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t base = blockIdx.x * (uint64_t)PT_TILE + (uint64_t)w * (32 * PT_IPT) + lane;
    uint32_t h[PT_IPT], pos[PT_IPT], valid = 0;
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        const uint64_t i = base + 32 * j;
        h[j] = 0; pos[j] = (uint32_t)i;
        if (i < a.n) {
            uint64_t canon;
            if (pk_window(a.words, a.mask64, a.p0 + i, a.k, canon)) {
                h[j] = pk_hash32(canon);
                valid |= 1u << j;
            } else {
                uint8_t *dst = a.rows + i * a.row_stride + a.col_offset;
                for (uint32_t q = 0; q < a.nbl; q++) dst[q] = 0;
            }
        }
    }
    block_partition<RANK>(S, h, pos, valid, 32 - a.pb1, 1u << a.pb1, a.buf1, 0, a.cap1, a.cursor1, a.spill, a.spill_cursor,
                          a.spill_cap, a.err);
}

// K2: coarse region c = blockIdx.y, tile blockIdx.x of it -> fine regions c * 2^pb2 + next pb2 bits.
template <int RANK>
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
    block_partition<RANK>(S, h, pos, valid, 32 - a.pb1 - a.pb2, 1u << a.pb2, a.buf2, (uint64_t)c << a.pb2, a.cap2, a.cursor2,
                          a.spill, a.spill_cursor, a.spill_cap, a.err);
}

// ------------------------------------------------------------------ K3: probe one partition per block
#define PP_THREADS 256
#define PP_CAP 3072                       // items per fine partition (shared memory: 16 B each)
#define PP_IPT (PP_CAP / PP_THREADS)      // 12
#define PP_ILP 4
#define PP_OBINS 128                      // position bins of the un-permute lists
#define PP_WARPS (PP_THREADS / 32)

__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct ProbeArgs {
    const uint2 *buf;                 // regions of `cap` items
    const uint32_t *counts;           // per-region fill (may exceed cap: clamp); NULL => flat list of *flat_total items
    const unsigned long long *flat_total;
    uint32_t cap;
    uint32_t n_regions;               // regions mode
    uint32_t pb;                      // total partition bits (regions mode; 0 disables prefetch)
    const uint64_t *words;
    uint64_t p0;
    uint32_t k;
    const PkTable *tables;
    uint32_t n_local;
    uint8_t *rows;
    uint32_t row_stride, col_offset, nbl;
    int prefetch;
    // un-permute: instead of scattering row bytes over the whole bitmap (32x write amplification and a
    // read-modify-write per byte in DRAM), results are appended as (pos, bits) to lists binned by
    // pos >> out_shift; unpermute_kernel then scatters each list inside its own L2-resident slice
    uint2 *out_list;                  // [(grp * PP_OBINS + bin) << out_shift]
    uint32_t *out_cursor;             // [n_groups * PP_OBINS]
    uint32_t out_shift;
};

__device__ __noinline__ bool pk_probe_slow_ca(const PkTable t, unsigned long long key, uint32_t b) {
    for (uint32_t tries = 1; tries < t.n_buckets; ++tries) {
        b = b + 1 == t.n_buckets ? 0 : b + 1;
        const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * b);
        if (v.a == key || v.b == key || v.c == key || v.d == key) return true;
        if (v.a == PK_EMPTY || v.b == PK_EMPTY || v.c == PK_EMPTY || v.d == PK_EMPTY) return false;
    }
    return false;
}

__global__ void __launch_bounds__(PP_THREADS, 3) probe_part_kernel(ProbeArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *s_canon = (uint64_t *)smem_raw;                    // [PP_CAP]
    uint32_t *s_h = (uint32_t *)(s_canon + PP_CAP);              // [PP_CAP]
    uint32_t *s_pos = s_h + PP_CAP;                              // [PP_CAP]
    __shared__ uint16_t o_wc[PP_WARPS][PP_OBINS];
    __shared__ uint32_t o_gb[PP_OBINS];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint64_t nreg = a.n_regions;
    unsigned long long flat = 0;
    if (!a.counts) { flat = *a.flat_total; nreg = (flat + a.cap - 1) / a.cap; }
    for (uint64_t q = blockIdx.x; q < nreg; q += gridDim.x) {
        uint32_t cnt;
        if (a.counts) cnt = min(a.counts[q], a.cap);
        else cnt = (uint32_t)min((unsigned long long)a.cap, flat - q * a.cap);
        if (cnt == 0) continue;
        const uint2 *src = a.buf + q * (uint64_t)a.cap;
        for (uint32_t i = tid; i < cnt; i += PP_THREADS) {
            const uint2 it = src[i];
            const uint64_t p = a.p0 + it.y;
            const uint64_t w0 = a.words[p >> 5], w1 = a.words[(p >> 5) + 1];
            const uint32_t s = 2 * ((uint32_t)p & 31);
            const uint64_t x = (w0 << s) | ((w1 >> 1) >> (63 - s));
            const uint64_t fwd = x >> (64 - 2 * a.k);
            const uint64_t rc = pk_revcomp(fwd, a.k);
            s_canon[i] = fwd < rc ? fwd : rc;
            s_h[i] = it.x;
            s_pos[i] = it.y;
        }
        __syncthreads();
        const bool do_pf = a.prefetch && a.counts && a.pb;
        // hash range of this partition: [q << (32-pb), ((q+1) << (32-pb)) - 1]
        const uint32_t h_lo = do_pf ? (uint32_t)(q << (32 - a.pb)) : 0;
        const uint32_t h_hi = do_pf ? (uint32_t)(((q + 1) << (32 - a.pb)) - 1) : 0;
        if (do_pf && tid == 0) {
            const PkTable t = a.tables[0];
            const uint32_t b0 = __umulhi(h_lo, t.n_buckets), b1 = __umulhi(h_hi, t.n_buckets);
            if (b1 - b0 < 8192) l2_prefetch_bulk(t.slots + 4ull * b0, (b1 - b0 + 1) * 32);
        }
        for (uint32_t g0 = 0; g0 < a.n_local; g0 += 32) {
            uint32_t bits[PP_IPT];
#pragma unroll
            for (int j = 0; j < PP_IPT; j++) bits[j] = 0;
            const uint32_t ng = min(32u, a.n_local - g0);
            for (uint32_t gg = 0; gg < ng; gg++) {
                const PkTable t = a.tables[g0 + gg];
                if (do_pf && tid == 0 && g0 + gg + 1 < a.n_local) {
                    const PkTable tn = a.tables[g0 + gg + 1];
                    const uint32_t b0 = __umulhi(h_lo, tn.n_buckets), b1 = __umulhi(h_hi, tn.n_buckets);
                    if (b1 - b0 < 8192) l2_prefetch_bulk(tn.slots + 4ull * b0, (b1 - b0 + 1) * 32);   // windows <= 256 KB
                }
#pragma unroll
                for (int j0 = 0; j0 < PP_IPT; j0 += PP_ILP) {
                    u64x4 v[PP_ILP];
                    uint32_t b[PP_ILP];
#pragma unroll
                    for (int u = 0; u < PP_ILP; u++) {
                        const uint32_t i = tid + (j0 + u) * PP_THREADS;
                        if (i < cnt) {
                            b[u] = __umulhi(s_h[i], t.n_buckets);
                            v[u] = pk_ld_bucket_ca(t.slots + 4ull * b[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PP_ILP; u++) {
                        const uint32_t i = tid + (j0 + u) * PP_THREADS;
                        if (i < cnt) {
                            const uint64_t key = s_canon[i];
                            bool hit = v[u].a == key || v[u].b == key || v[u].c == key || v[u].d == key;
                            if (!hit && v[u].a != PK_EMPTY && v[u].b != PK_EMPTY && v[u].c != PK_EMPTY && v[u].d != PK_EMPTY)
                                hit = pk_probe_slow_ca(t, key, b[u]);
                            bits[j0 + u] |= (uint32_t)hit << gg;
                        }
                    }
                }
            }
            if (a.out_list) {
                // append (pos, bits) to the list of bin = pos >> out_shift: rank inside the block by
                // warp match, one atomicAdd per (block, bin) reserves the run
                for (uint32_t q = tid; q < PP_WARPS * PP_OBINS; q += PP_THREADS) (&o_wc[0][0])[q] = 0;
                __syncthreads();
                uint16_t rank[PP_IPT];
#pragma unroll
                for (int j = 0; j < PP_IPT; j++) {
                    const uint32_t i = tid + j * PP_THREADS;
                    const bool v = i < cnt;
                    const uint32_t bin = v ? s_pos[i] >> a.out_shift : 0xffffffffu;
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
                for (uint32_t b = tid; b < PP_OBINS; b += PP_THREADS) {
                    uint32_t run = 0;
#pragma unroll
                    for (int ww = 0; ww < PP_WARPS; ww++) {
                        const uint32_t t = o_wc[ww][b];
                        o_wc[ww][b] = (uint16_t)run;
                        run += t;
                    }
                    if (run) o_gb[b] = atomicAdd(&a.out_cursor[(g0 / 32) * PP_OBINS + b], run);
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < PP_IPT; j++) {
                    const uint32_t i = tid + j * PP_THREADS;
                    if (i < cnt) {
                        const uint32_t pos = s_pos[i], bin = pos >> a.out_shift;
                        const uint64_t slot = (((uint64_t)(g0 / 32) * PP_OBINS + bin) << a.out_shift) + o_gb[bin] + o_wc[wid][bin] + rank[j];
                        a.out_list[slot] = make_uint2(pos, bits[j]);
                    }
                }
                __syncthreads();
            } else {
                const uint32_t nb = min(4u, a.nbl - g0 / 8);
                const bool al4 = nb == 4 && ((a.row_stride | a.col_offset) & 3) == 0;
#pragma unroll
                for (int j = 0; j < PP_IPT; j++) {
                    const uint32_t i = tid + j * PP_THREADS;
                    if (i < cnt) {
                        uint8_t *dst = a.rows + (uint64_t)s_pos[i] * a.row_stride + a.col_offset + g0 / 8;
                        if (al4) *(uint32_t *)dst = bits[j];
                        else for (uint32_t qb = 0; qb < nb; qb++) dst[qb] = (uint8_t)(bits[j] >> (8 * qb));
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K3, item-major variant
// One item per thread at a time, GILP genomes in flight for it: a block's round trips to memory drop
// from (genomes x rounds) to (items-per-thread x genome-groups); the walk-on to the next bucket is
// batched over the GILP genomes instead of serialised. Partitions are smaller (<= T*IPT items) so that
// the GILP windows a block has live stay small (8 KB each at 2^18 partitions of a 2 GB table).
template <int T, int IPT, int GILP, int MINB>
__global__ void __launch_bounds__(T, MINB) probe_item_kernel(ProbeArgs a) {
    __shared__ uint16_t o_wc[T / 32][PP_OBINS];
    __shared__ uint32_t o_gb[PP_OBINS];
    __shared__ PkTable s_tb[32];
    constexpr int QCAP = T * IPT;
    __shared__ uint32_t s_bits[QCAP];              // results of the deferred walk-ons, by item
    __shared__ unsigned long long q_key[QCAP];     // deferred walk-on queue
    __shared__ uint32_t q_h[QCAP], q_meta[QCAP];
    __shared__ uint32_t q_n;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint64_t nreg = a.n_regions;
    unsigned long long flat = 0;
    if (!a.counts) { flat = *a.flat_total; nreg = (flat + a.cap - 1) / a.cap; }
    for (uint64_t q = blockIdx.x; q < nreg; q += gridDim.x) {
        uint32_t cnt;
        if (a.counts) cnt = min(a.counts[q], a.cap);
        else cnt = (uint32_t)min((unsigned long long)a.cap, flat - q * a.cap);
        if (cnt == 0) continue;
        const bool do_pf = a.prefetch && a.counts && a.pb;
        const uint32_t h_lo = do_pf ? (uint32_t)(q << (32 - a.pb)) : 0;
        const uint32_t h_hi = do_pf ? (uint32_t)(((q + 1) << (32 - a.pb)) - 1) : 0;
        if (do_pf && tid < min((uint32_t)GILP, a.n_local)) {
            const PkTable t = a.tables[tid];
            const uint32_t b0 = __umulhi(h_lo, t.n_buckets), b1 = __umulhi(h_hi, t.n_buckets);
            if (b1 - b0 < 8192) l2_prefetch_bulk(t.slots + 4ull * b0, (b1 - b0 + 1) * 32);
        }
        const uint2 *src = a.buf + q * (uint64_t)a.cap;
        uint64_t canon[IPT];
        uint32_t h[IPT], pos[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const uint32_t i = tid + j * T;
            canon[j] = 0; h[j] = 0; pos[j] = 0;
            if (i < cnt) {
                const uint2 it = src[i];
                const uint64_t p = a.p0 + it.y;
                const uint64_t w0 = a.words[p >> 5], w1 = a.words[(p >> 5) + 1];
                const uint32_t s = 2 * ((uint32_t)p & 31);
                const uint64_t x = (w0 << s) | ((w1 >> 1) >> (63 - s));
                const uint64_t fwd = x >> (64 - 2 * a.k);
                const uint64_t rc = pk_revcomp(fwd, a.k);
                canon[j] = fwd < rc ? fwd : rc;
                h[j] = it.x; pos[j] = it.y;
            }
        }
        for (uint32_t g0 = 0; g0 < a.n_local; g0 += 32) {
            uint32_t bits[IPT];
#pragma unroll
            for (int j = 0; j < IPT; j++) bits[j] = 0;
            const uint32_t ng = min(32u, a.n_local - g0);
            __syncthreads();
            if (tid < 32) s_tb[tid] = a.tables[min(g0 + tid, a.n_local - 1)];
            if (tid == 0) q_n = 0;
            for (uint32_t qq = tid; qq < (uint32_t)QCAP; qq += T) s_bits[qq] = 0;
            __syncthreads();
            for (uint32_t gs = 0; gs < ng; gs += GILP) {
                if (do_pf && tid < GILP && g0 + gs + GILP + tid < a.n_local) {      // next group's windows
                    const PkTable t = a.tables[g0 + gs + GILP + tid];
                    const uint32_t b0 = __umulhi(h_lo, t.n_buckets), b1 = __umulhi(h_hi, t.n_buckets);
                    if (b1 - b0 < 8192) l2_prefetch_bulk(t.slots + 4ull * b0, (b1 - b0 + 1) * 32);
                }
                const PkTable *tb = s_tb + gs;            // descriptors stay in shared memory (uniform LDS)
                const uint32_t nu = min((uint32_t)GILP, ng - gs);
#pragma unroll
                for (int j = 0; j < IPT; j++) {
                    const uint32_t i = tid + j * T;
                    if (i < cnt) {
                        u64x4 v[GILP];
                        const uint64_t key = canon[j];
#pragma unroll
                        for (int u = 0; u < GILP; u++) {
                            if ((uint32_t)u < nu)
                                v[u] = pk_ld_bucket_ca(tb[u].slots + 4ull * __umulhi(h[j], tb[u].n_buckets));
                        }
#pragma unroll
                        for (int u = 0; u < GILP; u++) {
                            if ((uint32_t)u < nu) {
                                const bool hit = v[u].a == key || v[u].b == key || v[u].c == key || v[u].d == key;
                                bits[j] |= (uint32_t)hit << (gs + u);
                                // slots fill in order, so the bucket is full iff its last slot is taken. A full
                                // bucket without the key means the key may sit in a later bucket: that rare
                                // walk-on is queued and resolved densely below instead of diverging here.
                                if (!hit && v[u].d != PK_EMPTY) {
                                    const uint32_t slot = atomicAdd(&q_n, 1u);
                                    if (slot < QCAP) {
                                        q_key[slot] = key;
                                        q_h[slot] = h[j];
                                        q_meta[slot] = (i << 5) | (gs + u);
                                    } else {            // queue full (pathological): walk here
                                        const uint32_t nbk = tb[u].n_buckets;
                                        uint32_t bb = __umulhi(h[j], nbk);
                                        for (uint32_t r = 1; r < nbk; r++) {
                                            bb = bb + 1 == nbk ? 0 : bb + 1;
                                            const u64x4 w = pk_ld_bucket_ca(tb[u].slots + 4ull * bb);
                                            if (w.a == key || w.b == key || w.c == key || w.d == key) { bits[j] |= 1u << (gs + u); break; }
                                            if (w.d == PK_EMPTY) break;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            {   // deferred walk-ons, one per thread, all lanes busy
                const uint32_t nq = min(q_n, (uint32_t)QCAP);
                for (uint32_t e = tid; e < nq; e += T) {
                    const uint64_t key = q_key[e];
                    const uint32_t meta = q_meta[e], g = meta & 31;
                    const PkTable t = s_tb[g];
                    const uint32_t b0 = __umulhi(q_h[e], t.n_buckets);
                    uint32_t bb = b0;
                    for (uint32_t r = 1; r < t.n_buckets; r++) {
                        bb = bb + 1 == t.n_buckets ? 0 : bb + 1;
                        const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * bb);
                        if (v.a == key || v.b == key || v.c == key || v.d == key) { atomicOr(&s_bits[meta >> 5], 1u << g); break; }
                        if (v.d == PK_EMPTY) break;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; j++)
                if (tid + j * T < cnt) bits[j] |= s_bits[tid + j * T];
            if (a.out_list) {
                for (uint32_t qq = tid; qq < (T / 32) * PP_OBINS; qq += T) (&o_wc[0][0])[qq] = 0;
                __syncthreads();
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
                        const uint32_t t = o_wc[ww][b];
                        o_wc[ww][b] = (uint16_t)run;
                        run += t;
                    }
                    if (run) o_gb[b] = atomicAdd(&a.out_cursor[(g0 / 32) * PP_OBINS + b], run);
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < IPT; j++) {
                    if (tid + j * T < cnt) {
                        const uint32_t bin = pos[j] >> a.out_shift;
                        const uint64_t slot = (((uint64_t)(g0 / 32) * PP_OBINS + bin) << a.out_shift) + o_gb[bin] + o_wc[wid][bin] + rank[j];
                        a.out_list[slot] = make_uint2(pos[j], bits[j]);
                    }
                }
                __syncthreads();
            } else {
                const uint32_t nb = min(4u, a.nbl - g0 / 8);
                const bool al4 = nb == 4 && ((a.row_stride | a.col_offset) & 3) == 0;
#pragma unroll
                for (int j = 0; j < IPT; j++) {
                    if (tid + j * T < cnt) {
                        uint8_t *dst = a.rows + (uint64_t)pos[j] * a.row_stride + a.col_offset + g0 / 8;
                        if (al4) *(uint32_t *)dst = bits[j];
                        else for (uint32_t qb = 0; qb < nb; qb++) dst[qb] = (uint8_t)(bits[j] >> (8 * qb));
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ K3, bucket-sorted variant
// The L1TEX tag stage handles ~1 sector per clock per SM, and with one 32 B bucket per lane every probe is
// a sector of its own (profiles/r1_k3_*.md: l1tex 87 % busy). Here the block first counting-sorts its
// <= T*IPT items by the hash bits just below the partition bits (shared memory), so the 32 lanes of a warp
// probe ~16 neighbouring buckets: half the tag lookups, more L1 hits, and DRAM sees 512 B runs.
// One launch covers one group of <= 32 genomes whose table descriptors travel as kernel parameters
// (constant bank: no LDS/LDG per probe).
struct ProbeArgs2 {
    const uint2 *buf;
    const uint32_t *counts;
    const unsigned long long *flat_total;
    uint32_t cap, n_regions, pb;
    const uint64_t *words;
    uint64_t p0;
    uint32_t k, ng, grp;
    uint8_t *rows;
    uint32_t row_stride, col_offset, nbl;
    int prefetch;
    uint2 *out_list;
    uint32_t *out_cursor;
    uint32_t out_shift;
    PkTable tabs[32];
};

#define PS_BINS 1024
template <int T, int IPT, int MINB>
__global__ void __launch_bounds__(T, MINB) probe_sorted_kernel(const __grid_constant__ ProbeArgs2 a) {
    constexpr int CAP = T * IPT;
    __shared__ unsigned long long s_canon[CAP];
    __shared__ uint32_t s_h[CAP], s_pos[CAP], s_bits[CAP];
    __shared__ uint32_t s_cnt[PS_BINS];
    __shared__ unsigned long long q_key[CAP];
    __shared__ uint32_t q_meta[CAP], q_h[CAP];
    __shared__ uint32_t q_n, s_ws[T / 32];
    __shared__ uint16_t o_wc[T / 32][PP_OBINS];
    __shared__ uint32_t o_gb[PP_OBINS];
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
        // ---- load items, derive canonical k-mers, counting sort by the next 10 hash bits
        for (uint32_t i = tid; i < PS_BINS; i += T) s_cnt[i] = 0;
        for (uint32_t i = tid; i < (uint32_t)CAP; i += T) s_bits[i] = 0;
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
                const uint64_t p = a.p0 + it.y;
                const uint64_t w0 = a.words[p >> 5], w1 = a.words[(p >> 5) + 1];
                const uint32_t sh = 2 * ((uint32_t)p & 31);
                const uint64_t x = (w0 << sh) | ((w1 >> 1) >> (63 - sh));
                const uint64_t fwd = x >> (64 - 2 * a.k);
                const uint64_t rc = pk_revcomp(fwd, a.k);
                canon[j] = fwd < rc ? fwd : rc;
                h[j] = it.x; pos[j] = it.y;
                rnk[j] = atomicAdd(&s_cnt[(it.x >> sshift) & (PS_BINS - 1)], 1u);
            }
        }
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
        uint32_t bits[IPT];
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const uint32_t i = tid + j * T;
            bits[j] = 0;
            if (i < cnt) { canon[j] = s_canon[i]; h[j] = s_h[i]; }
        }
        // ---- probe: genome by genome, neighbouring lanes hit neighbouring buckets
        for (uint32_t g = 0; g < a.ng; g++) {
            const PkTable t = a.tabs[g];
            if (do_pf && tid == 0 && g + 2 < a.ng) {
                const PkTable tn = a.tabs[g + 2];
                const uint32_t b0 = __umulhi(h_lo, tn.n_buckets), b1 = __umulhi(h_hi, tn.n_buckets);
                if (b1 - b0 < 8192) l2_prefetch_bulk(tn.slots + 4ull * b0, (b1 - b0 + 1) * 32);
            }
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                const uint32_t i = tid + j * T;
                if (i < cnt) {
                    const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * __umulhi(h[j], t.n_buckets));
                    const uint64_t key = canon[j];
                    const bool hit = v.a == key || v.b == key || v.c == key || v.d == key;
                    bits[j] |= (uint32_t)hit << g;
                    if (!hit && v.d != PK_EMPTY) {          // full bucket without the key: deferred walk-on
                        const uint32_t slot = atomicAdd(&q_n, 1u);
                        if (slot < (uint32_t)CAP) {
                            q_key[slot] = key; q_h[slot] = h[j]; q_meta[slot] = (i << 5) | g;
                        } else {
                            uint32_t bb = __umulhi(h[j], t.n_buckets);
                            for (uint32_t r = 1; r < t.n_buckets; r++) {
                                bb = bb + 1 == t.n_buckets ? 0 : bb + 1;
                                const u64x4 w = pk_ld_bucket_ca(t.slots + 4ull * bb);
                                if (w.a == key || w.b == key || w.c == key || w.d == key) { bits[j] |= 1u << g; break; }
                                if (w.d == PK_EMPTY) break;
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
                const uint64_t key = q_key[e];
                const uint32_t meta = q_meta[e], g = meta & 31;
                const PkTable t = a.tabs[g];
                uint32_t bb = __umulhi(q_h[e], t.n_buckets);
                for (uint32_t r = 1; r < t.n_buckets; r++) {
                    bb = bb + 1 == t.n_buckets ? 0 : bb + 1;
                    const u64x4 v = pk_ld_bucket_ca(t.slots + 4ull * bb);
                    if (v.a == key || v.b == key || v.c == key || v.d == key) { atomicOr(&s_bits[meta >> 5], 1u << g); break; }
                    if (v.d == PK_EMPTY) break;
                }
            }
        }
        for (uint32_t qq = tid; qq < (T / 32) * PP_OBINS; qq += T) (&o_wc[0][0])[qq] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const uint32_t i = tid + j * T;
            if (i < cnt) { bits[j] |= s_bits[i]; pos[j] = s_pos[i]; }
        }
        if (a.out_list) {
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
                if (run) o_gb[b] = atomicAdd(&a.out_cursor[a.grp * PP_OBINS + b], run);
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

// variants of K3 selectable at run time (PK_K3_VARIANT) while the design is being tuned
struct K3Variant { int threads, cap; size_t shmem; void (*fn)(ProbeArgs); void (*fn2)(ProbeArgs2); };
static const K3Variant k3_variants[] = {
    {PP_THREADS, PP_CAP, (size_t)PP_CAP * 16, probe_part_kernel},          // 0: genome-sequential, items in smem
    {256, 768, 0, probe_item_kernel<256, 3, 8, 2>},                         // 1
    {256, 768, 0, probe_item_kernel<256, 3, 8, 3>},                         // 2
    {256, 768, 0, probe_item_kernel<256, 3, 4, 4>},                         // 3
    {128, 768, 0, probe_item_kernel<128, 6, 4, 8>},                         // 4
    {512, 1536, 0, probe_item_kernel<512, 3, 4, 2>},                        // 5
    {256, 1536, 0, probe_item_kernel<256, 6, 4, 4>},                        // 6
    {256, 768, 0, probe_item_kernel<256, 3, 2, 5>},                         // 7
    {256, 768, 0, probe_item_kernel<256, 3, 2, 4>},                         // 8
    {256, 768, 0, probe_item_kernel<256, 3, 4, 3>},                         // 9
    {256, 768, 0, probe_item_kernel<256, 3, 1, 6>},                         // 10
    {256, 768, 0, nullptr, probe_sorted_kernel<256, 3, 5>},                 // 11
    {256, 768, 0, nullptr, probe_sorted_kernel<256, 3, 4>},                 // 12
    {256, 512, 0, nullptr, probe_sorted_kernel<256, 2, 6>},                 // 13
    {128, 768, 0, nullptr, probe_sorted_kernel<128, 6, 8>},                 // 14
};
static int g_k3_variant = 10;
void pk_part_set_variant(int v) { if (v >= 0 && v < (int)(sizeof k3_variants / sizeof k3_variants[0])) g_k3_variant = v; }

// K4: scatter the (pos, bits) lists into rows. All blocks of one bin write inside a slice of
// 2^out_shift rows, which stays in L2 until its sectors are complete.
#define UP_TILE 4096
__global__ void __launch_bounds__(256) unpermute_kernel(const uint2 *__restrict__ list, const uint32_t *__restrict__ cursor,
                                                        uint32_t out_shift, uint8_t *__restrict__ rows, uint32_t row_stride,
                                                        uint32_t col_offset, uint32_t nbl) {
    const uint32_t bin = blockIdx.y, grp = blockIdx.z;
    const uint32_t cnt = cursor[grp * PP_OBINS + bin];
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

// ------------------------------------------------------------------ host orchestration
static uint32_t ceil_log2(uint64_t v) { uint32_t b = 0; while ((1ull << b) < v) b++; return b; }

void pk_part_plan(uint64_t n, PkPartPlan *pl) {
    // mean fill <= 5/6 of the K3 block capacity (>= 20% head-room for the Poisson spread)
    const uint32_t cap = (uint32_t)k3_variants[g_k3_variant].cap;
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
}
uint32_t pk_part_obins(void) { return PP_OBINS; }

int pk_launch_probe_partitioned(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, uint32_t k,
                                const PkTable *d_tables, const PkTable *h_tables, uint32_t n_local, uint8_t *d_rows,
                                uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl, const PkPartScratch &sc,
                                int prefetch, pk_stream_t s, cudaEvent_t *evs) {
    if (!n) return 0;
    static bool attr_set = false;
    const K3Variant &kv = k3_variants[g_k3_variant];
    const size_t shmem = kv.shmem;
    if (!attr_set) {
        if (cudaFuncSetAttribute(probe_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)PP_CAP * 16)) != cudaSuccess) return -1;
        attr_set = true;
    }
    cudaMemsetAsync(sc.cursor1, 0, sizeof(uint32_t) * pl.n_regions1, s);
    if (pl.n_regions2) cudaMemsetAsync(sc.cursor2, 0, sizeof(uint32_t) * pl.n_regions2, s);
    cudaMemsetAsync(sc.spill_cursor, 0, sizeof(unsigned long long), s);
    const uint32_t n_groups = (n_local + 31) / 32;
    const bool unperm = sc.out_list != nullptr;
    if (unperm) cudaMemsetAsync(sc.out_cursor, 0, sizeof(uint32_t) * n_groups * PP_OBINS, s);
    PartArgs a{};
    a.words = d_words; a.mask64 = (const uint64_t *)d_mask; a.p0 = p0; a.n = n; a.k = k;
    a.pb1 = pl.pb1; a.pb2 = pl.pb2; a.cap1 = pl.cap1; a.cap2 = pl.cap2;
    a.buf1 = (uint2 *)sc.buf1; a.buf2 = (uint2 *)sc.buf2; a.spill = (uint2 *)sc.spill;
    a.cursor1 = sc.cursor1; a.cursor2 = sc.cursor2; a.spill_cursor = sc.spill_cursor; a.spill_cap = pl.spill_items;
    a.err = sc.err;
    a.rows = d_rows; a.row_stride = row_stride; a.col_offset = col_offset; a.nbl = (n_local + 7) / 8;
    if (evs) cudaEventRecord(evs[0], s);
    static const int rank_mode = getenv("PK_PART_RANK") ? atoi(getenv("PK_PART_RANK")) : 1;
    if (rank_mode) partition_seq_kernel<1><<<(unsigned)((n + PT_TILE - 1) / PT_TILE), PT_THREADS, 0, s>>>(a);
    else partition_seq_kernel<0><<<(unsigned)((n + PT_TILE - 1) / PT_TILE), PT_THREADS, 0, s>>>(a);
    if (evs) cudaEventRecord(evs[1], s);
    if (pl.pb2) {
        dim3 grid((pl.cap1 + PT_TILE - 1) / PT_TILE, pl.n_regions1);
        if (rank_mode) partition_fine_kernel<1><<<grid, PT_THREADS, 0, s>>>(a);
        else partition_fine_kernel<0><<<grid, PT_THREADS, 0, s>>>(a);
    }
    ProbeArgs p{};
    p.buf = pl.pb2 ? (const uint2 *)sc.buf2 : (const uint2 *)sc.buf1;
    p.counts = pl.pb2 ? sc.cursor2 : sc.cursor1;
    p.cap = pl.pb2 ? pl.cap2 : pl.cap1;
    p.n_regions = pl.pb2 ? pl.n_regions2 : pl.n_regions1;
    p.pb = pl.pb1 + pl.pb2;
    p.words = d_words; p.p0 = p0; p.k = k; p.tables = d_tables; p.n_local = n_local;
    p.rows = d_rows; p.row_stride = row_stride; p.col_offset = col_offset; p.nbl = a.nbl;
    p.prefetch = prefetch;
    p.out_list = unperm ? (uint2 *)sc.out_list : nullptr;
    p.out_cursor = sc.out_cursor;
    p.out_shift = pl.out_shift;
    if (evs) cudaEventRecord(evs[2], s);
    if (kv.fn) {
        kv.fn<<<p.n_regions, kv.threads, shmem, s>>>(p);
        if (evs) cudaEventRecord(evs[3], s);
        ProbeArgs sp = p;          // drain the spill list (normally empty: the blocks exit at once)
        sp.buf = (const uint2 *)sc.spill; sp.counts = nullptr; sp.flat_total = sc.spill_cursor; sp.cap = kv.cap; sp.pb = 0;
        kv.fn<<<148 * 2, kv.threads, shmem, s>>>(sp);
    } else {
        ProbeArgs2 p2{};
        p2.buf = p.buf; p2.counts = p.counts; p2.cap = p.cap; p2.n_regions = p.n_regions; p2.pb = p.pb;
        p2.words = d_words; p2.p0 = p0; p2.k = k;
        p2.rows = d_rows; p2.row_stride = row_stride; p2.col_offset = col_offset; p2.nbl = a.nbl;
        p2.prefetch = prefetch; p2.out_list = p.out_list; p2.out_cursor = p.out_cursor; p2.out_shift = p.out_shift;
        for (uint32_t grp = 0; grp < n_groups; grp++) {       // one launch per group of 32 genomes
            p2.grp = grp; p2.ng = n_local - 32 * grp < 32 ? n_local - 32 * grp : 32;
            for (uint32_t g = 0; g < p2.ng; g++) p2.tabs[g] = h_tables[32 * grp + g];
            kv.fn2<<<p2.n_regions, kv.threads, 0, s>>>(p2);
        }
        if (evs) cudaEventRecord(evs[3], s);
        ProbeArgs2 sp = p2;
        sp.buf = (const uint2 *)sc.spill; sp.counts = nullptr; sp.flat_total = sc.spill_cursor; sp.cap = kv.cap; sp.pb = 0;
        for (uint32_t grp = 0; grp < n_groups; grp++) {
            sp.grp = grp; sp.ng = n_local - 32 * grp < 32 ? n_local - 32 * grp : 32;
            for (uint32_t g = 0; g < sp.ng; g++) sp.tabs[g] = h_tables[32 * grp + g];
            kv.fn2<<<148 * 2, kv.threads, 0, s>>>(sp);
        }
    }
    if (evs) cudaEventRecord(evs[4], s);
    if (unperm) {
        dim3 grid((unsigned)(((1ull << pl.out_shift) + UP_TILE - 1) / UP_TILE), pl.out_bins, n_groups);
        unpermute_kernel<<<grid, 256, 0, s>>>((const uint2 *)sc.out_list, sc.out_cursor, pl.out_shift, d_rows, row_stride,
                                              col_offset, a.nbl);
    }
    if (evs) cudaEventRecord(evs[5], s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
