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
__device__ __forceinline__ void block_partition(PartSmem &S, const uint32_t (&h)[PT_IPT], const uint32_t (&pos)[PT_IPT],
                                                uint32_t validmask, uint32_t shift, uint32_t nb, uint2 *__restrict__ dst,
                                                uint64_t region0, uint32_t cap, uint32_t *__restrict__ cursor,
                                                uint2 *__restrict__ spill, unsigned long long *__restrict__ spill_cursor,
                                                uint64_t spill_cap, uint32_t *__restrict__ err) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t lt = (1u << lane) - 1;
    for (uint32_t i = tid; i < PT_WARPS * PT_MAXB; i += PT_THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    uint16_t rank[PT_IPT];
#pragma unroll
    for (int j = 0; j < PT_IPT; j++) {
        const bool v = (validmask >> j) & 1;
        const uint32_t d = (h[j] >> shift) & (nb - 1);
        const uint32_t peers = __match_any_sync(0xffffffffu, v ? d : 0xffffffffu);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (v && lane == leader) {
            old = S.warp_cnt[w][d];
            S.warp_cnt[w][d] = (uint16_t)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = (uint16_t)(old + __popc(peers & lt));
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
__global__ void __launch_bounds__(PT_THREADS) partition_seq_kernel(PartArgs a) {
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
    block_partition(S, h, pos, valid, 32 - a.pb1, 1u << a.pb1, a.buf1, 0, a.cap1, a.cursor1, a.spill, a.spill_cursor,
                    a.spill_cap, a.err);
}

// K2: coarse region c = blockIdx.y, tile blockIdx.x of it -> fine regions c * 2^pb2 + next pb2 bits.
__global__ void __launch_bounds__(PT_THREADS) partition_fine_kernel(PartArgs a) {
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
    block_partition(S, h, pos, valid, 32 - a.pb1 - a.pb2, 1u << a.pb2, a.buf2, (uint64_t)c << a.pb2, a.cap2, a.cursor2,
                    a.spill, a.spill_cursor, a.spill_cap, a.err);
}

// ------------------------------------------------------------------ K3: probe one partition per block
#define PP_THREADS 256
#define PP_CAP 3072                       // items per fine partition (shared memory: 16 B each)
#define PP_IPT (PP_CAP / PP_THREADS)      // 12
#define PP_ILP 4

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

__global__ void __launch_bounds__(PP_THREADS) probe_part_kernel(ProbeArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t *s_canon = (uint64_t *)smem_raw;                    // [PP_CAP]
    uint32_t *s_h = (uint32_t *)(s_canon + PP_CAP);              // [PP_CAP]
    uint32_t *s_pos = s_h + PP_CAP;                              // [PP_CAP]
    const uint32_t tid = threadIdx.x;
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
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host orchestration
static uint32_t ceil_log2(uint64_t v) { uint32_t b = 0; while ((1ull << b) < v) b++; return b; }

void pk_part_plan(uint64_t n, PkPartPlan *pl) {
    // mean fill <= 2560 of PP_CAP = 3072 (>= 20% head-room for the Poisson spread)
    uint32_t pb = ceil_log2((n + 2559) / 2560);
    if (pb > 18) pb = 18;
    if (pb < 1) pb = 1;
    pl->pb1 = pb > 9 ? 9 : pb;
    pl->pb2 = pb - pl->pb1;
    pl->cap2 = PP_CAP;
    pl->cap1 = pl->pb2 ? PP_CAP << pl->pb2 : PP_CAP;
    pl->n_regions1 = 1u << pl->pb1;
    pl->n_regions2 = pl->pb2 ? 1u << pb : 0;
    pl->buf1_items = (uint64_t)pl->n_regions1 * pl->cap1;
    pl->buf2_items = (uint64_t)pl->n_regions2 * pl->cap2;
    pl->spill_items = n;
}

int pk_launch_probe_partitioned(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, uint32_t k,
                                const PkTable *d_tables, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride,
                                uint32_t col_offset, const PkPartPlan &pl, const PkPartScratch &sc, int prefetch,
                                pk_stream_t s, cudaEvent_t *evs) {
    if (!n) return 0;
    static bool attr_set = false;
    const size_t shmem = (size_t)PP_CAP * 16;
    if (!attr_set) {
        if (cudaFuncSetAttribute(probe_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem) != cudaSuccess) return -1;
        attr_set = true;
    }
    cudaMemsetAsync(sc.cursor1, 0, sizeof(uint32_t) * pl.n_regions1, s);
    if (pl.n_regions2) cudaMemsetAsync(sc.cursor2, 0, sizeof(uint32_t) * pl.n_regions2, s);
    cudaMemsetAsync(sc.spill_cursor, 0, sizeof(unsigned long long), s);
    PartArgs a{};
    a.words = d_words; a.mask64 = (const uint64_t *)d_mask; a.p0 = p0; a.n = n; a.k = k;
    a.pb1 = pl.pb1; a.pb2 = pl.pb2; a.cap1 = pl.cap1; a.cap2 = pl.cap2;
    a.buf1 = (uint2 *)sc.buf1; a.buf2 = (uint2 *)sc.buf2; a.spill = (uint2 *)sc.spill;
    a.cursor1 = sc.cursor1; a.cursor2 = sc.cursor2; a.spill_cursor = sc.spill_cursor; a.spill_cap = pl.spill_items;
    a.err = sc.err;
    a.rows = d_rows; a.row_stride = row_stride; a.col_offset = col_offset; a.nbl = (n_local + 7) / 8;
    if (evs) cudaEventRecord(evs[0], s);
    partition_seq_kernel<<<(unsigned)((n + PT_TILE - 1) / PT_TILE), PT_THREADS, 0, s>>>(a);
    if (evs) cudaEventRecord(evs[1], s);
    if (pl.pb2) {
        dim3 grid((pl.cap1 + PT_TILE - 1) / PT_TILE, pl.n_regions1);
        partition_fine_kernel<<<grid, PT_THREADS, 0, s>>>(a);
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
    if (evs) cudaEventRecord(evs[2], s);
    probe_part_kernel<<<p.n_regions, PP_THREADS, shmem, s>>>(p);
    if (evs) cudaEventRecord(evs[3], s);
    ProbeArgs sp = p;          // drain the spill list (normally empty: the blocks exit at once)
    sp.buf = (const uint2 *)sc.spill; sp.counts = nullptr; sp.flat_total = sc.spill_cursor; sp.cap = PP_CAP; sp.pb = 0;
    probe_part_kernel<<<148 * 2, PP_THREADS, shmem, s>>>(sp);
    if (evs) cudaEventRecord(evs[4], s);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
