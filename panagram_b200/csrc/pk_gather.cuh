// Position-split exchange of the genome-sharded path.
//
// Every rank r of R holds a PLANE: its shard's w bytes of every row, planes[r][plane_rows][w], in the rank's own
// ("cat") row numbering, exported over CUDA IPC. The full rows are assembled SLICE-WISE: rank r builds only the
// rows of its own slice of the output stream, reading that slice of all R planes in place over NVLink:
//
//     rows[dst_row + i][q*w + b] = planes[q][src_row + i][b]        0 <= i < n_rows, per segment
//
// so every byte crosses NVLink once (1/R of what an all-gather moves) and every row is written once, by the
// rank that goes on to reduce, compress and store it. A slice is a list of SEGMENTS because the output stream
// drops the rows between chromosomes that the plane numbering contains (one segment per chromosome piece).
//
// No shared memory (the probe kernels of the next step keep theirs), no barriers: a thread takes one 16-byte
// chunk of plane bytes (16 / w source rows, 16-byte aligned in every plane), issues the R peer loads back to
// back, transposes in registers and stores whole rows with the widest store the row stride allows.
//
// The per-thread body is __host__ __device__ so that tests/native/gather_host_check.cpp runs the very same code
// on the CPU (TEST INFRASTRUCTURE only; the library never calls it on the host).
#ifndef PK_GATHER_CUH
#define PK_GATHER_CUH
#include <stdint.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
struct pkg_u4 { uint32_t x, y, z, w; };
#else
typedef uint4 pkg_u4;
#endif

#define PKG_MAX_RANKS 16

struct PkgSeg {                 // one run of consecutive rows
    uint64_t src_row;           // first row in the planes
    uint64_t n_rows;
    uint64_t dst_row;           // first row in the output (relative to the output pointer)
    uint64_t chunk0;            // number of chunks of all earlier segments (exclusive prefix)
};

struct PkgArgs {
    const uint8_t *planes[PKG_MAX_RANKS];
    uint32_t n_ranks, w;        // bytes per row per rank
    uint64_t plane_rows;        // rows allocated in every plane (loads never go past plane_rows * w bytes)
    const PkgSeg *segs;         // n_segs entries
    uint32_t n_segs;
    uint32_t row_stride;        // bytes between output rows
    uint32_t row_bytes;         // bytes of a full row that exist: min(n_ranks * w, row_bytes) are written
    uint8_t *rows;
    uint64_t n_chunks;
};

// rows per 16-byte chunk; a w that does not divide 16 is handled row by row (chunk = 1 row)
__host__ __device__ __forceinline__ uint32_t pkg_chunk_rows(uint32_t w) { return (w <= 16 && (16 % w) == 0) ? 16 / w : 1; }

template <int I> __host__ __device__ __forceinline__ uint32_t pkg_word(const pkg_u4 &v) { return I == 0 ? v.x : I == 1 ? v.y : I == 2 ? v.z : v.w; }
template <int I> __host__ __device__ __forceinline__ uint32_t pkg_byte(const pkg_u4 &v) { return (pkg_word<(I >> 2)>(v) >> (8 * (I & 3))) & 0xffu; }
__host__ __device__ __forceinline__ uint32_t pkg_byte_rt(const pkg_u4 &v, uint32_t i) {
    const uint32_t wd = (i >> 2) == 0 ? v.x : (i >> 2) == 1 ? v.y : (i >> 2) == 2 ? v.z : v.w;
    return (wd >> (8 * (i & 3))) & 0xffu;
}

// row I (of the chunk's 16 / W rows) out of the R loaded 16-byte pieces; KIND selects the store shape
//   1: W == 1, R % 8 == 0, stride % 8 == 0   64-bit stores        2: W == 1, R % 4 == 0, stride % 4 == 0   32-bit stores
//   3: W == 1, R == 2, stride % 2 == 0       16-bit store         4: W == 2, R even, stride % 4 == 0       32-bit stores
//   5: W == 4, stride % 4 == 0                                     6: W == 8, stride % 8 == 0
//   7: W == 16, stride % 16 == 0                                   0: bytes (any W, clipped at `full`)
template <int RMAX, int W, int I, int KIND>
__host__ __device__ __forceinline__ void pkg_store_row(const pkg_u4 (&x)[RMAX], uint32_t R, uint32_t full, uint8_t *dst) {
    if (KIND == 1) {
#pragma unroll
        for (int j = 0; j < RMAX / 8; j++)
            if ((uint32_t)(8 * j) < R) {
                const uint32_t lo = pkg_byte<I>(x[8 * j]) | pkg_byte<I>(x[8 * j + 1]) << 8 | pkg_byte<I>(x[8 * j + 2]) << 16 | pkg_byte<I>(x[8 * j + 3]) << 24;
                const uint32_t hi = pkg_byte<I>(x[8 * j + 4]) | pkg_byte<I>(x[8 * j + 5]) << 8 | pkg_byte<I>(x[8 * j + 6]) << 16 | pkg_byte<I>(x[8 * j + 7]) << 24;
                *(uint64_t *)(dst + 8 * j) = (uint64_t)lo | (uint64_t)hi << 32;
            }
    } else if (KIND == 2) {
#pragma unroll
        for (int j = 0; j < RMAX / 4; j++)
            if ((uint32_t)(4 * j) < R)
                *(uint32_t *)(dst + 4 * j) = pkg_byte<I>(x[4 * j]) | pkg_byte<I>(x[4 * j + 1]) << 8 | pkg_byte<I>(x[4 * j + 2]) << 16 | pkg_byte<I>(x[4 * j + 3]) << 24;
    } else if (KIND == 3) {
        *(uint16_t *)dst = (uint16_t)(pkg_byte<I>(x[0]) | pkg_byte<I>(x[RMAX > 1 ? 1 : 0]) << 8);
    } else if (KIND == 4) {
#pragma unroll
        for (int j = 0; j < RMAX / 2; j++)
            if ((uint32_t)(2 * j) < R) {
                const uint32_t a0 = (pkg_word<((2 * I) >> 2) & 3>(x[2 * j]) >> (8 * ((2 * I) & 3))) & 0xffffu;
                const uint32_t a1 = (pkg_word<((2 * I) >> 2) & 3>(x[2 * j + 1]) >> (8 * ((2 * I) & 3))) & 0xffffu;
                *(uint32_t *)(dst + 4 * j) = a0 | a1 << 16;
            }
    } else if (KIND == 5) {
#pragma unroll
        for (int q = 0; q < RMAX; q++)
            if ((uint32_t)q < R) *(uint32_t *)(dst + 4 * q) = pkg_word<I & 3>(x[q]);
    } else if (KIND == 6) {
#pragma unroll
        for (int q = 0; q < RMAX; q++)
            if ((uint32_t)q < R) *(uint64_t *)(dst + 8 * q) = (uint64_t)pkg_word<(2 * I) & 3>(x[q]) | (uint64_t)pkg_word<(2 * I + 1) & 3>(x[q]) << 32;
    } else if (KIND == 7) {
#pragma unroll
        for (int q = 0; q < RMAX; q++)
            if ((uint32_t)q < R) *(pkg_u4 *)(dst + 16 * q) = x[q];
    } else {
#pragma unroll
        for (int q = 0; q < RMAX; q++)
            if ((uint32_t)q < R)
                for (uint32_t b = 0; b < (uint32_t)W; b++)
                    if (q * W + b < full) dst[q * W + b] = (uint8_t)pkg_byte_rt(x[q], I * W + b);
    }
}

template <int RMAX, int W, int KIND, int I>
struct PkgRows {
    __host__ __device__ static __forceinline__ void run(const pkg_u4 (&x)[RMAX], uint32_t R, uint32_t full, uint64_t r0, const PkgSeg &sg,
                                                        uint8_t *rows, uint32_t stride, bool whole) {
        const uint64_t sr = r0 + I;
        if (whole || (sr >= sg.src_row && sr < sg.src_row + sg.n_rows))
            pkg_store_row<RMAX, W, I, KIND>(x, R, full, rows + (sg.dst_row + (sr - sg.src_row)) * stride);
        PkgRows<RMAX, W, KIND, I + 1>::run(x, R, full, r0, sg, rows, stride, whole);
    }
};
template <int RMAX, int W, int KIND> struct PkgRows<RMAX, W, KIND, 16 / W> {
    __host__ __device__ static __forceinline__ void run(const pkg_u4 (&)[RMAX], uint32_t, uint32_t, uint64_t, const PkgSeg &, uint8_t *, uint32_t, bool) {}
};

// chunk t of the slice. W = a.w when it divides 16 (compile-time: the transposition is unrolled), else 0 (byte path)
template <int RMAX, int W>
__host__ __device__ __forceinline__ void pkg_gather_chunk(const PkgArgs &a, uint64_t t) {
    uint32_t lo = 0, hi = a.n_segs;                  // segment of chunk t: the last s with segs[s].chunk0 <= t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.segs[mid].chunk0 <= t) lo = mid; else hi = mid;
    }
    const PkgSeg sg = a.segs[lo];
    const uint32_t w = a.w, R = a.n_ranks;
    const uint32_t cr = W ? 16 / (W ? W : 1) : 1;
    const uint64_t r0 = sg.src_row / cr * cr + (t - sg.chunk0) * cr;            // chunks are aligned in the planes
    const uint32_t full = R * w < a.row_bytes ? R * w : a.row_bytes;             // bytes of a row that get written
    if (W != 0 && (r0 + cr) <= a.plane_rows && (uint32_t)RMAX >= R) {
        constexpr int WW = W ? W : 1;
        pkg_u4 x[RMAX];
#pragma unroll
        for (int q = 0; q < RMAX; q++)
            if ((uint32_t)q < R) x[q] = *(const pkg_u4 *)(a.planes[q] + r0 * WW);
            else x[q] = pkg_u4{0, 0, 0, 0};
        const bool whole = r0 >= sg.src_row && r0 + cr <= sg.src_row + sg.n_rows;
        const bool exact = full == R * w;
        const uint32_t st = a.row_stride;
        if (WW == 1 && exact && (R & 7) == 0 && (st & 7) == 0) PkgRows<RMAX, WW, (WW == 1 && RMAX >= 8) ? 1 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else if (WW == 1 && exact && (R & 3) == 0 && (st & 3) == 0) PkgRows<RMAX, WW, (WW == 1 && RMAX >= 4) ? 2 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else if (WW == 1 && exact && R == 2 && (st & 1) == 0) PkgRows<RMAX, WW, (WW == 1 && RMAX >= 2) ? 3 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else if (WW == 2 && exact && (R & 1) == 0 && (st & 3) == 0) PkgRows<RMAX, WW, (WW == 2 && RMAX >= 2) ? 4 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else if (WW == 4 && exact && (st & 3) == 0) PkgRows<RMAX, WW, WW == 4 ? 5 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else if (WW == 8 && exact && (st & 7) == 0) PkgRows<RMAX, WW, WW == 8 ? 6 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else if (WW == 16 && exact && (st & 15) == 0) PkgRows<RMAX, WW, WW == 16 ? 7 : 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
        else PkgRows<RMAX, WW, 0, 0>::run(x, R, full, r0, sg, a.rows, st, whole);
    } else {
        // byte path: the last chunk of a plane, a w that does not divide 16, or more ranks than the instantiation holds
        const uint32_t crr = pkg_chunk_rows(w);
        const uint64_t rr0 = sg.src_row / crr * crr + (t - sg.chunk0) * crr;
        for (uint32_t i = 0; i < crr; i++) {
            const uint64_t sr = rr0 + i;
            if (sr < sg.src_row || sr >= sg.src_row + sg.n_rows) continue;
            uint8_t *dst = a.rows + (sg.dst_row + (sr - sg.src_row)) * a.row_stride;
            for (uint32_t q = 0; q < R; q++)
                for (uint32_t b = 0; b < w; b++)
                    if (q * w + b < full) dst[q * w + b] = a.planes[q][sr * w + b];
        }
    }
}

// ---- narrow rows (R * W <= 8 bytes: 2 or 4 ranks of 8 genomes, ...) -------------------------------------------
// Row-at-a-time stores of 2 or 4 bytes waste the store path (measured: 0.41 ms for 135 MB of 2-byte rows, 2 ranks).
// Here a thread owns one 16-byte-aligned chunk of the OUTPUT (16 / (R * W) rows), gathers its 16 / R bytes per plane
// with byte loads (unaligned in the planes; consecutive lanes read consecutive bytes, L1 absorbs the re-touches) and
// writes one 16-byte store. Needs the segments back to back in the output (dst_row = running sum of n_rows), rows of
// exactly R * W bytes at a stride of R * W, and a 16-byte aligned output. Chunks that straddle a segment boundary or the
// end take a row-by-row path.
template <int R, int W>
__host__ __device__ __forceinline__ void pkg_gather_dst_chunk(const PkgArgs &a, uint64_t t, uint64_t total_rows) {
    constexpr int RW = R * W, NR = 16 / RW;
    const uint64_t d0 = t * NR;
    uint32_t lo = 0, hi = a.n_segs;                  // segment of output row d0: the last s with segs[s].dst_row <= d0
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.segs[mid].dst_row <= d0) lo = mid; else hi = mid;
    }
    const PkgSeg sg = a.segs[lo];
    if (d0 >= sg.dst_row && d0 + NR <= sg.dst_row + sg.n_rows) {
        const uint64_t src0 = sg.src_row + (d0 - sg.dst_row);
        uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < R; q++) {
            const uint8_t *p = a.planes[q] + src0 * W;
#pragma unroll
            for (int i = 0; i < NR; i++)
#pragma unroll
                for (int b = 0; b < W; b++) {
                    const int pos = i * RW + q * W + b;
                    out[pos >> 2] |= (uint32_t)p[i * W + b] << (8 * (pos & 3));
                }
        }
        *(pkg_u4 *)(a.rows + d0 * RW) = pkg_u4{out[0], out[1], out[2], out[3]};
    } else {
        for (uint64_t d = d0; d < d0 + NR && d < total_rows; d++) {
            uint32_t l2 = 0, h2 = a.n_segs;
            while (h2 - l2 > 1) {
                const uint32_t mid = (l2 + h2) >> 1;
                if (a.segs[mid].dst_row <= d) l2 = mid; else h2 = mid;
            }
            const PkgSeg s2 = a.segs[l2];
            if (d < s2.dst_row || d >= s2.dst_row + s2.n_rows) continue;       // a hole in the output numbering
            const uint64_t sr = s2.src_row + (d - s2.dst_row);
            for (int q = 0; q < R; q++)
                for (int b = 0; b < W; b++) a.rows[d * RW + q * W + b] = a.planes[q][sr * W + b];
        }
    }
}
// usable when rows are exactly R * W <= 8 bytes (a divisor of 16) at that stride, the output is 16-byte aligned and the
// segments follow each other in the output without gaps or overlaps, in order
static inline bool pkg_dst_mode_ok(const PkgSeg *segs, uint32_t n_segs, uint32_t R, uint32_t w, uint32_t row_stride, uint32_t row_bytes,
                                   const void *rows, uint64_t *total_rows) {
    const uint32_t rw = R * w;
    if (!(rw == 2 || rw == 4 || rw == 8) || !(R == 2 || R == 4 || R == 8) || row_stride != rw || row_bytes != rw) return false;
    if (((uintptr_t)rows) & 15) return false;
    uint64_t run = 0;
    for (uint32_t s = 0; s < n_segs; s++) {
        if (segs[s].dst_row != run) return false;
        run += segs[s].n_rows;
    }
    *total_rows = run;
    return true;
}

// host side: chunk prefix of a segment list (fills chunk0; returns the total)
static inline uint64_t pkg_plan_segments(PkgSeg *segs, uint32_t n_segs, uint32_t w) {
    const uint32_t cr = pkg_chunk_rows(w);
    uint64_t total = 0;
    for (uint32_t s = 0; s < n_segs; s++) {
        segs[s].chunk0 = total;
        if (segs[s].n_rows) {
            const uint64_t first = segs[s].src_row / cr, last = (segs[s].src_row + segs[s].n_rows - 1) / cr;
            total += last - first + 1;
        }
    }
    return total;
}
#endif
