// On-GPU BGZF writer: device bytes -> the image of a .gz (BGZF) file + the image of its .gzi index.
// Replaces bgzf_write / bgzf_index_dump / bgzf_close as cpp/anchor.cpp:46-54,102-106,167,177 calls them
// (Python path: bgzip.BGZipWriter + `bgzip -rI`, panagram/index.py:1035-1037,1089-1094). See
// pk_deflate.cuh for the format; three kernels:
//
//   bgzf_encode   one 256-thread block per BGZF block (0xff00 payload bytes): every thread deflates its
//                 255-byte sub-chunk into a staging slot and CRCs it; the block combines sizes and CRCs
//   bgzf_scan     one block: exclusive scan of the member sizes -> member offsets, the .gzi image, totals
//   bgzf_assemble one block per member: header + pieces (or one stored block) + trailer, contiguous
#include <cuda_runtime.h>

#include "pk_deflate.cuh"
#include "pk_internal.h"

struct PkzMeta { uint32_t cdata, crc, isize, stored; };

// one block (PKZ_LANES threads) per BGZF member; thread l deflates and CRCs piece l. The member's payload is
// staged in shared memory first: every thread walks its own 255 bytes (matches four at a time), and out of global memory
// those byte loads missed L1 (16 resident blocks x 64 KB of payload per SM) and cost an L2 round trip each:
// 1.07 ms for 135 MB, ~1100 cycles per byte per thread (profiles/r1h_launches.csv).
__global__ void __launch_bounds__(PKZ_LANES) bgzf_encode_kernel(const uint8_t *__restrict__ in, uint64_t n, uint32_t dist,
                                                                uint8_t *__restrict__ stage, uint16_t *__restrict__ piece_sizes,
                                                                PkzMeta *__restrict__ meta, const uint32_t *__restrict__ g_tables,
                                                                const uint64_t m0) {
    extern __shared__ __align__(16) uint8_t s_blk[];           // [PKZ_PAYLOAD + PKZ_PAD]
    __shared__ uint32_t s_tab[PKZ_CRC_TAB_WORDS + PKZ_CRC_MATS * 32 + 256];
    __shared__ uint32_t s_crc[PKZ_LANES / 32], s_size[PKZ_LANES / 32];
    const uint32_t l = threadIdx.x;
    const uint64_t b = m0 + blockIdx.x;                        // member of the stream
    const uint8_t *blk = in + b * PKZ_PAYLOAD;
    const uint32_t blen = (uint32_t)(n - b * PKZ_PAYLOAD < PKZ_PAYLOAD ? n - b * PKZ_PAYLOAD : PKZ_PAYLOAD);
    for (uint32_t i = l; i < PKZ_CRC_TAB_WORDS + PKZ_CRC_MATS * 32 + 256; i += PKZ_LANES) s_tab[i] = g_tables[i];
    if (l < PKZ_PAD) s_blk[blen + l] = 0;
    if ((((uintptr_t)blk) & 15) == 0) {
        for (uint32_t i = l * 16; i + 16 <= blen; i += PKZ_LANES * 16) *(uint4 *)(s_blk + i) = *(const uint4 *)(blk + i);
        for (uint32_t i = (blen & ~15u) + l; i < blen; i += PKZ_LANES) s_blk[i] = blk[i];
    } else {
        for (uint32_t i = l; i < blen; i += PKZ_LANES) s_blk[i] = blk[i];
    }
    __syncthreads();
    const uint32_t s = l * PKZ_SUB < blen ? l * PKZ_SUB : blen;
    const uint32_t e = (l + 1) * PKZ_SUB < blen ? (l + 1) * PKZ_SUB : blen;
    uint32_t size = 0;
    if (e > s) size = pkz_encode_piece(s_blk, s, e, dist, e == blen, stage + (b * PKZ_LANES + l) * PKZ_STAGE,
                                       s_tab + PKZ_CRC_TAB_WORDS + PKZ_CRC_MATS * 32);
    piece_sizes[b * PKZ_LANES + l] = (uint16_t)size;
    uint32_t crc = pkz_crc_update(s_tab, l == 0 ? 0xFFFFFFFFu : 0u, s_blk + s, e - s);
    crc = pkz_crc_shift(s_tab + PKZ_CRC_TAB_WORDS, crc, blen - e);           // bytes that follow this piece
    crc = __reduce_xor_sync(0xffffffffu, crc);
    const uint32_t total = __reduce_add_sync(0xffffffffu, size);
    if ((l & 31) == 0) { s_crc[l >> 5] = crc; s_size[l >> 5] = total; }
    __syncthreads();
    if (l == 0) {
        uint32_t c = 0xFFFFFFFFu, t = 0;
        for (uint32_t w = 0; w < PKZ_LANES / 32; w++) { c ^= s_crc[w]; t += s_size[w]; }
        PkzMeta m;
        m.stored = t >= blen + 5;
        m.cdata = m.stored ? blen + 5 : t;
        m.crc = c;
        m.isize = blen;
        meta[b] = m;
    }
}

// member offsets (exclusive scan of 26 + cdata), .gzi image = uint64 n_entries, then (compressed offset,
// uncompressed offset) of every member after the first (what bgzf_index_dump writes and
// Genome.load_bgz_blocks reads, index.py:793-799), totals[0] = bytes of the .gz image incl. the EOF member,
// totals[1] = bytes of the .gzi image.
__global__ void __launch_bounds__(1024) bgzf_scan_kernel(const PkzMeta *__restrict__ meta, uint64_t nblocks, unsigned long long *__restrict__ coff,
                                                         unsigned long long *__restrict__ gzi, unsigned long long *__restrict__ totals,
                                                         uint8_t *__restrict__ out) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nblocks; base += 1024) {
        const uint64_t b = base + tid;
        const unsigned long long v = b < nblocks ? (unsigned long long)(PKZ_HDR + meta[b].cdata + PKZ_TRAILER) : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += y;
        }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        unsigned long long woff = 0;
        for (uint32_t ww = 0; ww < w; ww++) woff += s_warp[ww];
        const unsigned long long off = s_carry + woff + inc - v;
        if (b < nblocks) {
            coff[b] = off;
            if (b > 0) { gzi[1 + 2 * (b - 1)] = off; gzi[2 + 2 * (b - 1)] = b * (unsigned long long)PKZ_PAYLOAD; }
        }
        __syncthreads();
        if (tid == 1023) s_carry = off + v;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned long long end = s_carry;
        gzi[0] = nblocks ? nblocks - 1 : 0;
        totals[0] = end + PKZ_EOF_BYTES;
        totals[1] = 8 + 16 * (nblocks ? nblocks - 1 : 0);
        pkz_write_eof(out + end);
    }
}

__global__ void __launch_bounds__(256) bgzf_assemble_kernel(const uint8_t *__restrict__ in, uint64_t n, const uint8_t *__restrict__ stage,
                                                            const uint16_t *__restrict__ piece_sizes, const PkzMeta *__restrict__ meta,
                                                            const unsigned long long *__restrict__ coff, uint8_t *__restrict__ out) {
    __shared__ uint32_t s_off[PKZ_LANES + 1];
    const uint64_t b = blockIdx.x;
    const PkzMeta m = meta[b];
    uint8_t *dst = out + coff[b];
    const uint32_t tid = threadIdx.x;
    __shared__ uint32_t s_wsum[PKZ_LANES / 32];
    if (tid < PKZ_LANES) {
        const uint32_t sz = piece_sizes[b * PKZ_LANES + tid];
        uint32_t inc = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if ((tid & 31) >= (uint32_t)o) inc += y;
        }
        s_off[tid + 1] = inc;                  // inclusive within the warp; warp offsets added below
        if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
        if (tid == 0) s_off[0] = 0;
    }
    __syncthreads();
    if (tid < PKZ_LANES) {
        uint32_t add = 0;
        for (uint32_t w = 0; w < (tid >> 5); w++) add += s_wsum[w];
        s_off[tid + 1] += add;
    }
    if (tid == 160) pkz_write_header(dst, PKZ_HDR + m.cdata + PKZ_TRAILER);
    if (tid == 192) pkz_write_trailer(dst + PKZ_HDR + m.cdata, m.crc, m.isize);
    __syncthreads();
    uint8_t *d = dst + PKZ_HDR;
    if (m.stored) {
        if (tid == 0) {
            d[0] = 1;                                             // BFINAL = 1, BTYPE = 00
            d[1] = (uint8_t)(m.isize & 0xff); d[2] = (uint8_t)(m.isize >> 8);
            d[3] = (uint8_t)(~m.isize & 0xff); d[4] = (uint8_t)((~m.isize >> 8) & 0xff);
        }
        const uint8_t *src = in + b * PKZ_PAYLOAD;
        for (uint32_t i = tid; i < m.isize; i += 256) d[5 + i] = src[i];
    } else {
        for (uint32_t l = tid >> 5; l < PKZ_LANES; l += 8) {          // one warp per piece
            const uint32_t o = s_off[l], sz = s_off[l + 1] - o;
            const uint8_t *src = stage + (b * PKZ_LANES + l) * PKZ_STAGE;
            for (uint32_t i = tid & 31; i < sz; i += 32) d[o + i] = src[i];
        }
    }
}

// ------------------------------------------------------------------ host side
uint64_t pk_bgzf_blocks_impl(uint64_t n) { return (n + PKZ_PAYLOAD - 1) / PKZ_PAYLOAD; }
uint64_t pk_bgzf_bound_impl(uint64_t n) { return pk_bgzf_blocks_impl(n) * (uint64_t)(PKZ_HDR + PKZ_PAYLOAD + 5 + PKZ_TRAILER) + PKZ_EOF_BYTES; }
uint64_t pk_bgzf_gzi_bound_impl(uint64_t n) { const uint64_t nb = pk_bgzf_blocks_impl(n); return 8 + 16 * (nb ? nb - 1 : 0); }
uint64_t pk_bgzf_scratch_bytes(uint64_t n) {
    const uint64_t nb = pk_bgzf_blocks_impl(n);
    // stage | coff (u64) | meta | piece sizes (u16), each 256-byte aligned
    return ((nb * PKZ_LANES * PKZ_STAGE + 255) & ~255ull) + ((nb * 8 + 255) & ~255ull) + ((nb * sizeof(PkzMeta) + 255) & ~255ull) +
           ((nb * PKZ_LANES * 2 + 255) & ~255ull) + 256;
}
void pk_bgzf_tables_host(uint32_t *dst /*[PK_BGZF_TABLE_WORDS]*/) {
    pkz_make_tables(dst, dst + PKZ_CRC_TAB_WORDS, dst + PKZ_CRC_TAB_WORDS + PKZ_CRC_MATS * 32);
}

// d_tables: PK_BGZF_TABLE_WORDS uint32 on the device; d_scratch: pk_bgzf_scratch_bytes(n) bytes.
struct BgzfScratch { uint8_t *stage; unsigned long long *coff; PkzMeta *meta; uint16_t *piece_sizes; };
static BgzfScratch bgzf_scratch(uint8_t *d_scratch, uint64_t nb) {
    BgzfScratch r;
    uint8_t *p = d_scratch;
    r.stage = p; p += (nb * PKZ_LANES * PKZ_STAGE + 255) & ~255ull;
    r.coff = (unsigned long long *)p; p += (nb * 8 + 255) & ~255ull;
    r.meta = (PkzMeta *)p; p += (nb * sizeof(PkzMeta) + 255) & ~255ull;
    r.piece_sizes = (uint16_t *)p;
    return r;
}
// members [m0, m1) of the n-byte stream: deflate + CRC into the scratch. The bytes of those members must be final;
// ranges may be encoded in any order and on any stream, pk_launch_bgzf_finish must be ordered after all of them.
void pk_launch_bgzf_encode(const uint8_t *d_in, uint64_t n, uint32_t dist, uint64_t m0, uint64_t m1, uint8_t *d_scratch,
                           const uint32_t *d_tables, pk_stream_t s) {
    const uint64_t nb = pk_bgzf_blocks_impl(n);
    if (m1 > nb) m1 = nb;
    if (m1 <= m0) return;
    const BgzfScratch sc = bgzf_scratch(d_scratch, nb);
    if (dist < 1) dist = 1;
    if (dist > 32768) dist = 32768;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(bgzf_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PKZ_PAYLOAD + PKZ_PAD)); attr_set = true; }
    bgzf_encode_kernel<<<(unsigned)(m1 - m0), PKZ_LANES, PKZ_PAYLOAD + PKZ_PAD, s>>>(d_in, n, dist, sc.stage, sc.piece_sizes, sc.meta, d_tables, m0);
}
// member offsets, the .gzi image, totals, and the members laid out contiguously in d_out
void pk_launch_bgzf_finish(const uint8_t *d_in, uint64_t n, uint8_t *d_out, unsigned long long *d_gzi, unsigned long long *d_totals,
                           uint8_t *d_scratch, pk_stream_t s) {
    const uint64_t nb = pk_bgzf_blocks_impl(n);
    const BgzfScratch sc = bgzf_scratch(d_scratch, nb);
    bgzf_scan_kernel<<<1, 1024, 0, s>>>(sc.meta, nb, sc.coff, d_gzi, d_totals, d_out);
    if (nb) bgzf_assemble_kernel<<<(unsigned)nb, 256, 0, s>>>(d_in, n, sc.stage, sc.piece_sizes, sc.meta, sc.coff, d_out);
}
void pk_launch_bgzf(const uint8_t *d_in, uint64_t n, uint32_t dist, uint8_t *d_out, unsigned long long *d_gzi,
                    unsigned long long *d_totals, uint8_t *d_scratch, const uint32_t *d_tables, pk_stream_t s) {
    pk_launch_bgzf_encode(d_in, n, dist, 0, pk_bgzf_blocks_impl(n), d_scratch, d_tables, s);
    pk_launch_bgzf_finish(d_in, n, d_out, d_gzi, d_totals, d_scratch, s);
}
