// pk_deflate.cuh — the per-lane pieces of the on-GPU BGZF writer (pk_bgzf.cu).
//
// Replaces bgzf_write / bgzf_index_dump of htslib as cpp/anchor.cpp:46-54,102-106,167,177 uses them
// (and bgzip.BGZipWriter + `bgzip -rI` in panagram/index.py:1035-1037,1089-1094). A BGZF file is a
// sequence of independent gzip members of <= 64 KiB, each holding <= 0xff00 payload bytes; the reader
// side (index.py:793-845) only needs (a) valid gzip members and (b) the .gzi table of block starts.
// Parity with the reference is defined on the DECOMPRESSED bytes, so the encoder is free to be
// GPU-shaped: one 256-thread block per BGZF block, each thread deflates its own 255-byte sub-chunk into a
// byte-aligned piece of the member's deflate stream:
//
//     lane piece = fixed-Huffman block (BFINAL only on the last piece) [+ empty stored block = sync marker]
//
// The empty stored block (00 00 FF FF after bit padding, the Z_SYNC_FLUSH marker) re-aligns the stream to
// a byte boundary, so the 256 pieces concatenate byte-wise (5 marker bytes per 255 payload bytes at worst).
// LZ77 matches use ONE distance: the row width (`dist` bytes) — consecutive bitmap rows repeat for as long as no genome's k-mer membership changes, so
// "same as the previous row" is where the redundancy of this data is; a member whose pieces do not beat
// the raw size is emitted as a single stored block instead.
//
// The functions below are plain sequential code and also compile on the host (tests/native/) so that the
// bit-level format is checked against zlib's inflate on the CPU; the product only ever runs them on the GPU.
#ifndef PK_DEFLATE_CUH
#define PK_DEFLATE_CUH
#include <stdint.h>

#ifdef __CUDACC__
#define PKZ_FN __host__ __device__ __forceinline__
#else
#define PKZ_FN static inline
#endif

#define PKZ_PAYLOAD 0xFF00u                   // payload bytes per BGZF block (htslib BGZF_BLOCK_SIZE)
#define PKZ_LANES 256u                        // pieces (threads) per BGZF block: the encoder is a sequential loop per
                                              // piece, so its run time is the latency of ONE piece (32 pieces of 2040
                                              // bytes: 2.0 ms for 135 MB on B200, 128 of 510: 1.37 ms; profiles/)
#define PKZ_SUB (PKZ_PAYLOAD / PKZ_LANES)     // 255 bytes per piece
#define PKZ_STAGE 304u                        // staging bytes per piece: 255 * 9/8 + 10 bits + 5 + slack; multiple of 16
#define PKZ_PAD 16u                           // bytes readable past a payload (word-wise match compares over-read <= 3)
#define PKZ_MAX_MATCH 258u
#define PKZ_MIN_MATCH 3u
#define PKZ_HDR 18u                           // gzip header with the BC extra field
#define PKZ_TRAILER 8u                        // CRC32 + ISIZE
#define PKZ_EOF_BYTES 28u

struct PkzBits {
    uint64_t acc;
    uint32_t nbits;
    uint32_t pos;        // bytes written so far (multiple of 4 until pkz_finish)
    uint8_t *out;        // 4-byte aligned
};

PKZ_FN uint32_t pkz_rev(uint32_t v, uint32_t n) {          // reverse the low n bits (Huffman codes go MSB first)
#ifdef __CUDA_ARCH__
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

PKZ_FN void pkz_put(PkzBits &w, uint32_t value, uint32_t n) {      // n <= 32, value < 2^n, LSB first
    w.acc |= (uint64_t)value << w.nbits;
    w.nbits += n;
    if (w.nbits >= 32) {
        *(uint32_t *)(w.out + w.pos) = (uint32_t)w.acc;
        w.pos += 4;
        w.acc >>= 32;
        w.nbits -= 32;
    }
}
PKZ_FN void pkz_align(PkzBits &w) {                                // pad with zero bits to a byte boundary
    const uint32_t r = w.nbits & 7;
    if (r) pkz_put(w, 0, 8 - r);
}
PKZ_FN uint32_t pkz_finish(PkzBits &w) {                           // byte-aligned stream -> total bytes
    for (; w.nbits >= 8; w.nbits -= 8, w.acc >>= 8) w.out[w.pos++] = (uint8_t)w.acc;
    return w.pos;
}

// fixed Huffman literal/length code of RFC 1951 3.2.6
PKZ_FN void pkz_put_litlen(PkzBits &w, uint32_t sym) {
    if (sym < 144) pkz_put(w, pkz_rev(0x30 + sym, 8), 8);
    else if (sym < 256) pkz_put(w, pkz_rev(0x190 + (sym - 144), 9), 9);
    else if (sym < 280) pkz_put(w, pkz_rev(sym - 256, 7), 7);
    else pkz_put(w, pkz_rev(0xC0 + (sym - 280), 8), 8);
}
PKZ_FN uint32_t pkz_ilog2(uint32_t v) {       // floor(log2(v)), v >= 1
#ifdef __CUDA_ARCH__
    return 31 - __clz(v);
#else
    uint32_t r = 0;
    while (v >>= 1) r++;
    return r;
#endif
}
// length 3..258 -> symbol 257..285 + extra bits
PKZ_FN void pkz_put_length(PkzBits &w, uint32_t len) {
    const uint32_t x = len - 3;
    if (x < 8) { pkz_put_litlen(w, 257 + x); return; }
    if (x == 255) { pkz_put_litlen(w, 285); return; }
    const uint32_t e = pkz_ilog2(x) - 2;                    // extra bits: 1..5
    pkz_put_litlen(w, 265 + 4 * (e - 1) + ((x - (4u << e)) >> e));
    pkz_put(w, x & ((1u << e) - 1), e);
}
// distance 1..32768 -> 5-bit code 0..29 + extra bits
struct PkzDist { uint32_t code, ebits, eval; };
PKZ_FN PkzDist pkz_dist(uint32_t d) {
    PkzDist r;
    if (d <= 4) { r.code = d - 1; r.ebits = 0; r.eval = 0; return r; }
    const uint32_t e = pkz_ilog2(d - 1) - 1;
    r.code = 2 * e + 2 + (((d - 1) >> e) & 1);
    r.ebits = e;
    r.eval = (d - 1) & ((1u << e) - 1);
    return r;
}

// bytes p[0..3] as a little-endian word, p unaligned (device: two aligned loads + a funnel shift; at most 3 bytes
// past p + 3 are touched, never used)
PKZ_FN uint32_t pkz_ld32u(const uint8_t *p) {
#ifdef __CUDA_ARCH__
    const uint32_t sh = ((uint32_t)(uintptr_t)p & 3u) * 8u;
    const uint32_t *q = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    return sh ? __funnelshift_r(q[0], q[1], sh) : q[0];
#else
    uint32_t v;
    __builtin_memcpy(&v, p, 4);
    return v;
#endif
}
PKZ_FN uint32_t pkz_ctz(uint32_t v) {         // v != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}
// number of positions j < maxl with blk[i + j] == blk[i + j - dist], counted from j = 0 up to the first mismatch:
// bytes until i + j is word-aligned, then four at a time
PKZ_FN uint32_t pkz_match_len(const uint8_t *blk, uint32_t i, uint32_t dist, uint32_t maxl) {
    uint32_t len = 0;
    while (len < maxl && ((i + len) & 3u)) {
        if (blk[i + len] != blk[i + len - dist]) return len;
        len++;
    }
    while (len + 4 <= maxl) {
        const uint32_t x = pkz_ld32u(blk + i + len) ^ pkz_ld32u(blk + i + len - dist);
        if (x) return len + (pkz_ctz(x) >> 3);
        len += 4;
    }
    while (len < maxl && blk[i + len] == blk[i + len - dist]) len++;
    return len;
}

// Deflate bytes [s, e) of the member payload `blk` (history = blk[0, s); blk 4-byte aligned, PKZ_PAD bytes readable
// past e) into `out` as one byte-aligned piece. lit[b] = fixed-Huffman code of literal b, bit-reversed, | length << 16.
// Returns the piece length (<= PKZ_STAGE for e - s <= PKZ_SUB).
PKZ_FN uint32_t pkz_encode_piece(const uint8_t *blk, uint32_t s, uint32_t e, uint32_t dist, bool final, uint8_t *out,
                                 const uint32_t *lit) {
    PkzBits w;
    w.acc = 0; w.nbits = 0; w.pos = 0; w.out = out;
    const PkzDist dc = pkz_dist(dist);
    // the distance code and its extra bits follow every length code: one put
    const uint32_t dbits = pkz_rev(dc.code, 5) | (dc.eval << 5), dn = 5 + dc.ebits;
    pkz_put(w, final ? 1u : 0u, 1);
    pkz_put(w, 1, 2);                                        // BTYPE = 01: fixed Huffman
    uint32_t i = s;
    while (i < e) {
        uint32_t len = 0;
        if (i >= dist) len = pkz_match_len(blk, i, dist, e - i < PKZ_MAX_MATCH ? e - i : PKZ_MAX_MATCH);
        if (len >= PKZ_MIN_MATCH) {
            pkz_put_length(w, len);
            pkz_put(w, dbits, dn);
            i += len;
        } else {
            const uint32_t c = lit[blk[i]];
            pkz_put(w, c & 0xffffu, c >> 16);
            i++;
        }
    }
    pkz_put_litlen(w, 256);                                  // end of block
    if (!final) {
        pkz_put(w, 0, 3);                                    // BFINAL = 0, BTYPE = 00: stored ...
        pkz_align(w);
        pkz_put(w, 0x0000, 16);                              // ... LEN = 0
        pkz_put(w, 0xFFFF, 16);                              //     NLEN = ~LEN
    } else {
        pkz_align(w);
    }
    return pkz_finish(w);
}

// ---- CRC-32 (gzip polynomial, reflected) -------------------------------------------------------------
// tab: four 256-entry tables (slicing-by-4: tab[256 * j + b] = CRC state of byte b followed by j zero bytes), so four
// input bytes cost four INDEPENDENT lookups instead of a chain of four dependent ones. The update is linear in the
// state, so lane 0 starts from the CRC preset (0xFFFFFFFF) and the others from 0; pkz_crc_shift moves a state over
// n following zero bytes with the operators "append 2^j zero bytes" (32 x 32 bit matrices over GF(2),
// mats[j][bit]), and the XOR of the shifted lane states, complemented, is the member's CRC.
#define PKZ_CRC_TAB_WORDS 1024
PKZ_FN uint32_t pkz_crc_update(const uint32_t *tab, uint32_t crc, const uint8_t *p, uint32_t n) {
    uint32_t i = 0;
    for (; i < n && ((uintptr_t)(p + i) & 3u); i++) crc = tab[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    for (; i + 4 <= n; i += 4) {
        uint32_t wv;
#ifdef __CUDA_ARCH__
        wv = *(const uint32_t *)(p + i);
#else
        __builtin_memcpy(&wv, p + i, 4);
#endif
        crc ^= wv;
        crc = tab[768 + (crc & 0xff)] ^ tab[512 + ((crc >> 8) & 0xff)] ^ tab[256 + ((crc >> 16) & 0xff)] ^ tab[crc >> 24];
    }
    for (; i < n; i++) crc = tab[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
PKZ_FN uint32_t pkz_gf2_times(const uint32_t *mat, uint32_t vec) {
    uint32_t sum = 0;
    for (uint32_t b = 0; b < 32; b++) sum ^= ((vec >> b) & 1u) ? mat[b] : 0u;
    return sum;
}
#define PKZ_CRC_MATS 17                        // 2^0 .. 2^16 zero bytes
PKZ_FN uint32_t pkz_crc_shift(const uint32_t *mats /*[PKZ_CRC_MATS][32]*/, uint32_t crc, uint32_t nbytes) {
    for (uint32_t j = 0; nbytes; j++, nbytes >>= 1)
        if (nbytes & 1) crc = pkz_gf2_times(mats + 32 * j, crc);
    return crc;
}

// host-side construction of the tables (uploaded once by pk_bgzf.cu): tab[PKZ_CRC_TAB_WORDS], mats, lit[256]
static inline void pkz_make_tables(uint32_t *tab, uint32_t mats[PKZ_CRC_MATS * 32], uint32_t lit[256]) {
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[n] = c;
    }
    for (uint32_t n = 0; n < 256; n++)
        for (int j = 1; j < 4; j++) tab[256 * j + n] = tab[tab[256 * (j - 1) + n] & 0xff] ^ (tab[256 * (j - 1) + n] >> 8);
    for (uint32_t b = 0; b < 256; b++) {
        uint32_t code = b < 144 ? 0x30 + b : 0x190 + (b - 144), nb = b < 144 ? 8 : 9, r = 0;
        for (uint32_t i = 0; i < nb; i++) r |= ((code >> i) & 1u) << (nb - 1 - i);
        lit[b] = r | (nb << 16);
    }
    uint32_t a[32], b[32];
    a[0] = 0xEDB88320u;                                      // operator for one zero BIT
    for (int n = 1; n < 32; n++) a[n] = 1u << (n - 1);
    for (int sq = 0; sq < 3; sq++) {                         // 1 bit -> 2 -> 4 -> 8 bits
        for (int n = 0; n < 32; n++) {
            uint32_t sum = 0, v = a[n];
            for (int q = 0; v; q++, v >>= 1) if (v & 1) sum ^= a[q];
            b[n] = sum;
        }
        for (int n = 0; n < 32; n++) a[n] = b[n];
    }
    for (int j = 0; j < PKZ_CRC_MATS; j++) {                 // a = operator for 2^j zero bytes
        for (int n = 0; n < 32; n++) mats[32 * j + n] = a[n];
        for (int n = 0; n < 32; n++) {
            uint32_t sum = 0, v = a[n];
            for (int q = 0; v; q++, v >>= 1) if (v & 1) sum ^= a[q];
            b[n] = sum;
        }
        for (int n = 0; n < 32; n++) a[n] = b[n];
    }
}

// the fixed parts of a BGZF member
PKZ_FN void pkz_write_header(uint8_t *p, uint32_t member_bytes) {       // 18 bytes; BSIZE = member size - 1
    const uint8_t h[16] = {0x1F, 0x8B, 8, 4, 0, 0, 0, 0, 0, 0xFF, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; i++) p[i] = h[i];
    p[16] = (uint8_t)((member_bytes - 1) & 0xff);
    p[17] = (uint8_t)((member_bytes - 1) >> 8);
}
PKZ_FN void pkz_write_trailer(uint8_t *p, uint32_t crc, uint32_t isize) {
    for (int i = 0; i < 4; i++) { p[i] = (uint8_t)(crc >> (8 * i)); p[4 + i] = (uint8_t)(isize >> (8 * i)); }
}
PKZ_FN void pkz_write_eof(uint8_t *p) {
    const uint8_t e[28] = {0x1F, 0x8B, 8, 4, 0, 0, 0, 0, 0, 0xFF, 6, 0, 0x42, 0x43, 2, 0, 0x1B, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 28; i++) p[i] = e[i];
}
#endif
