/* TEST INFRASTRUCTURE ONLY (oracle/): a minimal zlib-backed stand-in for the
 * subset of htslib's BGZF writer API that the reference's cpp/anchor.cpp uses
 * (cpp/anchor.cpp:43-55,102-106,167,177). htslib is not present in this image
 * and cannot be fetched, so the unmodified reference anchor.cpp is compiled
 * against this header. Parity is defined on the DECOMPRESSED bytes.
 *
 * Format written: standard BGZF (RFC1952 members with a 'BC' extra subfield
 * holding block_size-1), payloads of at most 0xff00 bytes, 28-byte EOF block;
 * .gzi = u64 n, then n x (u64 compressed_offset, u64 uncompressed_offset), one
 * entry per block boundary after the first block (what index.py:793-799 reads).
 */
#ifndef PK_ORACLE_BGZF_SHIM_H
#define PK_ORACLE_BGZF_SHIM_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <zlib.h>

#define PK_BGZF_PAYLOAD 0xff00

typedef struct BGZF {
    FILE *fp;
    uint8_t *buf;       /* pending uncompressed bytes */
    size_t fill;
    uint64_t coff, uoff; /* offsets of the next block */
    int want_index;
    uint64_t *idx;      /* pairs */
    size_t nidx, capidx;
} BGZF;

static inline BGZF *bgzf_open(const char *path, const char *mode) {
    (void)mode;
    FILE *fp = fopen(path, "wb");
    if (!fp) return NULL;
    BGZF *b = (BGZF *)calloc(1, sizeof(BGZF));
    b->fp = fp;
    b->buf = (uint8_t *)malloc(PK_BGZF_PAYLOAD);
    return b;
}

static inline int bgzf_index_build_init(BGZF *b) { b->want_index = 1; return 0; }

static inline int pk_bgzf_flush_block(BGZF *b) {
    if (b->fill == 0) return 0;
    uint8_t out[0x10000];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    zs.next_in = b->buf; zs.avail_in = (uInt)b->fill;
    zs.next_out = out + 18; zs.avail_out = sizeof(out) - 18 - 8;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return -1; }
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    size_t bsize = clen + 18 + 8;
    static const uint8_t hdr[16] = {0x1f,0x8b,8,4,0,0,0,0,0,0xff,6,0,'B','C',2,0};
    memcpy(out, hdr, 16);
    out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), b->buf, (uInt)b->fill);
    uint32_t isz = (uint32_t)b->fill;
    memcpy(out + 18 + clen, &crc, 4);
    memcpy(out + 18 + clen + 4, &isz, 4);
    if (fwrite(out, 1, bsize, b->fp) != bsize) return -1;
    b->coff += bsize; b->uoff += b->fill; b->fill = 0;
    if (b->want_index) {
        if (b->nidx == b->capidx) {
            b->capidx = b->capidx ? b->capidx * 2 : 1024;
            b->idx = (uint64_t *)realloc(b->idx, b->capidx * 2 * sizeof(uint64_t));
        }
        b->idx[2 * b->nidx] = b->coff; b->idx[2 * b->nidx + 1] = b->uoff; b->nidx++;
    }
    return 0;
}

static inline ssize_t bgzf_write(BGZF *b, const void *data, size_t len) {
    const uint8_t *p = (const uint8_t *)data;
    size_t left = len;
    while (left) {
        size_t n = PK_BGZF_PAYLOAD - b->fill;
        if (n > left) n = left;
        memcpy(b->buf + b->fill, p, n);
        b->fill += n; p += n; left -= n;
        if (b->fill == PK_BGZF_PAYLOAD && pk_bgzf_flush_block(b) < 0) return -1;
    }
    return (ssize_t)len;
}

static inline int bgzf_index_dump(BGZF *b, const char *bname, const char *suffix) {
    if (pk_bgzf_flush_block(b) < 0) return -1;
    char path[4096];
    snprintf(path, sizeof path, "%s%s", bname, suffix ? suffix : "");
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    /* one entry per flushed data block (= start of the following block); the
       implicit (0,0) entry of the first block is not stored */
    uint64_t n = b->nidx;
    fwrite(&n, 8, 1, f);
    fwrite(b->idx, 16, n, f);
    fclose(f);
    return 0;
}

static inline int bgzf_close(BGZF *b) {
    int rc = pk_bgzf_flush_block(b);
    static const uint8_t eof[28] = {0x1f,0x8b,8,4,0,0,0,0,0,0xff,6,0,'B','C',2,0,0x1b,0,3,0,0,0,0,0,0,0,0,0};
    fwrite(eof, 1, 28, b->fp);
    fclose(b->fp);
    free(b->buf); free(b->idx); free(b);
    return rc;
}
#endif
