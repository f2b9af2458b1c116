// TEST INFRASTRUCTURE. Compiles the per-lane pieces of the on-GPU BGZF writer
// (panagram_b200/csrc/pk_deflate.cuh) for the host and runs them the way the kernels of pk_bgzf.cu do
// (one "warp" per BGZF block, 32 lanes), so that the bit-level format can be checked against zlib's
// inflate without a GPU. Usage: bgzf_host_check <in> <out.gz> <out.gzi> <dist>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../panagram_b200/csrc/pk_deflate.cuh"

int main(int argc, char **argv) {
    if (argc != 5) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    std::vector<uint8_t> in;
    uint8_t buf[1 << 16];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + r);
    fclose(f);
    const uint32_t dist = (uint32_t)atoi(argv[4]);
    static uint32_t tab[PKZ_CRC_TAB_WORDS], mats[PKZ_CRC_MATS * 32], lit[256];
    pkz_make_tables(tab, mats, lit);
    std::vector<uint8_t> out;
    std::vector<uint64_t> gzi;
    const uint64_t n = in.size();
    in.resize(n + PKZ_PAD);              // the encoder's word-wise compares may read (never use) a few bytes past the end
    const uint64_t nblocks = (n + PKZ_PAYLOAD - 1) / PKZ_PAYLOAD;
    std::vector<uint8_t> stage(PKZ_LANES * PKZ_STAGE + 16);
    uint8_t *st = stage.data() + ((16 - ((uintptr_t)stage.data() & 15)) & 15);
    for (uint64_t b = 0; b < nblocks; b++) {
        const uint8_t *blk = in.data() + b * PKZ_PAYLOAD;
        const uint32_t blen = (uint32_t)(n - b * PKZ_PAYLOAD < PKZ_PAYLOAD ? n - b * PKZ_PAYLOAD : PKZ_PAYLOAD);
        uint32_t sizes[PKZ_LANES], crcs[PKZ_LANES], lens[PKZ_LANES], total = 0;
        for (uint32_t l = 0; l < PKZ_LANES; l++) {
            const uint32_t s = l * PKZ_SUB < blen ? l * PKZ_SUB : blen;
            const uint32_t e = (l + 1) * PKZ_SUB < blen ? (l + 1) * PKZ_SUB : blen;
            lens[l] = e - s;
            sizes[l] = e > s ? pkz_encode_piece(blk, s, e, dist, e == blen, st + l * PKZ_STAGE, lit) : 0;
            if (sizes[l] > PKZ_STAGE) { fprintf(stderr, "piece overflow %u\n", sizes[l]); return 4; }
            total += sizes[l];
            crcs[l] = pkz_crc_update(tab, l == 0 ? 0xFFFFFFFFu : 0u, blk + s, e - s);
        }
        uint32_t crc = 0, after = 0;
        for (int l = PKZ_LANES - 1; l >= 0; l--) { crc ^= pkz_crc_shift(mats, crcs[l], after); after += lens[l]; }
        crc ^= 0xFFFFFFFFu;
        const bool stored = total >= blen + 5;
        const uint32_t cdata = stored ? blen + 5 : total;
        const uint32_t member = PKZ_HDR + cdata + PKZ_TRAILER;
        if (b) { gzi.push_back(out.size()); gzi.push_back(b * PKZ_PAYLOAD); }
        const size_t o = out.size();
        out.resize(o + member);
        pkz_write_header(&out[o], member);
        uint8_t *d = &out[o + PKZ_HDR];
        if (stored) {
            d[0] = 1; d[1] = blen & 0xff; d[2] = blen >> 8; d[3] = ~blen & 0xff; d[4] = (~blen >> 8) & 0xff;
            memcpy(d + 5, blk, blen);
        } else {
            for (uint32_t l = 0; l < PKZ_LANES; l++) { memcpy(d, st + l * PKZ_STAGE, sizes[l]); d += sizes[l]; }
        }
        pkz_write_trailer(&out[o + PKZ_HDR + cdata], crc, blen);
    }
    const size_t o = out.size();
    out.resize(o + PKZ_EOF_BYTES);
    pkz_write_eof(&out[o]);
    f = fopen(argv[2], "wb"); fwrite(out.data(), 1, out.size(), f); fclose(f);
    f = fopen(argv[3], "wb");
    const uint64_t cnt = gzi.size() / 2;
    fwrite(&cnt, 8, 1, f); fwrite(gzi.data(), 8, gzi.size(), f); fclose(f);
    return 0;
}
