// Host FASTA reader of libpkanchor.so. Replaces the getline/stringstream loop of KMCdb::anchor_fasta
// (cpp/anchor.cpp:74-100) and Genome.iter_fasta (panagram/index.py:922-930) on the way INTO the engine: one pass
// over the file with memchr, sequences land in page-locked memory so that pk_anchor_genome / pk_engine_add_sequence
// copy them to the device at full PCIe speed.
//
// Rules (identical to panagram_b200/anchor.py::parse_fasta, which tests/test_host.py compares against):
//   * a record starts at a line whose first byte is '>'; bytes before the first such line are ignored
//   * name = header up to the first ' '  (cpp/anchor.cpp:84,97)
//   * sequence = the following lines concatenated verbatim, '\n' removed — a '\r' stays, as it does with getline
//   * strip_cr (what `kmc -fm` and Biopython see): a trailing '\r' is cut from every line, every byte < 32 is
//     dropped from the sequence (kmc_core/splitter.cpp), the name ends at the first whitespace
// gzip / bgzip input (Genome.iter_fasta is gzip-aware, index.py:922-930) is inflated with zlib first, member after
// member (a .bgz file is a chain of gzip members), then parsed by the same pass.
#include <cuda_runtime.h>
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "pk_internal.h"

struct pk_fasta {
    struct Rec { std::string name; uint64_t off, len; };
    std::vector<Rec> recs;
    uint8_t *seq = nullptr;      // all sequences back to back
    uint64_t seq_bytes = 0;
    bool pinned = false;
};

static uint8_t *alloc_host(uint64_t n, bool *pinned) {
    void *p = nullptr;
    if (cudaMallocHost(&p, n ? n : 1) == cudaSuccess) { *pinned = true; return (uint8_t *)p; }
    cudaGetLastError();          // no driver / no device: plain memory (copies then do not overlap, nothing else changes)
    *pinned = false;
    return (uint8_t *)malloc(n ? n : 1);
}

extern "C" int pk_fasta_open(const char *path, int strip_cr, pk_fasta **out) {
    if (!path || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { pk_set_error("%s: cannot open", path); return PK_EIO; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 0) { close(fd); pk_set_error("%s: cannot stat", path); return PK_EIO; }
    const size_t size = (size_t)st.st_size;
    const uint8_t *raw = nullptr;
    if (size) {
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); pk_set_error("%s: mmap failed", path); return PK_EIO; }
        madvise(m, size, MADV_SEQUENTIAL);
        raw = (const uint8_t *)m;
    }
    close(fd);
    struct Unmap { const uint8_t *p; size_t n; ~Unmap() { if (p) munmap((void *)p, n); } } unmap{raw, size};
    std::vector<uint8_t> inflated;
    const uint8_t *text = raw;
    size_t text_size = size;
    if (size >= 2 && raw[0] == 0x1F && raw[1] == 0x8B) {
        // every gzip member of the file, back to back (zlib stops at a member's end: reset and go on)
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 16) != Z_OK) { pk_set_error("%s: zlib initialisation failed", path); return PK_EIO; }
        inflated.resize(std::max<size_t>(size * 4, 1 << 20));
        zs.next_in = const_cast<Bytef *>(raw);
        size_t in_left = size, out_pos = 0;
        int zr = Z_OK;
        while (in_left || zs.avail_in) {
            if (!zs.avail_in) { const size_t n = std::min<size_t>(in_left, 1u << 30); zs.avail_in = (uInt)n; in_left -= n; }
            if (out_pos == inflated.size()) inflated.resize(inflated.size() * 2);
            const size_t room = std::min<size_t>(inflated.size() - out_pos, 1u << 30);
            zs.next_out = inflated.data() + out_pos;
            zs.avail_out = (uInt)room;
            zr = inflate(&zs, Z_NO_FLUSH);
            out_pos += room - zs.avail_out;
            if (zr == Z_STREAM_END) {
                if (!zs.avail_in && !in_left) break;
                if (inflateReset(&zs) != Z_OK) { zr = Z_DATA_ERROR; break; }
            } else if (zr != Z_OK && zr != Z_BUF_ERROR) {
                break;
            }
        }
        inflateEnd(&zs);
        if (zr != Z_STREAM_END && zr != Z_OK) { pk_set_error("%s: corrupt gzip stream (zlib %d)", path, zr); return PK_EIO; }
        inflated.resize(out_pos);
        text = inflated.data();
        text_size = out_pos;
    }
    pk_fasta *fa = new pk_fasta();
    fa->seq = alloc_host((uint64_t)text_size, &fa->pinned);
    if (!fa->seq) { delete fa; pk_set_error("out of host memory"); return PK_ENOMEM; }
    const uint8_t *p = text, *end = p + text_size;
    uint64_t w = 0;
    bool in_record = false;
    while (p < end) {
        const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
        const uint8_t *le = nl ? nl : end;               // line = [p, le)
        const uint8_t *lc = le;                          // ... content end after the optional '\r' cut
        if (strip_cr && lc > p && lc[-1] == '\r') lc--;
        if (lc > p && p[0] == '>') {
            if (in_record) fa->recs.back().len = w - fa->recs.back().off;
            const uint8_t *n0 = p + 1, *q = n0;
            while (q < lc && *q != ' ') q++;                               // header up to the first space
            if (strip_cr) {                                                // ... then its first whitespace-separated token
                auto ws = [](uint8_t c) { return c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f' || c == ' '; };
                const uint8_t *e = q;
                while (n0 < e && ws(*n0)) n0++;
                q = n0;
                while (q < e && !ws(*q)) q++;
            }
            pk_fasta::Rec r;
            r.name.assign((const char *)n0, (size_t)(q - n0));
            r.off = w; r.len = 0;
            fa->recs.push_back(r);
            in_record = true;
        } else if (in_record) {
            if (!strip_cr) {
                memcpy(fa->seq + w, p, (size_t)(le - p));
                w += (uint64_t)(le - p);
            } else {
                const size_t n = (size_t)(lc - p);
                size_t ctl = 0;
                for (size_t i = 0; i < n; i++) ctl += p[i] < 32;           // vectorised; control bytes are the exception
                if (!ctl) { memcpy(fa->seq + w, p, n); w += n; }
                else for (size_t i = 0; i < n; i++) if (p[i] >= 32) fa->seq[w++] = p[i];
            }
        }
        p = nl ? nl + 1 : end;
    }
    if (in_record) fa->recs.back().len = w - fa->recs.back().off;
    fa->seq_bytes = w;
    *out = fa;
    return PK_OK;
}

extern "C" uint32_t pk_fasta_n_records(const pk_fasta *fa) { return fa ? (uint32_t)fa->recs.size() : 0; }

extern "C" int pk_fasta_record(const pk_fasta *fa, uint32_t i, const char **name, const uint8_t **seq, uint64_t *len) {
    if (!fa || i >= fa->recs.size()) { pk_set_error("record %u out of range", i); return PK_EINVAL; }
    if (name) *name = fa->recs[i].name.c_str();
    if (seq) *seq = fa->seq + fa->recs[i].off;
    if (len) *len = fa->recs[i].len;
    return PK_OK;
}

extern "C" void pk_fasta_close(pk_fasta *fa) {
    if (!fa) return;
    if (fa->pinned) cudaFreeHost(fa->seq); else free(fa->seq);
    delete fa;
}
