// Host-side reader of KMC databases (.kmc_pre / .kmc_suf), KMC1 and KMC2 layouts.
//
// Replaces the header/LUT parsing of CKMCFile::OpenForRA and
// ReadParamsFrom_prefix_file_buf (KMC/kmc_api/kmc_file.cpp:25-53,133-172,178-325).
// Unlike the reference it does not slurp the suffix file into host RAM: records
// are streamed in chunks to the GPU, where pk_kernels.cu decodes and inserts them.
#include "pk_internal.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static bool read_all(const std::string &path, const char *marker, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { pk_set_error("cannot open %s: %s", path.c_str(), strerror(errno)); return false; }
    fseeko(f, 0, SEEK_END);
    off_t n = ftello(f);
    rewind(f);
    if (n < 8) { fclose(f); pk_set_error("%s: too short", path.c_str()); return false; }
    buf.resize((size_t)n);
    bool ok = fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    if (!ok) { pk_set_error("%s: short read", path.c_str()); return false; }
    // kmc_file.cpp:133-172 — 4-byte marker at both ends
    if (memcmp(buf.data(), marker, 4) || memcmp(buf.data() + n - 4, marker, 4)) {
        pk_set_error("%s: bad %s marker", path.c_str(), marker);
        return false;
    }
    return true;
}

template <typename T> static T rd(const uint8_t *p) { T v; memcpy(&v, p, sizeof v); return v; }

int pk_kmcdb_open_impl(const char *prefix, pk_kmcdb **out) {
    if (!prefix || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    std::vector<uint8_t> pre;
    if (!read_all(std::string(prefix) + ".kmc_pre", "KMCP", pre)) return PK_EIO;
    const uint64_t n = pre.size();
    if (n < 8 + 12) { pk_set_error("%s.kmc_pre: truncated", prefix); return PK_EIO; }
    auto db = new pk_kmcdb();
    db->prefix = prefix;
    pk_kmcdb_info &I = db->info;
    memset(&I, 0, sizeof I);
    I.kmc_version = rd<uint32_t>(pre.data() + n - 12);             // kmc_file.cpp:183
    if (I.kmc_version != 0 && I.kmc_version != 0x200) {
        pk_set_error("%s: unsupported KMC version 0x%x", prefix, I.kmc_version);
        delete db; return PK_EUNSUPPORTED;
    }
    const uint64_t header_offset = pre[n - 8];                      // fgetc: one byte (:192,:266)
    if (header_offset + 8 > n - 4 || header_offset < 40) { pk_set_error("%s: bad header offset", prefix); delete db; return PK_EIO; }
    const uint8_t *h = pre.data() + n - (header_offset + 8);
    I.kmer_length = rd<uint32_t>(h);
    I.mode = rd<uint32_t>(h + 4);
    I.counter_size = rd<uint32_t>(h + 8);
    I.lut_prefix_length = rd<uint32_t>(h + 12);
    if (I.mode != 0) { pk_set_error("%s: Quake-mode counters are not supported", prefix); delete db; return PK_EUNSUPPORTED; }
    if (I.counter_size > 8 || I.lut_prefix_length > 15 || I.lut_prefix_length > I.kmer_length) {
        pk_set_error("%s: implausible header", prefix); delete db; return PK_EIO;
    }
    const uint64_t single_lut = 1ull << (2 * I.lut_prefix_length);
    uint64_t lut_entries;   // without guard
    if (I.kmc_version == 0x200) {                                   // :188-261
        I.signature_len = rd<uint32_t>(h + 16);
        I.min_count = rd<uint32_t>(h + 20);
        I.max_count = rd<uint32_t>(h + 24);
        I.total_kmers = rd<uint64_t>(h + 28);
        I.both_strands = !h[36];
        const uint64_t sigmap_bytes = ((1ull << (2 * I.signature_len)) + 1) * 4;
        const uint64_t size = n - 8 - 4;
        if (size < sigmap_bytes + header_offset + 8) { pk_set_error("%s: truncated KMC2 prefix file", prefix); delete db; return PK_EIO; }
        lut_entries = (size - (sigmap_bytes + header_offset + 8)) / 8;
        if (lut_entries % single_lut) { pk_set_error("%s: LUT area is not a multiple of 4^lut", prefix); delete db; return PK_EIO; }
    } else {                                                        // :262-322
        I.min_count = rd<uint32_t>(h + 16);
        I.max_count = rd<uint32_t>(h + 20);
        I.total_kmers = rd<uint64_t>(h + 24);
        I.both_strands = !h[32];
        // the reference READS max_count_hi right after the 1-byte flag (:291-293), i.e. at
        // header+33, although kmc_tools wrote it at +36 (kmc1_db_writer.h:344-349); the
        // lookup filter (:1396) uses the value as read, so mirror the reader
        I.max_count += (uint64_t)rd<uint32_t>(h + 33) << 32;
        lut_entries = single_lut;
        if (4 + (lut_entries + 0) * 8 > n) { pk_set_error("%s: truncated KMC1 prefix file", prefix); delete db; return PK_EIO; }
    }
    db->single_lut = single_lut;
    db->lut.resize(lut_entries + 1);
    memcpy(db->lut.data(), pre.data() + 4, lut_entries * 8);
    db->lut[lut_entries] = I.total_kmers;   // guard (the reference stores total+1, :237/:307; see oracle note)
    db->suf_size = (I.kmer_length - I.lut_prefix_length) / 4;      // :255 / :317
    db->rec_size = db->suf_size + I.counter_size;

    const std::string sufp = std::string(prefix) + ".kmc_suf";
    FILE *f = fopen(sufp.c_str(), "rb");
    if (!f) { pk_set_error("cannot open %s: %s", sufp.c_str(), strerror(errno)); delete db; return PK_EIO; }
    char m0[4], m1[4];
    fseeko(f, 0, SEEK_END);
    off_t sn = ftello(f);
    bool ok = sn >= 8;
    if (ok) { rewind(f); ok = fread(m0, 1, 4, f) == 4; }
    if (ok) { fseeko(f, sn - 4, SEEK_SET); ok = fread(m1, 1, 4, f) == 4; }
    if (!ok || memcmp(m0, "KMCS", 4) || memcmp(m1, "KMCS", 4)) {
        fclose(f); pk_set_error("%s: bad KMCS marker", sufp.c_str()); delete db; return PK_EIO;
    }
    if ((uint64_t)sn - 8 < I.total_kmers * db->rec_size) {
        fclose(f); pk_set_error("%s: %llu records of %u bytes do not fit the file", sufp.c_str(),
                                (unsigned long long)I.total_kmers, db->rec_size); delete db; return PK_EIO;
    }
    db->suf = f;
    *out = db;
    return PK_OK;
}

// read records [first, first+count) into dst; returns false on I/O error
bool pk_kmcdb_read_records(pk_kmcdb *db, uint64_t first, uint64_t count, uint8_t *dst) {
    if (fseeko(db->suf, (off_t)(4 + first * db->rec_size), SEEK_SET)) return false;
    const size_t want = (size_t)(count * db->rec_size);
    return fread(dst, 1, want, db->suf) == want;
}

void pk_kmcdb_close_impl(pk_kmcdb *db) {
    if (!db) return;
    if (db->suf) fclose(db->suf);
    delete db;
}
