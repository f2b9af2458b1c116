/* TEST INFRASTRUCTURE ONLY — NOT PART OF THE PRODUCT.
 *
 * Plain-C CPU restatement of the reference's anchoring path, used solely as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * The product (panagram_b200/) never loads this file; it fails loudly if its
 * CUDA library is missing.
 *
 * PARITY PINNED: this restatement is checked (tests/test_oracle.py) against
 *   (1) the known-answer semantics of KMC/tests/py_kmc_api/test_py_kmc_file.py
 *       ::test_get_counters_for_read / ::test_check_kmer (:174-218) on the
 *       reads fixture at :28-32, k=17, via committed golden vectors produced by
 *       the reference's own kmc + py_kmc_api (tests/golden/make_golden.py), and
 *   (2) outputs of the UNMODIFIED reference cpp/anchor.cpp built into
 *       oracle/_ref/run_anchor (oracle/Makefile), byte for byte on the
 *       decompressed bitmap.1 / bitmap.100 and the text of chrs.tsv /
 *       bitsum.bins.tsv.
 *
 * Each function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates. Scope: k <= 32 (one uint64 row; the reference's
 * CKmerAPI supports k <= 256 with multi-row compares, kmer_api.h:373-386, which
 * for one row degenerates to a plain uint64 compare as used here).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct pko_db {
    uint32_t k, mode, counter_size, lut_prefix_len, signature_len, version;
    uint64_t min_count, max_count, total_kmers;
    int both_strands;
    uint64_t *lut;          /* prefix_file_buf */
    uint64_t lut_size;      /* entries incl. guard */
    uint64_t single_lut;    /* 4^lut_prefix_len */
    uint32_t *sigmap;       /* KMC2 only */
    uint64_t sigmap_size;
    uint8_t *suf;           /* sufix_file_buf (without markers) */
    uint64_t suf_bytes;
    uint32_t suf_size, rec_size;
} pko_db;

static int code_of(unsigned char c) {
    /* kmer_api.h:264-275 — only ACGTacgt are symbols, everything else is -1 */
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

static uint8_t *slurp(const char *path, const char *marker, uint64_t *size) {
    /* kmc_file.cpp:133-172 OpenASingleFile — leading and trailing 4-byte marker */
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    rewind(f);
    if (n < 8) { fclose(f); return NULL; }
    uint8_t *buf = (uint8_t *)malloc((size_t)n);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
    fclose(f);
    if (memcmp(buf, marker, 4) || memcmp(buf + n - 4, marker, 4)) { free(buf); return NULL; }
    *size = (uint64_t)n;
    return buf;
}

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

void pko_db_close(pko_db *db) {
    if (!db) return;
    free(db->lut); free(db->sigmap); free(db->suf); free(db);
}

/* kmc_file.cpp:25-53 OpenForRA + :178-325 ReadParamsFrom_prefix_file_buf */
pko_db *pko_db_open(const char *prefix) {
    char path[4096];
    uint64_t n = 0;
    snprintf(path, sizeof path, "%s.kmc_pre", prefix);
    uint8_t *pre = slurp(path, "KMCP", &n);
    if (!pre) return NULL;
    pko_db *db = (pko_db *)calloc(1, sizeof *db);
    db->version = rd32(pre + n - 12);                      /* :183 */
    if (db->version != 0 && db->version != 0x200) { free(pre); free(db); return NULL; }
    uint64_t header_offset = pre[n - 8];                   /* :192 / :266 (fgetc: one byte) */
    const uint8_t *h = pre + n - (header_offset + 8);
    db->k = rd32(h); db->mode = rd32(h + 4); db->counter_size = rd32(h + 8);
    db->lut_prefix_len = rd32(h + 12);
    if (db->mode != 0 || db->k > 32 || db->k == 0) { free(pre); free(db); return NULL; }
    db->single_lut = 1ull << (2 * db->lut_prefix_len);
    if (db->version == 0x200) {                            /* :188-261 */
        db->signature_len = rd32(h + 16);
        db->min_count = rd32(h + 20);
        db->max_count = rd32(h + 24);
        db->total_kmers = rd64(h + 28);
        db->both_strands = !h[36];
        db->sigmap_size = (1ull << (2 * db->signature_len)) + 1;
        uint64_t size = n - 8 - 4;                         /* minus markers, minus header_offset field */
        uint64_t lut_area = size - (db->sigmap_size * 4 + header_offset + 8);
        uint64_t last = lut_area / 8;
        db->lut_size = last + 1;
        db->lut = (uint64_t *)malloc(db->lut_size * 8);
        memcpy(db->lut, pre + 4, db->lut_size * 8);
        db->lut[last] = db->total_kmers + 1;               /* :237 */
        db->sigmap = (uint32_t *)malloc(db->sigmap_size * 4);
        memcpy(db->sigmap, pre + 4 + lut_area + 8, db->sigmap_size * 4);
    } else {                                               /* :262-322 */
        db->min_count = rd32(h + 16);
        db->max_count = rd32(h + 20);
        db->total_kmers = rd64(h + 24);
        db->both_strands = !h[32];
        /* the reader fetches max_count_hi directly after the 1-byte flag (:291-293),
           i.e. at header+33, not at the +36 the writer used (kmc1_db_writer.h:344-349) */
        db->max_count += (uint64_t)rd32(h + 33) << 32;
        db->lut_size = db->single_lut + 1;
        db->lut = (uint64_t *)malloc(db->lut_size * 8);
        memcpy(db->lut, pre + 4, db->lut_size * 8);
        db->lut[db->lut_size - 1] = db->total_kmers + 1;   /* :307 */
    }
    free(pre);
    db->suf_size = (db->k - db->lut_prefix_len) / 4;       /* :255 / :317 */
    db->rec_size = db->suf_size + db->counter_size;
    snprintf(path, sizeof path, "%s.kmc_suf", prefix);
    uint8_t *suf = slurp(path, "KMCS", &n);
    if (!suf) { pko_db_close(db); return NULL; }
    db->suf_bytes = n - 8;
    db->suf = (uint8_t *)malloc(db->suf_bytes + 16);
    memcpy(db->suf, suf + 4, db->suf_bytes);
    memset(db->suf + db->suf_bytes, 0, 16);
    free(suf);
    return db;
}

/* An in-memory KMC1 canonical DB from (kmer, counter) pairs sorted ascending by
   kmer: same LUT + sorted-suffix-record structure the kmc_tools writer emits
   (kmc1_db_writer.h:377-399), so lookups run through the very same search. */
pko_db *pko_db_from_sorted(uint32_t k, uint32_t lut_prefix_len, uint32_t counter_size,
                           const uint64_t *kmers, const uint32_t *counters, uint64_t n) {
    if (k == 0 || k > 32 || lut_prefix_len >= k || (k - lut_prefix_len) % 4) return NULL;
    pko_db *db = (pko_db *)calloc(1, sizeof *db);
    db->k = k; db->counter_size = counter_size; db->lut_prefix_len = lut_prefix_len;
    db->min_count = 1; db->max_count = 0xffffffffull; db->total_kmers = n; db->both_strands = 1;
    db->single_lut = 1ull << (2 * lut_prefix_len);
    db->lut_size = db->single_lut + 1;
    db->lut = (uint64_t *)calloc(db->lut_size, 8);
    db->suf_size = (k - lut_prefix_len) / 4;
    db->rec_size = db->suf_size + counter_size;
    db->suf_bytes = n * db->rec_size;
    db->suf = (uint8_t *)calloc(db->suf_bytes + 16, 1);
    uint32_t sbits = 2 * (k - lut_prefix_len);
    uint64_t cur = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t pfx = sbits >= 64 ? 0 : kmers[i] >> sbits;
        while (cur < pfx) db->lut[++cur] = i;
        uint8_t *rec = db->suf + i * db->rec_size;
        for (uint32_t a = 0; a < db->suf_size; a++)
            rec[a] = (uint8_t)(kmers[i] >> (8 * (db->suf_size - 1 - a)));
        for (uint32_t b = 0; b < counter_size; b++)
            rec[db->suf_size + b] = (uint8_t)(counters[i] >> (8 * b));
    }
    while (cur < db->single_lut) db->lut[++cur] = n;
    db->lut[db->lut_size - 1] = n + 1;
    return db;
}

void pko_db_info(const pko_db *db, uint64_t out[10]) {
    out[0] = db->k; out[1] = db->counter_size; out[2] = db->lut_prefix_len;
    out[3] = db->signature_len; out[4] = db->min_count; out[5] = db->max_count;
    out[6] = db->total_kmers; out[7] = (uint64_t)db->both_strands; out[8] = db->version;
    out[9] = db->mode;
}

/* kmc_file.cpp:1321-1399 BinarySearch over suffix records in [lo, hi]; the kmer is
   held as the 2k-bit integer (first base most significant) instead of the
   byte-aligned rows, which compares identically byte by byte. */
static int binary_search(const pko_db *db, int64_t lo, int64_t hi, uint64_t kmer, uint64_t *counter) {
    if (lo >= (int64_t)db->total_kmers) return 0;          /* :1323 */
    /* the guard LUT entry total_kmers+1 (:307) lets hi reach one record past the
       buffer in the reference (an out-of-bounds read that cannot match a real
       k-mer); clamp instead of reading past the end */
    if (hi >= (int64_t)db->total_kmers) hi = (int64_t)db->total_kmers - 1;
    uint32_t sbits = 8 * db->suf_size;
    uint64_t pattern = sbits >= 64 ? kmer : (kmer & ((1ull << sbits) - 1));
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        const uint8_t *p = db->suf + (uint64_t)mid * db->rec_size;
        uint64_t s = 0;
        for (uint32_t a = 0; a < db->suf_size; a++) s = (s << 8) | p[a];   /* MSB first :1346-1365 */
        if (s == pattern) {
            if (db->counter_size == 0) { *counter = 1; return 1; }        /* :1380-1381 */
            uint64_t c = 0;
            for (uint32_t b = 0; b < db->counter_size; b++) c |= (uint64_t)p[db->suf_size + b] << (8 * b);
            *counter = c;
            return c >= db->min_count && c <= db->max_count;               /* :1396 */
        }
        if (s < pattern) lo = mid + 1; else hi = mid - 1;
    }
    return 0;
}

/* kmc_file.cpp:905-925 count_for_kmer_kmc1 */
static uint32_t count_for_kmer_kmc1(const pko_db *db, uint64_t kmer) {
    uint32_t sbits = 2 * (db->k - db->lut_prefix_len);
    uint64_t prefix = sbits >= 64 ? 0 : kmer >> sbits;
    if (prefix >= db->lut_size) return 0;
    int64_t lo = (int64_t)db->lut[prefix];
    int64_t hi = (int64_t)db->lut[prefix + 1] - 1;
    uint64_t c = 0;
    if (binary_search(db, lo, hi, kmer, &c)) return (uint32_t)c;
    return 0;
}

/* kmc_file.cpp:873-900 dispatch + :954-1027 GetCountersForRead_kmc1_both_strands.
   Returns 1 on success (counters[len-k+1] filled), 0 when len < k (:878-882) or
   the DB is not a KMC1 canonical DB (the only kind the anchoring path queries). */
int pko_get_counters_for_read(const pko_db *db, const char *read, uint64_t len, uint32_t *counters) {
    uint32_t k = db->k;
    if (len < k) return 0;
    if (db->version != 0 || !db->both_strands) return 0;
    uint64_t n = len - k + 1;
    uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t i = 0, cpos = 0;
    uint64_t kmer = 0, rev = 0;
    uint32_t pos = 0;
    while (i + k - 1 < len) {
        int contains_n = 0;
        while (i < len && pos < k) {                       /* fill phase :972-993 */
            int c = code_of((unsigned char)read[i]);
            if (c < 0) {
                pos = 0; kmer = 0; rev = 0;
                ++i;
                uint64_t wrong = i - cpos;
                if (wrong > n - cpos) wrong = n - cpos;
                memset(counters + cpos, 0, wrong * sizeof(uint32_t));
                cpos += wrong;
                contains_n = 1;
                break;
            }
            /* insert2bits(pos, c) / insert2bits(rev_pos, 3-c)  (kmer_api.h:44-47) */
            kmer |= (uint64_t)c << (2 * (k - 1 - pos));
            rev |= (uint64_t)(3 - c) << (2 * pos);
            ++pos; ++i;
        }
        if (contains_n) continue;
        if (pos == k)                                       /* :996-1002; operator< kmer_api.h:373-386 */
            counters[cpos++] = count_for_kmer_kmc1(db, kmer < rev ? kmer : rev);
        else
            break;
        while (i < len) {                                   /* slide phase :1006-1019 */
            int c = code_of((unsigned char)read[i]);
            if (c < 0) { pos = 0; break; }
            rev = (rev >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));     /* SHR_insert2bits kmer_api.h:68-81 */
            kmer = ((kmer << 2) | (uint64_t)c) & mask;                    /* SHL_insert2bits kmer_api.h:54-66 */
            ++i;
            counters[cpos++] = count_for_kmer_kmc1(db, kmer < rev ? kmer : rev);
        }
        if (pos == 0) { kmer = 0; rev = 0; }                /* re-enters the fill phase, which clears */
    }
    if (cpos < n) memset(counters + cpos, 0, (n - cpos) * sizeof(uint32_t));   /* :1021-1025 */
    return 1;
}

/* Every canonical k-mer of a sequence, window by window — what `kmc -ci1 -fm` counts for a FASTA record
   (kmc_core/splitter.cpp:44-47: only ACGTacgt are symbols, every other byte ends the current super-k-mer; canonical
   form = min(kmer, reverse complement), kmer_api.h:373-386) — with the same fill/slide structure as
   pko_get_counters_for_read above. Windows holding a non-ACGT byte yield nothing. out may be NULL to only count;
   duplicates are NOT removed (the caller sorts and uniques). Returns the number of k-mers written. */
uint64_t pko_kmers_of_seq(const char *seq, uint64_t len, uint32_t k, uint64_t *out) {
    if (len < k) return 0;
    const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t kmer = 0, rev = 0, n = 0;
    uint32_t pos = 0;
    for (uint64_t i = 0; i < len; i++) {
        int c = code_of((unsigned char)seq[i]);
        if (c < 0) { pos = 0; kmer = 0; rev = 0; continue; }
        kmer = ((kmer << 2) | (uint64_t)c) & mask;
        rev = (rev >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
        if (++pos >= k) {
            if (out) out[n] = kmer < rev ? kmer : rev;
            n++;
        }
    }
    return n;
}

/* Listing of every (canonical k-mer, counter) record, as CKMCFile::ReadNextKmer
   decodes them (kmc_file.cpp:421-490) with the LUT slot -> prefix mapping of
   kmc_file.h:89 / CPrefixFileBufferForListingMode (prefix = slot mod 4^lut for
   KMC2's per-bin LUTs). Applies the min/max filter of :487. Returns the number
   written (<= total_kmers); kmers/counters may be NULL to only count. */
uint64_t pko_db_list(const pko_db *db, uint64_t *kmers, uint32_t *counters) {
    uint64_t nslots = db->lut_size - 1, out = 0;
    uint32_t sbits = 8 * db->suf_size;
    for (uint64_t j = 0; j < nslots; j++) {
        uint64_t lo = db->lut[j], hi = db->lut[j + 1];
        if (hi > db->total_kmers) hi = db->total_kmers;
        uint64_t prefix = j % db->single_lut;
        for (uint64_t r = lo; r < hi; r++) {
            const uint8_t *p = db->suf + r * db->rec_size;
            uint64_t s = 0;
            for (uint32_t a = 0; a < db->suf_size; a++) s = (s << 8) | p[a];
            uint64_t c = 1;
            if (db->counter_size) {
                c = 0;
                for (uint32_t b = 0; b < db->counter_size; b++) c |= (uint64_t)p[db->suf_size + b] << (8 * b);
                if (c < db->min_count || c > db->max_count) continue;
            }
            if (kmers) kmers[out] = (sbits >= 64 ? 0 : prefix << sbits) | s;
            if (counters) counters[out] = (uint32_t)c;
            out++;
        }
    }
    return out;
}

/* cpp/anchor.cpp:114-118 — bin length rule. Returns 0 when nkmers < 100 (the
   reference then divides by zero at :120; callers must treat it as an error). */
uint64_t pko_binlen(uint64_t nkmers) {
    uint64_t binlen = 200000;
    if (nkmers / binlen < 100) binlen = nkmers / 100;
    return binlen;
}

static int popcount32(uint32_t v) { int c = 0; while (v) { v &= v - 1; c++; } return c; }

/* cpp/anchor.cpp:112-195 KMCdb::write_bits for one chromosome, writing into
   caller buffers instead of BGZF streams:
     bitmap1   [nkmers * nbytes]                         (:138-167)
     bitmap100 [ceil(nkmers/100) * nbytes]               (:169-177; phase = chunk_start % 100)
     bin_start [nchunks], bin_hist [nchunks * (N+1)]     (:179-190)
     paircounts[N]  += per-genome column sums            (index.py:1051, Python path only)
   dbs = the ceil(N/32) bitvec DBs (cpp/anchor.cpp:21-35). Returns nkmers, or
   (uint64_t)-1 if len < k or nkmers < 100 (undefined in the reference). */
uint64_t pko_anchor_chrom(pko_db *const *dbs, int ndb, int ngenomes, const char *seq, uint64_t len,
                          uint8_t *bitmap1, uint8_t *bitmap100, uint64_t *n100_out,
                          uint64_t *bin_start, uint64_t *bin_hist, uint64_t *nbins_out,
                          uint64_t *paircounts) {
    uint32_t k = dbs[0]->k;
    if (len < k) return (uint64_t)-1;
    uint64_t nbytes = ((uint64_t)ngenomes + 7) / 8;
    uint64_t nkmers = len - k + 1;
    uint64_t binlen = pko_binlen(nkmers);
    if (binlen == 0) return (uint64_t)-1;
    uint64_t nchunks = nkmers / binlen + (nkmers % binlen != 0);          /* :120 */
    uint64_t chunk_start = 0, n100 = 0;
    uint32_t *ints = (uint32_t *)malloc((binlen + 1) * sizeof(uint32_t));
    int *popc = (int *)malloc((binlen + 1) * sizeof(int));
    for (uint64_t chunk = 0; chunk < nchunks; chunk++) {
        uint64_t chunk_end = chunk_start + binlen < nkmers ? chunk_start + binlen : nkmers;
        uint64_t cn = chunk_end - chunk_start;
        uint8_t *bytes = bitmap1 + chunk_start * nbytes;
        memset(popc, 0, cn * sizeof(int));
        memset(bytes, 0, cn * nbytes);
        uint64_t offs = 0;
        for (int dbi = 0; dbi < ndb; dbi++) {
            uint64_t nb;                                                   /* :139-145 */
            if (nbytes <= 4) nb = nbytes;
            else if (dbi == ndb - 1 && nbytes % 4 > 0) nb = nbytes % 4;
            else nb = 4;
            /* substr(chunk_start, cn + k - 1)  (:127) */
            if (!pko_get_counters_for_read(dbs[dbi], seq + chunk_start, cn + k - 1, ints)) {
                free(ints); free(popc); return (uint64_t)-1;
            }
            for (uint64_t j = 0; j < cn; j++) {                            /* :154-161 */
                for (uint64_t sh = 0; sh < nb; sh++) bytes[j * nbytes + offs + sh] = (uint8_t)(ints[j] >> (8 * sh));
                popc[j] += popcount32(ints[j]);
                if (paircounts)
                    for (int b = 0; b < 32 && dbi * 32 + b < ngenomes; b++)
                        paircounts[dbi * 32 + b] += (ints[j] >> b) & 1;
            }
            offs += nb;
        }
        for (uint64_t j = (100 - (chunk_start % 100)) % 100; j < cn; j += 100) {    /* :170-175 */
            memcpy(bitmap100 + n100 * nbytes, bytes + j * nbytes, nbytes);
            n100++;
        }
        uint64_t *hist = bin_hist + chunk * (uint64_t)(ngenomes + 1);
        memset(hist, 0, (size_t)(ngenomes + 1) * 8);
        for (uint64_t j = 0; j < cn; j++) hist[popc[j] <= ngenomes ? popc[j] : ngenomes]++;   /* :181-183 */
        bin_start[chunk] = chunk_start;
        chunk_start += binlen;
    }
    free(ints); free(popc);
    *n100_out = n100; *nbins_out = nchunks;
    return nkmers;
}
