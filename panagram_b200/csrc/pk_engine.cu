// Host side of libpkanchor.so: engine lifetime, table construction, the pipelined
// per-chromosome anchoring call and the C ABI declared in include/pk_anchor.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "pk_internal.h"

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void pk_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            pk_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? PK_ENOMEM : PK_ECUDA;                         \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------ engine
struct HostTable {
    PkTable dev{nullptr, 0, 0};
    uint64_t capacity = 0;      // keys reserved for
    uint64_t n_keys = 0, n_overflow = 0;
    bool reserved = false;
};

struct Slot {                    // one pipeline slot of pk_anchor_chrom
    cudaStream_t stream = nullptr;
    uint8_t *d_ascii = nullptr;
    uint64_t *d_words = nullptr;
    uint32_t *d_mask = nullptr;
    uint8_t *d_rows = nullptr, *d_low = nullptr;
    cudaEvent_t ev[6] = {};      // h2d start, pack start, probe start, reduce start, d2h start, end
};

struct pk_engine {
    pk_config cfg{};
    uint32_t n_local = 0, row_bytes = 0;
    std::vector<HostTable> tabs;
    PkTable *d_tables = nullptr;
    unsigned long long *d_counters = nullptr;   // [3 * n_local]
    bool finalized = false;
    cudaStream_t stream = nullptr;              // build stream / default stream for device-level calls
    static const int kSlots = 3;
    Slot slots[kSlots];
    uint64_t chunk = 0;                         // positions per chunk
    bool slots_ready = false;
    unsigned long long *d_hist = nullptr;       // per-chromosome bin histogram
    uint64_t hist_cap = 0;
    unsigned long long *d_colsums = nullptr;    // [n_local]
    // staging for KMC ingestion
    uint8_t *h_stage = nullptr, *d_stage = nullptr;
    size_t stage_bytes = 0;
    pk_stats stats{};
};

static int set_device(const pk_engine *e) {
    CU(cudaSetDevice(e->cfg.device));
    return PK_OK;
}

static bool is_local(const pk_engine *e, uint32_t g) { return g >= e->cfg.genome_begin && g < e->cfg.genome_end; }

extern "C" int pk_abi_version(void) { return PK_ABI_VERSION; }
extern "C" const char *pk_last_error(void) { return g_err; }
extern "C" int pk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int pk_kmcdb_open(const char *prefix, pk_kmcdb **out) { return pk_kmcdb_open_impl(prefix, out); }
extern "C" int pk_kmcdb_info_get(const pk_kmcdb *db, pk_kmcdb_info *info) {
    if (!db || !info) { pk_set_error("null argument"); return PK_EINVAL; }
    *info = db->info;
    return PK_OK;
}
extern "C" void pk_kmcdb_close(pk_kmcdb *db) { pk_kmcdb_close_impl(db); }

extern "C" uint64_t pk_bin_len(const pk_config *cfg, uint64_t nkmers) {
    // cpp/anchor.cpp:114-118: binlen = 200000; if (nkmers / binlen < 100) binlen = nkmers / 100
    const uint64_t maxlen = cfg && cfg->max_bin_len ? cfg->max_bin_len : 200000;
    const uint64_t mincnt = cfg && cfg->min_bin_count ? cfg->min_bin_count : 100;
    uint64_t binlen = maxlen;
    if (nkmers / binlen < mincnt) binlen = nkmers / mincnt;
    return binlen;
}

extern "C" uint64_t pk_packed_words(uint64_t len) { return ((len + 31) / 32 + 4 + 1) & ~1ull; }

extern "C" int pk_engine_create(const pk_config *cfg, pk_engine **out) {
    if (!cfg || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    if (cfg->k < 1 || cfg->k > 32) { pk_set_error("k=%u unsupported (1..32)", cfg->k); return PK_EUNSUPPORTED; }
    if (cfg->n_genomes == 0 || cfg->genome_begin >= cfg->genome_end || cfg->genome_end > cfg->n_genomes) {
        pk_set_error("bad genome shard [%u,%u) of %u", cfg->genome_begin, cfg->genome_end, cfg->n_genomes);
        return PK_EINVAL;
    }
    if (cfg->genome_begin % 8 || (cfg->genome_end % 8 && cfg->genome_end != cfg->n_genomes)) {
        pk_set_error("genome shard [%u,%u) must fall on byte (8-genome) boundaries", cfg->genome_begin, cfg->genome_end);
        return PK_EINVAL;
    }
    if (cfg->load_factor < 0.f || cfg->load_factor > 0.9f) { pk_set_error("load_factor %.3f out of (0, 0.9]", cfg->load_factor); return PK_EINVAL; }
    int ndev = pk_device_count();
    if (ndev <= 0) { pk_set_error("no CUDA device visible: libpkanchor has no CPU path"); return PK_ECUDA; }
    if (cfg->device < 0 || cfg->device >= ndev) { pk_set_error("device %d out of range (%d visible)", cfg->device, ndev); return PK_EINVAL; }
    std::unique_ptr<pk_engine> e(new pk_engine());
    e->cfg = *cfg;
    if (e->cfg.lowres_step == 0) e->cfg.lowres_step = 100;
    if (e->cfg.max_bin_len == 0) e->cfg.max_bin_len = 200000;
    if (e->cfg.min_bin_count == 0) e->cfg.min_bin_count = 100;
    if (e->cfg.load_factor == 0.f) e->cfg.load_factor = 0.5f;
    e->chunk = cfg->chunk_positions ? cfg->chunk_positions : (4u << 20);
    e->n_local = cfg->genome_end - cfg->genome_begin;
    e->row_bytes = (e->n_local + 7) / 8;
    e->tabs.resize(e->n_local);
    CU(cudaSetDevice(cfg->device));
    CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&e->d_tables, sizeof(PkTable) * e->n_local));
    CU(cudaMemset(e->d_tables, 0, sizeof(PkTable) * e->n_local));
    CU(cudaMalloc(&e->d_counters, sizeof(unsigned long long) * 3 * e->n_local));
    CU(cudaMemset(e->d_counters, 0, sizeof(unsigned long long) * 3 * e->n_local));
    CU(cudaMalloc(&e->d_colsums, sizeof(unsigned long long) * e->n_local));
    *out = e.release();
    return PK_OK;
}

static void free_slots(pk_engine *e) {
    for (auto &s : e->slots) {
        if (s.stream) cudaStreamDestroy(s.stream);
        cudaFree(s.d_ascii); cudaFree(s.d_words); cudaFree(s.d_mask); cudaFree(s.d_rows); cudaFree(s.d_low);
        for (auto &ev : s.ev) if (ev) cudaEventDestroy(ev);
        s = Slot();
    }
    e->slots_ready = false;
}

extern "C" void pk_engine_destroy(pk_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (auto &t : e->tabs) cudaFree(t.dev.slots);
    free_slots(e);
    cudaFree(e->d_tables); cudaFree(e->d_counters); cudaFree(e->d_colsums); cudaFree(e->d_hist);
    cudaFree(e->d_stage);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int upload_tables(pk_engine *e) {
    std::vector<PkTable> h(e->n_local);
    for (uint32_t i = 0; i < e->n_local; i++) h[i] = e->tabs[i].dev;
    CU(cudaMemcpyAsync(e->d_tables, h.data(), sizeof(PkTable) * e->n_local, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PK_OK;
}

extern "C" int pk_engine_reserve(pk_engine *e, uint32_t genome, uint64_t max_keys) {
    if (!e) { pk_set_error("null engine"); return PK_EINVAL; }
    if (genome >= e->cfg.n_genomes) { pk_set_error("genome %u out of range", genome); return PK_EINVAL; }
    if (!is_local(e, genome)) return PK_OK;
    int rc = set_device(e); if (rc) return rc;
    HostTable &t = e->tabs[genome - e->cfg.genome_begin];
    if (t.reserved) { pk_set_error("genome %u already reserved", genome); return PK_ESTATE; }
    // buckets of 4 slots; at least 2 so the walk-on probe terminates on an EMPTY slot
    uint64_t nb = (uint64_t)((double)max_keys / (4.0 * e->cfg.load_factor)) + 2;
    if (nb >= 0xFFFFFFFFull) { pk_set_error("genome %u: %llu keys exceed the 2^32-bucket table limit", genome, (unsigned long long)max_keys); return PK_EUNSUPPORTED; }
    CU(cudaMalloc(&t.dev.slots, nb * 32));
    t.dev.n_buckets = (uint32_t)nb;
    t.capacity = max_keys;
    t.reserved = true;
    pk_launch_fill_empty(t.dev.slots, nb * 4, e->stream);
    CU(cudaGetLastError());
    e->finalized = false;
    return upload_tables(e);
}

static int ensure_stage(pk_engine *e, size_t bytes) {
    if (e->stage_bytes >= bytes) return PK_OK;
    cudaFree(e->d_stage); e->d_stage = nullptr;
    if (e->h_stage) { cudaFreeHost(e->h_stage); e->h_stage = nullptr; }
    e->stage_bytes = 0;
    CU(cudaMalloc(&e->d_stage, bytes));
    CU(cudaMallocHost(&e->h_stage, bytes));
    e->stage_bytes = bytes;
    return PK_OK;
}

static int ingest_kmc(pk_engine *e, const char *prefix, bool bitvec, uint32_t genome_or_first) {
    pk_kmcdb *db = nullptr;
    int rc = pk_kmcdb_open_impl(prefix, &db);
    if (rc) return rc;
    struct Closer { pk_kmcdb *d; ~Closer() { pk_kmcdb_close_impl(d); } } closer{db};
    const pk_kmcdb_info &I = db->info;
    if (I.kmer_length != e->cfg.k) { pk_set_error("%s: k=%u but the engine was created with k=%u", prefix, I.kmer_length, e->cfg.k); return PK_EINVAL; }
    if (!I.both_strands) { pk_set_error("%s: non-canonical database (counted with -b)", prefix); return PK_EUNSUPPORTED; }
    rc = set_device(e); if (rc) return rc;
    if (!bitvec) {
        if (!is_local(e, genome_or_first)) return PK_OK;
        if (!e->tabs[genome_or_first - e->cfg.genome_begin].reserved) {
            rc = pk_engine_reserve(e, genome_or_first, I.total_kmers); if (rc) return rc;
        }
    } else {
        for (uint32_t j = 0; j < 32 && genome_or_first + j < e->cfg.n_genomes; j++) {
            const uint32_t g = genome_or_first + j;
            if (is_local(e, g) && !e->tabs[g - e->cfg.genome_begin].reserved) {
                // upper bound: every k-mer of the union could belong to this genome
                rc = pk_engine_reserve(e, g, I.total_kmers); if (rc) return rc;
            }
        }
    }
    uint64_t *d_lut = nullptr;
    CU(cudaMalloc(&d_lut, db->lut.size() * 8));
    struct Freer { void *p; ~Freer() { cudaFree(p); } } freer{d_lut};
    CU(cudaMemcpyAsync(d_lut, db->lut.data(), db->lut.size() * 8, cudaMemcpyHostToDevice, e->stream));
    const uint64_t recs_per_chunk = std::max<uint64_t>(1, (32ull << 20) / std::max<uint32_t>(1, db->rec_size));
    rc = ensure_stage(e, recs_per_chunk * std::max<uint32_t>(1, db->rec_size)); if (rc) return rc;
    PkDecodeArgs a{};
    a.d_lut = d_lut; a.n_lut_slots = db->lut.size() - 1; a.single_lut = db->single_lut;
    a.suf_size = db->suf_size; a.counter_size = I.counter_size; a.rec_size = db->rec_size;
    a.min_count = I.min_count; a.max_count = I.max_count;
    a.bitvec = bitvec; a.first_genome = genome_or_first; a.gbegin = e->cfg.genome_begin; a.gend = e->cfg.genome_end;
    a.d_tables = e->d_tables; a.local_genome = bitvec ? 0 : genome_or_first - e->cfg.genome_begin;
    a.d_counters = e->d_counters;
    for (uint64_t r0 = 0; r0 < I.total_kmers; r0 += recs_per_chunk) {
        const uint64_t n = std::min(recs_per_chunk, I.total_kmers - r0);
        if (db->rec_size) {
            CU(cudaStreamSynchronize(e->stream));   // staging buffer reuse
            if (!pk_kmcdb_read_records(db, r0, n, e->h_stage)) { pk_set_error("%s.kmc_suf: read error", prefix); return PK_EIO; }
            CU(cudaMemcpyAsync(e->d_stage, e->h_stage, n * db->rec_size, cudaMemcpyHostToDevice, e->stream));
        }
        a.d_recs = e->d_stage; a.rec0 = r0; a.n = n;
        pk_launch_decode_insert(a, e->stream);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(e->stream));
    e->finalized = false;
    return PK_OK;
}

extern "C" int pk_engine_add_kmc(pk_engine *e, uint32_t genome, const char *kmc_prefix) {
    if (!e || !kmc_prefix) { pk_set_error("null argument"); return PK_EINVAL; }
    if (genome >= e->cfg.n_genomes) { pk_set_error("genome %u out of range", genome); return PK_EINVAL; }
    return ingest_kmc(e, kmc_prefix, false, genome);
}
extern "C" int pk_engine_add_bitvec(pk_engine *e, uint32_t first_genome, const char *kmc_prefix) {
    if (!e || !kmc_prefix) { pk_set_error("null argument"); return PK_EINVAL; }
    if (first_genome >= e->cfg.n_genomes || first_genome % 32) { pk_set_error("first_genome %u must be a multiple of 32 below N", first_genome); return PK_EINVAL; }
    return ingest_kmc(e, kmc_prefix, true, first_genome);
}

extern "C" int pk_engine_add_keys(pk_engine *e, uint32_t genome, const uint64_t *keys, uint64_t n) {
    if (!e || (!keys && n)) { pk_set_error("null argument"); return PK_EINVAL; }
    if (genome >= e->cfg.n_genomes) { pk_set_error("genome %u out of range", genome); return PK_EINVAL; }
    if (!is_local(e, genome)) return PK_OK;
    int rc = set_device(e); if (rc) return rc;
    HostTable &t = e->tabs[genome - e->cfg.genome_begin];
    if (!t.reserved) { rc = pk_engine_reserve(e, genome, n); if (rc) return rc; }
    const uint64_t kmask = e->cfg.k == 32 ? ~0ull : ((1ull << (2 * e->cfg.k)) - 1);
    for (uint64_t i = 0; i < n; i++)
        if (keys[i] & ~kmask) { pk_set_error("key %llu has bits above 2k", (unsigned long long)i); return PK_EINVAL; }
    uint64_t *d_keys = nullptr;
    if (n) {
        CU(cudaMalloc(&d_keys, n * 8));
        struct Freer { void *p; ~Freer() { cudaFree(p); } } freer{d_keys};
        CU(cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, e->stream));
        pk_launch_insert_keys(d_keys, n, t.dev, e->d_counters + 3 * (genome - e->cfg.genome_begin), e->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(e->stream));
    }
    e->finalized = false;
    return PK_OK;
}

static int add_sequence_device(pk_engine *e, uint32_t genome, const uint8_t *d_ascii, uint64_t len) {
    HostTable &t = e->tabs[genome - e->cfg.genome_begin];
    if (!t.reserved) { pk_set_error("genome %u: pk_engine_reserve must precede add_sequence", genome); return PK_ESTATE; }
    if (len < e->cfg.k) return PK_OK;
    const uint64_t nw = pk_packed_words(len);
    uint64_t *d_words = nullptr; uint32_t *d_mask = nullptr;
    CU(cudaMalloc(&d_words, nw * 8));
    struct Freer { void *p; ~Freer() { cudaFree(p); } } f1{d_words};
    CU(cudaMalloc(&d_mask, nw * 4));
    Freer f2{d_mask};
    pk_launch_pack(d_ascii, len, nw, d_words, d_mask, e->stream);
    pk_launch_insert_seq(d_words, d_mask, len - e->cfg.k + 1, e->cfg.k, t.dev,
                         e->d_counters + 3 * (genome - e->cfg.genome_begin), e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    e->finalized = false;
    return PK_OK;
}

extern "C" int pk_engine_add_sequence(pk_engine *e, uint32_t genome, const char *ascii, uint64_t len) {
    if (!e || (!ascii && len)) { pk_set_error("null argument"); return PK_EINVAL; }
    if (genome >= e->cfg.n_genomes) { pk_set_error("genome %u out of range", genome); return PK_EINVAL; }
    if (!is_local(e, genome)) return PK_OK;
    int rc = set_device(e); if (rc) return rc;
    if (len < e->cfg.k) {
        if (!e->tabs[genome - e->cfg.genome_begin].reserved) { pk_set_error("genome %u: pk_engine_reserve must precede add_sequence", genome); return PK_ESTATE; }
        return PK_OK;
    }
    uint8_t *d_ascii = nullptr;
    CU(cudaMalloc(&d_ascii, len));
    struct Freer { void *p; ~Freer() { cudaFree(p); } } f{d_ascii};
    CU(cudaMemcpyAsync(d_ascii, ascii, len, cudaMemcpyHostToDevice, e->stream));
    return add_sequence_device(e, genome, d_ascii, len);
}

extern "C" int pk_engine_add_sequence_device(pk_engine *e, uint32_t genome, const void *d_ascii, uint64_t len) {
    if (!e || (!d_ascii && len)) { pk_set_error("null argument"); return PK_EINVAL; }
    if (genome >= e->cfg.n_genomes) { pk_set_error("genome %u out of range", genome); return PK_EINVAL; }
    if (!is_local(e, genome)) return PK_OK;
    int rc = set_device(e); if (rc) return rc;
    CU(cudaDeviceSynchronize());   // the caller's producer stream is unknown: order after everything
    return add_sequence_device(e, genome, (const uint8_t *)d_ascii, len);
}

extern "C" int pk_engine_finalize(pk_engine *e) {
    if (!e) { pk_set_error("null engine"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    for (uint32_t i = 0; i < e->n_local; i++) {
        if (!e->tabs[i].reserved) {       // a genome without k-mers still needs a (tiny) table
            rc = pk_engine_reserve(e, e->cfg.genome_begin + i, 0); if (rc) return rc;
        }
    }
    std::vector<unsigned long long> c(3 * e->n_local);
    CU(cudaMemcpyAsync(c.data(), e->d_counters, c.size() * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (uint32_t i = 0; i < e->n_local; i++) {
        e->tabs[i].n_keys = c[3 * i]; e->tabs[i].n_overflow = c[3 * i + 1];
        if (c[3 * i + 2]) {
            pk_set_error("genome %u: table full (%llu keys did not fit a table reserved for %llu)",
                         e->cfg.genome_begin + i, c[3 * i + 2], (unsigned long long)e->tabs[i].capacity);
            return PK_ENOMEM;
        }
    }
    rc = upload_tables(e); if (rc) return rc;
    e->finalized = true;
    return PK_OK;
}

extern "C" int pk_engine_table_stats(const pk_engine *e, uint32_t genome, pk_table_stats *out) {
    if (!e || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    if (!is_local(e, genome)) { pk_set_error("genome %u is not in this engine's shard", genome); return PK_EINVAL; }
    const HostTable &t = e->tabs[genome - e->cfg.genome_begin];
    out->n_keys = t.n_keys; out->n_buckets = t.dev.n_buckets; out->n_overflow = t.n_overflow;
    out->bytes = (uint64_t)t.dev.n_buckets * 32;
    return PK_OK;
}

// ------------------------------------------------------------------ pinned host memory
extern "C" int pk_host_alloc(void **out, size_t bytes) {
    if (!out) { pk_set_error("null argument"); return PK_EINVAL; }
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return PK_OK;
}
extern "C" int pk_host_free(void *p) {
    if (p) CU(cudaFreeHost(p));
    return PK_OK;
}

// ------------------------------------------------------------------ device-level building blocks
#define NEED_FINAL(e)                                                                              \
    do {                                                                                           \
        if (!(e)) { pk_set_error("null engine"); return PK_EINVAL; }                               \
        if (!(e)->finalized) { pk_set_error("pk_engine_finalize has not been called"); return PK_ESTATE; } \
    } while (0)

extern "C" int pk_pack_device(pk_engine *e, const void *d_ascii, uint64_t len, void *d_words, void *d_mask, void *stream) {
    if (!e || !d_words || !d_mask || (!d_ascii && len)) { pk_set_error("null argument"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    pk_launch_pack((const uint8_t *)d_ascii, len, pk_packed_words(len), (uint64_t *)d_words, (uint32_t *)d_mask,
                   stream ? (pk_stream_t)stream : e->stream);
    CU(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_probe_device(pk_engine *e, const void *d_words, const void *d_mask, uint64_t p0, uint64_t n,
                               void *d_rows, uint32_t row_stride, uint32_t col_offset, void *stream) {
    NEED_FINAL(e);
    if (!d_words || !d_mask || !d_rows) { pk_set_error("null argument"); return PK_EINVAL; }
    if (col_offset + e->row_bytes > row_stride) { pk_set_error("row_stride %u too small for %u bytes at offset %u", row_stride, e->row_bytes, col_offset); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    pk_launch_probe((const uint64_t *)d_words, (const uint32_t *)d_mask, p0, n, e->cfg.k, e->d_tables, e->n_local,
                    (uint8_t *)d_rows, row_stride, col_offset, stream ? (pk_stream_t)stream : e->stream);
    CU(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_reduce_device(pk_engine *e, const void *d_rows, uint32_t row_stride, uint32_t n_cols, uint64_t p_first,
                                uint64_t n, uint64_t binlen, void *d_bin_hist, void *d_col_sums, void *d_rows_low,
                                uint32_t lowres_step, void *stream) {
    if (!e || !d_rows) { pk_set_error("null argument"); return PK_EINVAL; }
    if (n_cols == 0 || n_cols > 8 * row_stride || n_cols > 4096) { pk_set_error("bad n_cols %u for row_stride %u", n_cols, row_stride); return PK_EINVAL; }
    if (d_rows_low && !lowres_step) { pk_set_error("lowres_step must be > 0"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    pk_launch_reduce((const uint8_t *)d_rows, row_stride, n_cols, p_first, n, binlen, (unsigned long long *)d_bin_hist,
                     (unsigned long long *)d_col_sums, (uint8_t *)d_rows_low, lowres_step,
                     stream ? (pk_stream_t)stream : e->stream);
    CU(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_interleave_device(pk_engine *e, const void *d_planes, uint32_t n_ranks, uint64_t n, uint32_t w,
                                    void *d_rows, uint32_t row_stride, void *stream) {
    if (!e || !d_planes || !d_rows) { pk_set_error("null argument"); return PK_EINVAL; }
    if ((uint64_t)n_ranks * w > row_stride) { pk_set_error("row_stride %u < n_ranks*w", row_stride); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    pk_launch_interleave((const uint8_t *)d_planes, n_ranks, n, w, (uint8_t *)d_rows, row_stride,
                         stream ? (pk_stream_t)stream : e->stream);
    CU(cudaGetLastError());
    return PK_OK;
}

// ------------------------------------------------------------------ host-level hot path
static int ensure_slots(pk_engine *e) {
    if (e->slots_ready) return PK_OK;
    const uint64_t C = e->chunk, k = e->cfg.k;
    const uint64_t nw = pk_packed_words(C + k);
    const uint64_t low_rows = C / e->cfg.lowres_step + 2;
    for (auto &s : e->slots) {
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&s.d_ascii, C + k + 32));
        CU(cudaMalloc(&s.d_words, nw * 8));
        CU(cudaMalloc(&s.d_mask, nw * 4));
        CU(cudaMalloc(&s.d_rows, C * e->row_bytes));
        CU(cudaMalloc(&s.d_low, low_rows * e->row_bytes));
        for (auto &ev : s.ev) CU(cudaEventCreate(&ev));
    }
    e->slots_ready = true;
    return PK_OK;
}

struct ChunkTimes { int slot; };

extern "C" int pk_anchor_chrom(pk_engine *e, const char *ascii, uint64_t len, uint8_t *bitmap1, uint8_t *bitmap_low,
                               uint64_t *bin_hist, uint64_t *col_sums, uint64_t *nkmers_out) {
    NEED_FINAL(e);
    if (!ascii && len) { pk_set_error("null sequence"); return PK_EINVAL; }
    const uint32_t k = e->cfg.k, step = e->cfg.lowres_step, rb = e->row_bytes;
    if (nkmers_out) *nkmers_out = 0;
    memset(&e->stats, 0, sizeof e->stats);
    if (len < k) return PK_OK;                      // GetCountersForRead: clears and returns false (kmc_file.cpp:878-882)
    const uint64_t nk = len - k + 1;
    const uint64_t binlen = pk_bin_len(&e->cfg, nk);
    if (bin_hist && binlen == 0) {
        pk_set_error("chromosome with %llu k-mers (< min_bin_count %u) has no defined bins", (unsigned long long)nk, e->cfg.min_bin_count);
        return PK_EINVAL;
    }
    int rc = set_device(e); if (rc) return rc;
    rc = ensure_slots(e); if (rc) return rc;
    const uint64_t nbins = binlen ? (nk + binlen - 1) / binlen : 0;
    const uint64_t hist_words = nbins * (e->n_local + 1);
    if (bin_hist) {
        if (e->hist_cap < hist_words) {
            cudaFree(e->d_hist); e->d_hist = nullptr; e->hist_cap = 0;
            CU(cudaMalloc(&e->d_hist, hist_words * 8));
            e->hist_cap = hist_words;
        }
        CU(cudaMemsetAsync(e->d_hist, 0, hist_words * 8, e->stream));
    }
    if (col_sums) CU(cudaMemsetAsync(e->d_colsums, 0, e->n_local * 8, e->stream));
    cudaEvent_t ev_begin, ev_ready, ev_end;
    CU(cudaEventCreate(&ev_begin)); CU(cudaEventCreate(&ev_ready)); CU(cudaEventCreate(&ev_end));
    CU(cudaEventRecord(ev_begin, e->stream));
    CU(cudaEventRecord(ev_ready, e->stream));
    const uint64_t C = e->chunk;
    const uint64_t nchunks = (nk + C - 1) / C;
    std::vector<float> t_h2d, t_pack, t_probe, t_red, t_d2h;
    auto harvest = [&](Slot &s) -> int {      // the slot's previous chunk has completed: collect its timings
        float ms;
        CU(cudaEventElapsedTime(&ms, s.ev[0], s.ev[1])); e->stats.h2d_ms += ms;
        CU(cudaEventElapsedTime(&ms, s.ev[1], s.ev[2])); e->stats.pack_ms += ms;
        CU(cudaEventElapsedTime(&ms, s.ev[2], s.ev[3])); e->stats.probe_ms += ms;
        CU(cudaEventElapsedTime(&ms, s.ev[3], s.ev[4])); e->stats.reduce_ms += ms;
        CU(cudaEventElapsedTime(&ms, s.ev[4], s.ev[5])); e->stats.d2h_ms += ms;
        return PK_OK;
    };
    for (uint64_t c = 0; c < nchunks; c++) {
        Slot &s = e->slots[c % pk_engine::kSlots];
        const uint64_t p0 = c * C, n = std::min(C, nk - p0);
        if (c >= (uint64_t)pk_engine::kSlots) {
            CU(cudaStreamSynchronize(s.stream));
            rc = harvest(s); if (rc) return rc;
        } else {
            CU(cudaStreamWaitEvent(s.stream, ev_ready, 0));   // hist/colsum memsets
        }
        const uint64_t nbytes_in = n + k - 1;
        CU(cudaEventRecord(s.ev[0], s.stream));
        CU(cudaMemcpyAsync(s.d_ascii, ascii + p0, nbytes_in, cudaMemcpyHostToDevice, s.stream));
        CU(cudaEventRecord(s.ev[1], s.stream));
        pk_launch_pack(s.d_ascii, nbytes_in, pk_packed_words(nbytes_in), s.d_words, s.d_mask, s.stream);
        CU(cudaEventRecord(s.ev[2], s.stream));
        pk_launch_probe(s.d_words, s.d_mask, 0, n, k, e->d_tables, e->n_local, s.d_rows, rb, 0, s.stream);
        CU(cudaEventRecord(s.ev[3], s.stream));
        const bool want_low = bitmap_low != nullptr;
        if (bin_hist || col_sums || want_low)
            pk_launch_reduce(s.d_rows, rb, e->n_local, p0, n, bin_hist ? binlen : 0, bin_hist ? e->d_hist : nullptr,
                             col_sums ? e->d_colsums : nullptr, want_low ? s.d_low : nullptr, step, s.stream);
        CU(cudaEventRecord(s.ev[4], s.stream));
        if (bitmap1) CU(cudaMemcpyAsync(bitmap1 + p0 * rb, s.d_rows, n * rb, cudaMemcpyDeviceToHost, s.stream));
        if (want_low) {
            const uint64_t l0 = (p0 + step - 1) / step, l1 = (p0 + n + step - 1) / step;   // low rows in [l0, l1)
            if (l1 > l0) CU(cudaMemcpyAsync(bitmap_low + l0 * rb, s.d_low, (l1 - l0) * rb, cudaMemcpyDeviceToHost, s.stream));
        }
        CU(cudaEventRecord(s.ev[5], s.stream));
        CU(cudaGetLastError());
        e->stats.probe_launches += 1;
        e->stats.kernel_launches += 2 + ((bin_hist || col_sums || want_low) ? 1 : 0);
    }
    for (uint64_t c = 0; c < std::min<uint64_t>(nchunks, pk_engine::kSlots); c++) {
        Slot &s = e->slots[(nchunks - 1 - c) % pk_engine::kSlots];
        CU(cudaStreamSynchronize(s.stream));
        rc = harvest(s); if (rc) return rc;
    }
    std::vector<unsigned long long> tmp;
    if (bin_hist) {
        CU(cudaMemcpyAsync(bin_hist, e->d_hist, hist_words * 8, cudaMemcpyDeviceToHost, e->stream));
    }
    if (col_sums) {
        tmp.resize(e->n_local);
        CU(cudaMemcpyAsync(tmp.data(), e->d_colsums, e->n_local * 8, cudaMemcpyDeviceToHost, e->stream));
    }
    CU(cudaEventRecord(ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (col_sums) for (uint32_t g = 0; g < e->n_local; g++) col_sums[g] += tmp[g];
    CU(cudaEventElapsedTime(&e->stats.total_ms, ev_begin, ev_end));
    cudaEventDestroy(ev_begin); cudaEventDestroy(ev_ready); cudaEventDestroy(ev_end);
    e->stats.positions = nk;
    e->stats.probes = nk * e->n_local;
    if (nkmers_out) *nkmers_out = nk;
    return PK_OK;
}

extern "C" int pk_get_counters_for_read(pk_engine *e, uint32_t dbi, const char *read, uint64_t len, uint32_t *counters,
                                        uint64_t *n_out) {
    NEED_FINAL(e);
    if ((!read && len) || !counters || !n_out) { pk_set_error("null argument"); return PK_EINVAL; }
    if (dbi >= (e->cfg.n_genomes + 31) / 32) { pk_set_error("bitvec index %u out of range", dbi); return PK_EINVAL; }
    *n_out = 0;
    const uint32_t k = e->cfg.k, rb = e->row_bytes;
    if (len < k) return PK_OK;
    const uint64_t nk = len - k + 1;
    int rc = set_device(e); if (rc) return rc;
    // genomes [32*dbi, 32*dbi+32) intersected with the local shard
    const uint32_t g_lo = std::max(32 * dbi, e->cfg.genome_begin), g_hi = std::min({32 * dbi + 32, e->cfg.genome_end, e->cfg.n_genomes});
    if (g_lo >= g_hi) { memset(counters, 0, nk * 4); *n_out = nk; return PK_OK; }
    const uint32_t byte_off = (g_lo - e->cfg.genome_begin) / 8;          // shard starts on a byte boundary
    const uint32_t shift = g_lo - 32 * dbi;                              // multiple of 8
    const uint32_t nbits = g_hi - g_lo, nbytes = (nbits + 7) / 8;
    const uint32_t mask = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1);
    uint8_t *d_ascii = nullptr, *d_rows = nullptr; uint64_t *d_words = nullptr; uint32_t *d_mask = nullptr, *d_out = nullptr;
    const uint64_t nw = pk_packed_words(len);
    struct Freer { void *p = nullptr; ~Freer() { cudaFree(p); } } f[5];
    CU(cudaMalloc(&d_ascii, len)); f[0].p = d_ascii;
    CU(cudaMalloc(&d_words, nw * 8)); f[1].p = d_words;
    CU(cudaMalloc(&d_mask, nw * 4)); f[2].p = d_mask;
    CU(cudaMalloc(&d_rows, nk * rb)); f[3].p = d_rows;
    CU(cudaMalloc(&d_out, nk * 4)); f[4].p = d_out;
    cudaStream_t s = e->stream;
    CU(cudaMemcpyAsync(d_ascii, read, len, cudaMemcpyHostToDevice, s));
    pk_launch_pack(d_ascii, len, nw, d_words, d_mask, s);
    pk_launch_probe(d_words, d_mask, 0, nk, k, e->d_tables, e->n_local, d_rows, rb, 0, s);
    pk_launch_rows_to_u32(d_rows, rb, byte_off, nbytes, mask, nk, d_out, s);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counters, d_out, nk * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (shift) for (uint64_t i = 0; i < nk; i++) counters[i] <<= shift;
    *n_out = nk;
    return PK_OK;
}

extern "C" int pk_engine_stats(const pk_engine *e, pk_stats *out) {
    if (!e || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    *out = e->stats;
    return PK_OK;
}
