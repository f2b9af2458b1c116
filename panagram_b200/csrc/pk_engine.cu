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
#include "pk_gather.cuh"

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
    bool sealed = false;        // its keys live in the group table only (group_only): no inserts, no per-genome lookups
};

struct pk_engine {
    pk_config cfg{};
    PkKeySpec ks{};                             // k + slot format of the tables
    uint32_t n_local = 0, row_bytes = 0;
    std::vector<HostTable> tabs;
    PkTable *d_tables = nullptr;
    std::vector<PkTable> h_tables;              // host copy of the descriptors (kernel-parameter path)
    // group tables: one per PK_GROUP (8) local genomes, derived from the per-genome tables by finalize; the
    // partitioned probe answers 8 genomes per probe out of them (pk_device.cuh)
    std::vector<HostTable> utabs;
    std::vector<PkTable> h_utables;             // empty when group tables are off or could not be built
    PkTable *d_utables = nullptr;               // the same descriptors on the device (direct probe, spill)
    int union_tables = 1;
    int group_only = 0;                         // free the per-genome tables once their group table is built
    int allow_g32 = 1;                          // 32-bit group slots where k and the table size allow (pk_device.cuh, G32)
    int gfmt = -1;                              // format of this engine's group tables: -1 undecided, 0 G64, 1 G32
    unsigned long long *d_ucounters = nullptr;  // [4]
    unsigned long long *d_counters = nullptr;   // [3 * n_local]
    bool finalized = false;
    cudaStream_t stream = nullptr;              // build stream / default stream for device-level calls
    cudaStream_t copy_stream = nullptr;         // D2H of finished chromosomes
    cudaStream_t in_stream = nullptr;           // H2D of the anchor's chromosomes
    cudaEvent_t ev[6] = {};                     // begin, h2d done, pack done, probe done, reduce done, end
    // genome-batch buffers (grow-only)
    uint8_t *g_ascii = nullptr; uint64_t g_ascii_cap = 0;
    uint64_t *g_words = nullptr; uint64_t g_words_cap = 0;
    uint32_t *g_mask = nullptr; uint64_t g_mask_cap = 0;
    uint8_t *g_rows = nullptr; uint64_t g_rows_cap = 0;
    uint8_t *g_low = nullptr; uint64_t g_low_cap = 0;
    uint32_t *g_u32 = nullptr; uint64_t g_u32_cap = 0;
    uint64_t last_lowtot = 0;                   // low-res rows the last anchoring call left in g_low
    unsigned long long *d_hist = nullptr;       // per-chromosome bin histograms
    uint64_t hist_cap = 0;
    cudaEvent_t pev[6] = {};                    // around K1 / K2 / K3 / spill / K4 of the last partitioned launch
    int unpermute = 1;
    int e2e_batches = 2;                        // batches of whole chromosomes per pk_anchor_genome call (copy/compute overlap)
    uint64_t e2e_batch_min = 32ull << 20;       // ... for genomes of at least this many positions
    int e2e_batch_force = 0;                    // set by an explicit "e2e_batches" knob: skip the table-size rule
    int e2e_front_small = 0;                    // even splits: the straddling chromosome goes to the later batch
    bool pev_valid = false;
    PkPartTune tune{};                          // tuning state of the partitioned probe (pk_engine_tune / PK_K3* environment)
    PkPartScratch sc{};                         // partitioned-probe scratch (grow-only)
    PkPartPlan sc_plan{};
    int l2_prefetch = 1;
    int gather_dst_mode = 1;                    // narrow rows: output-aligned 16-byte stores in the exchange kernel (pk_gather.cuh)
    unsigned long long *d_colsums = nullptr;    // [n_local]
    // staging for KMC ingestion
    uint8_t *h_stage = nullptr, *d_stage = nullptr;
    size_t stage_bytes = 0;
    // on-GPU BGZF writer: CRC tables, scratch, the contiguous row stream and the two file images (grow-only)
    uint32_t *z_tables = nullptr;
    uint8_t *z_scratch = nullptr; uint64_t z_scratch_cap = 0;
    uint8_t *z_cat = nullptr; uint64_t z_cat_cap = 0;
    uint8_t *z_gz[2] = {nullptr, nullptr}; uint64_t z_gz_cap[2] = {0, 0};
    unsigned long long *z_gzi[2] = {nullptr, nullptr}; uint64_t z_gzi_cap[2] = {0, 0};
    unsigned long long *z_totals = nullptr;     // [4]
    // position-split exchange: the last segment list, on the device (pk_gather_slice_device)
    std::vector<PkgSeg> seg_host;
    PkgSeg *d_segs = nullptr; uint64_t d_segs_cap = 0;
    uint32_t seg_w = 0;
    uint64_t seg_chunks = 0;
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
    if (e->cfg.probe_mode > 2) { pk_set_error("probe_mode %u out of range", e->cfg.probe_mode); return PK_EINVAL; }
    // slot format: S32 (8 x 32-bit slots per sector) whenever the 2k-29 high key bits fit the bucket index
    e->ks.k = cfg->k;
    e->ks.eb = 2 * cfg->k > PK_S32_REM_BITS ? 2 * cfg->k - PK_S32_REM_BITS : 0;
    e->ks.fmt = e->ks.eb <= PK_S32_MAX_EB ? PK_FMT_S32 : PK_FMT_S64;
    if (const char *tf = getenv("PK_TABLE_FMT")) { if (atoi(tf) == 64) e->ks.fmt = PK_FMT_S64; }
    if (const char *pf = getenv("PK_L2_PREFETCH")) e->l2_prefetch = atoi(pf);
    if (const char *up = getenv("PK_UNPERMUTE")) e->unpermute = atoi(up);
    if (const char *ut = getenv("PK_GROUP_TABLES")) e->union_tables = atoi(ut) ? 1 : 0;
    if (const char *g3 = getenv("PK_GROUP_G32")) e->allow_g32 = atoi(g3) ? 1 : 0;
    if (const char *kv = getenv("PK_K3_VARIANT")) { const int v = atoi(kv); if (v >= -1 && v < pk_part_n_variants()) e->tune.variant = v; }
    {
        const char *we = getenv("PK_K3_WINDOW"), *wv = getenv("PK_K3W_VARIANT"), *ws = getenv("PK_K3W_GROUP");
        if (we) e->tune.window = atoi(we);
        if (wv && atoi(wv) >= -1 && atoi(wv) < pk_part_n_wvariants()) e->tune.wvariant = atoi(wv);
        if (ws && (atoi(ws) == 0 || atoi(ws) == 1 || atoi(ws) == 2 || atoi(ws) == 4)) e->tune.wgroup = atoi(ws);
    }
    e->sc.tune = &e->tune; e->sc.last_window = &e->tune.last_window;
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
    for (auto &ev : e->pev) CU(cudaEventCreate(&ev));
    CU(cudaMalloc(&e->ks.stash, sizeof(unsigned long long) * PK_STASH_SLOTS));
    CU(cudaMemset(e->ks.stash, 0xFF, sizeof(unsigned long long) * PK_STASH_SLOTS));
    CU(cudaMalloc(&e->ks.stash_n, sizeof(unsigned int)));
    CU(cudaMemset(e->ks.stash_n, 0, sizeof(unsigned int)));
    *out = e.release();
    return PK_OK;
}

extern "C" void pk_engine_destroy(pk_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (auto &t : e->tabs) cudaFree(t.dev.slots);
    for (auto &t : e->utabs) cudaFree(t.dev.slots);
    cudaFree(e->d_ucounters); cudaFree(e->d_utables);
    cudaFree(e->sc.buf1); cudaFree(e->sc.buf2); cudaFree(e->sc.spill);
    cudaFree(e->sc.cursor1); cudaFree(e->sc.cursor2); cudaFree(e->sc.spill_cursor); cudaFree(e->sc.err);
    cudaFree(e->g_ascii); cudaFree(e->g_words); cudaFree(e->g_mask); cudaFree(e->g_rows); cudaFree(e->g_low); cudaFree(e->g_u32);
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->pev) if (ev) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->in_stream) cudaStreamDestroy(e->in_stream);
    cudaFree(e->ks.stash); cudaFree(e->ks.stash_n);
    cudaFree(e->d_tables); cudaFree(e->d_counters); cudaFree(e->d_colsums); cudaFree(e->d_hist);
    cudaFree(e->d_stage);
    cudaFree(e->z_tables); cudaFree(e->z_scratch); cudaFree(e->z_cat); cudaFree(e->z_totals);
    for (int i = 0; i < 2; i++) { cudaFree(e->z_gz[i]); cudaFree(e->z_gzi[i]); }
    cudaFree(e->sc.out_list); cudaFree(e->sc.out_cursor);
    cudaFree(e->d_segs);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int upload_tables(pk_engine *e) {
    std::vector<PkTable> &h = e->h_tables;
    h.resize(e->n_local);
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
    if (t.sealed) { pk_set_error("genome %u is sealed in its group table", genome); return PK_ESTATE; }
    // 32-byte buckets of 4 (S64) or 8 (S32) slots; S32 needs n_buckets >= 2^eb so that k-mers with equal
    // low bits never share a home bucket (pk_key_hash), and always a free slot somewhere for walk-ons to stop
    const double slots_per_bucket = e->ks.fmt == PK_FMT_S32 ? 8.0 : 4.0;
    uint64_t nb = (uint64_t)((double)max_keys / (slots_per_bucket * e->cfg.load_factor)) + 2;
    if (e->ks.fmt == PK_FMT_S32) nb = std::max<uint64_t>(nb, std::max<uint64_t>(1ull << e->ks.eb, 16));
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

// a genome's table is about to change: not allowed once sealed; a group table already built from it is dropped
// (rebuilt by the next finalize)
static int touch_genome(pk_engine *e, uint32_t g_local) {
    if (e->tabs[g_local].sealed) { pk_set_error("genome %u is sealed in its group table (group_only): no more k-mers can be added", e->cfg.genome_begin + g_local); return PK_ESTATE; }
    const uint32_t gi = g_local / 8;
    if (gi < e->utabs.size() && e->utabs[gi].reserved) {
        cudaFree(e->utabs[gi].dev.slots);
        e->utabs[gi] = HostTable{};
        e->h_utables.clear();
        bool any = false;
        for (auto &u : e->utabs) any = any || u.reserved;
        if (!any) e->gfmt = -1;
    }
    e->finalized = false;
    return PK_OK;
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
        rc = touch_genome(e, genome_or_first - e->cfg.genome_begin); if (rc) return rc;
        if (!e->tabs[genome_or_first - e->cfg.genome_begin].reserved) {
            rc = pk_engine_reserve(e, genome_or_first, I.total_kmers); if (rc) return rc;
        }
    }
    const uint64_t recs_per_chunk = std::max<uint64_t>(1, (32ull << 20) / std::max<uint32_t>(1, db->rec_size));
    rc = ensure_stage(e, recs_per_chunk * std::max<uint32_t>(1, db->rec_size)); if (rc) return rc;
    if (bitvec) {
        // size every genome's table from ITS k-mer count, not from the union's (up to 32x too much): one pass over
        // the counters, popcount per bit
        unsigned long long *d_bits = nullptr;
        CU(cudaMalloc(&d_bits, 32 * sizeof(unsigned long long)));
        struct Freer { void *p; ~Freer() { cudaFree(p); } } fb{d_bits};
        CU(cudaMemsetAsync(d_bits, 0, 32 * sizeof(unsigned long long), e->stream));
        for (uint64_t r0 = 0; r0 < I.total_kmers && db->rec_size; r0 += recs_per_chunk) {
            const uint64_t n = std::min(recs_per_chunk, I.total_kmers - r0);
            CU(cudaStreamSynchronize(e->stream));   // staging buffer reuse
            if (!pk_kmcdb_read_records(db, r0, n, e->h_stage)) { pk_set_error("%s.kmc_suf: read error", prefix); return PK_EIO; }
            CU(cudaMemcpyAsync(e->d_stage, e->h_stage, n * db->rec_size, cudaMemcpyHostToDevice, e->stream));
            pk_launch_count_bits(e->d_stage, n, db->rec_size, db->suf_size, I.counter_size, I.min_count, I.max_count, d_bits, e->stream);
            CU(cudaGetLastError());
        }
        unsigned long long bits[32];
        CU(cudaMemcpyAsync(bits, d_bits, sizeof bits, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        for (uint32_t j = 0; j < 32 && genome_or_first + j < e->cfg.n_genomes; j++) {
            const uint32_t g = genome_or_first + j;
            if (!is_local(e, g)) continue;
            rc = touch_genome(e, g - e->cfg.genome_begin); if (rc) return rc;
            if (!e->tabs[g - e->cfg.genome_begin].reserved) {
                rc = pk_engine_reserve(e, g, I.counter_size ? bits[j] : I.total_kmers); if (rc) return rc;
            }
        }
    }
    uint64_t *d_lut = nullptr;
    CU(cudaMalloc(&d_lut, db->lut.size() * 8));
    struct Freer { void *p; ~Freer() { cudaFree(p); } } freer{d_lut};
    CU(cudaMemcpyAsync(d_lut, db->lut.data(), db->lut.size() * 8, cudaMemcpyHostToDevice, e->stream));
    PkDecodeArgs a{};
    a.d_lut = d_lut; a.n_lut_slots = db->lut.size() - 1; a.single_lut = db->single_lut;
    a.suf_size = db->suf_size; a.counter_size = I.counter_size; a.rec_size = db->rec_size;
    a.ks = e->ks;
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
    rc = touch_genome(e, genome - e->cfg.genome_begin); if (rc) return rc;
    HostTable &t = e->tabs[genome - e->cfg.genome_begin];
    if (!t.reserved) { rc = pk_engine_reserve(e, genome, n); if (rc) return rc; }
    const uint32_t k = e->cfg.k;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t x = keys[i];
        if (x & ~kmask) { pk_set_error("key %llu has bits above 2k", (unsigned long long)i); return PK_EINVAL; }
        // canonical = min(x, revcomp(x)) (kmer_api.h:373-386); a non-canonical key could never be found by a probe,
        // and ~0 (T^32, whose canonical form is A^32 = 0) is the EMPTY slot
        uint64_t r = ~x, rc = 0;
        for (uint32_t b = 0; b < k; b++) { rc = (rc << 2) | (r & 3); r >>= 2; }
        if (rc < x) { pk_set_error("key %llu is not canonical (its reverse complement is smaller)", (unsigned long long)i); return PK_EINVAL; }
    }
    uint64_t *d_keys = nullptr;
    if (n) {
        CU(cudaMalloc(&d_keys, n * 8));
        struct Freer { void *p; ~Freer() { cudaFree(p); } } freer{d_keys};
        CU(cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, e->stream));
        pk_launch_insert_keys(d_keys, n, e->ks, t.dev, genome - e->cfg.genome_begin,
                              e->d_counters + 3 * (genome - e->cfg.genome_begin), e->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(e->stream));
    }
    e->finalized = false;
    return PK_OK;
}

static int add_sequence_device(pk_engine *e, uint32_t genome, const uint8_t *d_ascii, uint64_t len) {
    int trc = touch_genome(e, genome - e->cfg.genome_begin); if (trc) return trc;
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
    pk_launch_insert_seq(d_words, d_mask, len - e->cfg.k + 1, e->ks, t.dev, genome - e->cfg.genome_begin,
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

// Group tables (pk_device.cuh): merge the per-genome tables of every 8 local genomes into one table whose slots
// carry an 8-bit membership mask. The number of distinct k-mers of a group is not known in advance (between the
// largest genome's count and the sum): it is estimated by merging 1/64 of the hash range into a scratch table
// first. A group that cannot be built (allocation failure, a neighbourhood of 15 full buckets with 64-bit keys)
// switches group tables off for the engine: the per-genome tables answer every query on their own.
//
// group_only (pk_engine_tune "group_only" 1): the per-genome tables of a group are FREED once its group table
// holds their keys — they are build intermediates then, and every lookup (partitioned probe, direct probe, spill,
// pk_get_counters_for_read) is answered from the group tables. 64 genomes of configs[3] (k = 31) need 154 GB as
// per-genome tables + ~110 GB as group tables; sealed group by group (pk_engine_seal_group) the peak is the
// group tables + ONE group's per-genome tables.
#define PK_GROUP 8u
static void drop_group_tables(pk_engine *e) {
    for (auto &t : e->utabs) cudaFree(t.dev.slots);
    e->utabs.clear();
    e->h_utables.clear();
    e->gfmt = -1;
}
// the key spec of a probe launch: positions are hashed for the tables the launch probes
static PkKeySpec probe_ks(const pk_engine *e) {
    PkKeySpec ks = e->ks;
    ks.ghash = (!e->h_utables.empty() && e->gfmt == 1) ? 20u : 0u;
    return ks;
}
static int refresh_counts(pk_engine *e) {
    std::vector<unsigned long long> c(3 * e->n_local);
    CU(cudaMemcpyAsync(c.data(), e->d_counters, c.size() * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (uint32_t i = 0; i < e->n_local; i++) {
        if (e->tabs[i].sealed) continue;
        e->tabs[i].n_keys = c[3 * i]; e->tabs[i].n_overflow = c[3 * i + 1];
        if (c[3 * i + 2]) {
            pk_set_error("genome %u: table too small (%llu keys did not fit a table reserved for %llu)",
                         e->cfg.genome_begin + i, c[3 * i + 2], (unsigned long long)e->tabs[i].capacity);
            return PK_ENOMEM;
        }
    }
    return PK_OK;
}
// 1 = built, 0 = could not be built (a note went to stderr; the caller decides), < 0 = error
static int build_group(pk_engine *e, uint32_t gi) {
    const uint32_t n_groups = (e->n_local + PK_GROUP - 1) / PK_GROUP;
    if (e->utabs.size() != n_groups) e->utabs.assign(n_groups, HostTable{});
    HostTable &u = e->utabs[gi];
    if (u.reserved) return 1;
    if (!e->d_ucounters) CU(cudaMalloc(&e->d_ucounters, 4 * sizeof(unsigned long long)));
    const double load = std::min(e->ks.fmt == PK_FMT_S32 ? 0.5 : 0.35, (double)e->cfg.load_factor);
    const uint32_t ebu = 2 * e->cfg.k > 52 ? 2 * e->cfg.k - 52 : 0;
    const int use_stash = e->ks.fmt == PK_FMT_S32;
    unsigned long long c[4] = {0, 0, 0, 0};
    uint64_t sum = 0, mx = 0;
    auto fail = [&](const char *why) {
        const cudaError_t ce = cudaGetLastError();
        fprintf(stderr, "[pkanchor] group table %u not built: %s (slots %llu, bits %llu, stashed %llu, failed %llu of %llu keys; cuda: %s)\n",
                gi, why, c[0], c[1], c[2], c[3], (unsigned long long)sum, cudaGetErrorString(ce));
        if (u.dev.slots) { cudaFree(u.dev.slots); u = HostTable{}; }
        return 0;
    };
    // the group tables are an acceleration structure: their fill is capped whatever the per-genome tables use. With 4
    // slots per bucket, 2 of 315 M keys met 15 full buckets in a row at a fill of 0.49 (configs[1]); k <= 24 sends
    // those to the stash, longer k-mers (no stash for > 48-bit keys) get a fill of 0.35 instead
    const uint32_t g0 = gi * PK_GROUP, ng = std::min(PK_GROUP, e->n_local - g0);
    for (uint32_t g = g0; g < g0 + ng; g++) { sum += e->tabs[g].n_keys; mx = std::max(mx, e->tabs[g].n_keys); }
    uint64_t distinct = sum;
    if (ng > 1 && sum >= (1ull << 22)) {
        // estimate: merge the first 1/64 of every table's buckets (= of the hash range) into a scratch table
        PkTable tmp{nullptr, 0, 0};
        const uint64_t nbt = (uint64_t)((double)sum / 64 / (4.0 * 0.5)) + 4096;
        if (cudaMalloc(&tmp.slots, nbt * 32) != cudaSuccess) return fail("scratch allocation");
        tmp.n_buckets = (uint32_t)nbt;
        pk_launch_fill_empty(tmp.slots, nbt * 4, e->stream);
        cudaMemsetAsync(e->d_ucounters, 0, sizeof c, e->stream);
        for (uint32_t g = g0; g < g0 + ng; g++)
            pk_launch_union_merge(e->tabs[g].dev, e->tabs[g].dev.n_buckets / 64, 6, e->ks, tmp, g - g0, g, 0, e->d_ucounters, e->stream);
        cudaMemcpyAsync(c, e->d_ucounters, sizeof c, cudaMemcpyDeviceToHost, e->stream);
        const cudaError_t er = cudaStreamSynchronize(e->stream);
        cudaFree(tmp.slots);
        if (er != cudaSuccess) return fail("estimate");
        distinct = std::min<uint64_t>(sum, std::max<uint64_t>(mx, (uint64_t)((double)(c[0] + c[3]) * 64 * 1.03) + 65536));
    }
    // slot format, one for all group tables of the engine (they share the hash of a launch): G32 (32-bit slots, 8 per
    // bucket, 2k - 20 key bits implied by the home bucket) when the first group's table has at least 2^(2k-20) buckets
    // at its natural size — any genome-scale table at k <= 23 — or k is so short that the minimum is tiny; later
    // (smaller) groups are padded to that minimum
    const uint32_t ebg = 2 * e->cfg.k > 20 ? 2 * e->cfg.k - 20 : 0;
    const uint64_t nb32 = (uint64_t)((double)distinct / (8.0 * std::min(0.5, (double)e->cfg.load_factor))) + 2;
    if (e->gfmt < 0)
        e->gfmt = (e->allow_g32 && use_stash && ebg <= 26 && (ebg <= 12 || nb32 >= (1ull << ebg) || e->allow_g32 == 2)) ? 1 : 0;
    uint64_t nb;
    if (e->gfmt == 1) nb = std::max<uint64_t>(nb32, std::max<uint64_t>(1ull << ebg, 16));
    else nb = std::max<uint64_t>((uint64_t)((double)distinct / (4.0 * load)) + 2, std::max<uint64_t>(1ull << ebu, 16));
    if (nb >= 0xFFFFFFFFull) return fail("too many buckets");
    if (cudaMalloc(&u.dev.slots, nb * 32) != cudaSuccess) { u.dev.slots = nullptr; return fail("allocation"); }
    u.dev.n_buckets = (uint32_t)nb;
    u.dev.fmt = (uint32_t)e->gfmt;
    u.capacity = distinct;
    pk_launch_fill_empty(u.dev.slots, nb * 4, e->stream);
    cudaMemsetAsync(e->d_ucounters, 0, sizeof c, e->stream);
    for (uint32_t g = g0; g < g0 + ng; g++)
        pk_launch_union_merge(e->tabs[g].dev, e->tabs[g].dev.n_buckets, 0, e->ks, u.dev, g - g0, g, use_stash, e->d_ucounters, e->stream);
    if (use_stash) pk_launch_union_merge_stash(e->ks, u.dev, g0, ng, e->d_ucounters, e->stream);
    cudaMemcpyAsync(c, e->d_ucounters, sizeof c, cudaMemcpyDeviceToHost, e->stream);
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return fail("merge");
    // every per-genome key must have arrived: bits set + stashed == sum of the genomes' key counts
    if (c[3] || c[1] + c[2] != sum) return fail("merge incomplete");
    u.n_keys = c[0];
    u.reserved = true;
    if (e->group_only) {
        for (uint32_t g = g0; g < g0 + ng; g++) {
            cudaFree(e->tabs[g].dev.slots);
            e->tabs[g].dev = PkTable{nullptr, 0, 0};
            e->tabs[g].sealed = true;
        }
    }
    return 1;
}

extern "C" int pk_engine_seal_group(pk_engine *e, uint32_t group) {
    if (!e) { pk_set_error("null engine"); return PK_EINVAL; }
    const uint32_t n_groups = (e->n_local + PK_GROUP - 1) / PK_GROUP;
    if (group >= n_groups) { pk_set_error("group %u out of range (%u groups of 8 local genomes)", group, n_groups); return PK_EINVAL; }
    if (!e->union_tables) { pk_set_error("group tables are switched off"); return PK_ESTATE; }
    int rc = set_device(e); if (rc) return rc;
    const uint32_t g0 = group * PK_GROUP, ng = std::min(PK_GROUP, e->n_local - g0);
    for (uint32_t g = g0; g < g0 + ng; g++)
        if (!e->tabs[g].reserved) { rc = pk_engine_reserve(e, e->cfg.genome_begin + g, 0); if (rc) return rc; }
    rc = refresh_counts(e); if (rc) return rc;
    rc = build_group(e, group);
    if (rc < 0) return rc;
    if (rc == 0) { pk_set_error("group table %u could not be built (see stderr)", group); return PK_ENOMEM; }
    e->finalized = false;
    return PK_OK;
}

static int build_group_tables(pk_engine *e) {
    const uint32_t n_groups = (e->n_local + PK_GROUP - 1) / PK_GROUP;
    e->h_utables.clear();
    if (!e->union_tables) { drop_group_tables(e); return PK_OK; }
    for (uint32_t gi = 0; gi < n_groups; gi++) {
        const int rc = build_group(e, gi);
        if (rc < 0) return rc;
        if (rc == 0) {
            if (e->group_only) { pk_set_error("group table %u could not be built and group_only is set (see stderr)", gi); return PK_ENOMEM; }
            drop_group_tables(e);          // not an error for the caller: the per-genome tables answer everything
            return PK_OK;
        }
    }
    e->h_utables.resize(n_groups);
    for (uint32_t gi = 0; gi < n_groups; gi++) e->h_utables[gi] = e->utabs[gi].dev;
    if (!e->d_utables) CU(cudaMalloc(&e->d_utables, sizeof(PkTable) * n_groups));
    CU(cudaMemcpy(e->d_utables, e->h_utables.data(), sizeof(PkTable) * n_groups, cudaMemcpyHostToDevice));
    return PK_OK;
}

extern "C" int pk_engine_finalize(pk_engine *e) {
    if (!e) { pk_set_error("null engine"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    for (uint32_t i = 0; i < e->n_local; i++) {
        if (!e->tabs[i].reserved) {       // a genome without k-mers still needs a (tiny) table
            rc = pk_engine_reserve(e, e->cfg.genome_begin + i, 0); if (rc) return rc;
        }
    }
    rc = refresh_counts(e); if (rc) return rc;
    rc = upload_tables(e); if (rc) return rc;
    rc = build_group_tables(e); if (rc) return rc;
    rc = upload_tables(e); if (rc) return rc;          // sealed genomes: null descriptors
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

extern "C" int pk_engine_group_stats(const pk_engine *e, uint32_t group, pk_table_stats *out) {
    if (!e || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    if (group >= e->utabs.size() || e->h_utables.empty()) { pk_set_error("no group table %u (group tables off, or not built)", group); return PK_ESTATE; }
    const HostTable &t = e->utabs[group];
    out->n_keys = t.n_keys; out->n_buckets = t.dev.n_buckets; out->n_overflow = 0;
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

// ------------------------------------------------------------------ k-mer sample for genome distances
extern "C" int pk_engine_sample_kmers(pk_engine *e, uint32_t hmax, uint64_t *keys, uint32_t *tags, uint64_t cap, uint64_t *n_out) {
    NEED_FINAL(e);
    if (!n_out || (cap && (!keys || !tags))) { pk_set_error("null argument"); return PK_EINVAL; }
    if (e->h_utables.empty()) { pk_set_error("k-mer sampling reads the group tables: they are switched off or could not be built"); return PK_ESTATE; }
    int rc = set_device(e); if (rc) return rc;
    unsigned long long *d_keys = nullptr, *d_n = nullptr;
    uint32_t *d_tags = nullptr;
    struct Freer { void *p; ~Freer() { cudaFree(p); } };
    CU(cudaMalloc(&d_keys, std::max<uint64_t>(cap, 1) * 8)); Freer f1{d_keys};
    CU(cudaMalloc(&d_tags, std::max<uint64_t>(cap, 1) * 4)); Freer f2{d_tags};
    CU(cudaMalloc(&d_n, 8)); Freer f3{d_n};
    CU(cudaMemsetAsync(d_n, 0, 8, e->stream));
    const int take_all = hmax == 0;
    for (uint32_t u = 0; u < e->h_utables.size(); u++)
        pk_launch_sample_group(e->h_utables[u], e->ks, u, hmax, take_all, d_keys, d_tags, cap, d_n, e->stream);
    pk_launch_sample_stash(e->ks, e->gfmt == 1, hmax, take_all, d_keys, d_tags, cap, d_n, e->stream);
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    *n_out = n;
    if (n > cap) { pk_set_error("sample holds %llu k-mers, the buffers %llu", n, (unsigned long long)cap); return PK_ENOMEM; }
    if (n) {
        CU(cudaMemcpy(keys, d_keys, n * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(tags, d_tags, n * 4, cudaMemcpyDeviceToHost));
    }
    return PK_OK;
}

// ------------------------------------------------------------------ peer memory (multi-GPU, process per GPU)
extern "C" int pk_device_alloc(pk_engine *e, void **d_ptr, size_t bytes) {
    if (!e || !d_ptr) { pk_set_error("null argument"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
    CU(cudaMemset(*d_ptr, 0, bytes ? bytes : 1));
    return PK_OK;
}
extern "C" int pk_device_free(pk_engine *e, void *d_ptr) {
    if (!e) { pk_set_error("null argument"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    if (d_ptr) CU(cudaFree(d_ptr));
    return PK_OK;
}
extern "C" int pk_ipc_export(pk_engine *e, const void *d_ptr, uint8_t handle[64]) {
    if (!e || !d_ptr || !handle) { pk_set_error("null argument"); return PK_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int rc = set_device(e); if (rc) return rc;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, 64);
    return PK_OK;
}
extern "C" int pk_ipc_open(pk_engine *e, const uint8_t handle[64], void **d_ptr) {
    if (!e || !d_ptr || !handle) { pk_set_error("null argument"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PK_OK;
}
extern "C" int pk_ipc_close(pk_engine *e, void *d_ptr) {
    if (!e) { pk_set_error("null argument"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr));
    return PK_OK;
}
extern "C" int pk_gather_interleave_device(pk_engine *e, const void *const *d_planes, uint32_t n_ranks, uint64_t n, uint32_t w,
                                           void *d_rows, uint32_t row_stride, void *stream) {
    if (!e || !d_planes || !d_rows) { pk_set_error("null argument"); return PK_EINVAL; }
    if (n_ranks == 0 || n_ranks > 16 || w == 0) { pk_set_error("n_ranks %u / w %u out of range (<= 16 ranks)", n_ranks, w); return PK_EINVAL; }
    if ((uint64_t)n_ranks * w > row_stride) { pk_set_error("row_stride %u < n_ranks*w", row_stride); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    if (pk_launch_gather_interleave(d_planes, n_ranks, n, w, (uint8_t *)d_rows, row_stride,
                                    stream ? (pk_stream_t)stream : e->stream)) {
        pk_set_error("gather_interleave launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PK_ECUDA;
    }
    return PK_OK;
}

extern "C" int pk_gather_slice_device(pk_engine *e, const void *const *d_planes, uint32_t n_ranks, uint64_t plane_rows, uint32_t w,
                                      const pk_segment *segs, uint32_t n_segs, void *d_rows, uint32_t row_stride, uint32_t row_bytes,
                                      void *stream) {
    if (!e || !d_planes || !d_rows || (!segs && n_segs)) { pk_set_error("null argument"); return PK_EINVAL; }
    if (n_ranks == 0 || n_ranks > PKG_MAX_RANKS || w == 0) { pk_set_error("n_ranks %u / w %u out of range (<= %d ranks)", n_ranks, w, PKG_MAX_RANKS); return PK_EINVAL; }
    if (row_bytes == 0 || row_bytes > row_stride || (uint64_t)(n_ranks - 1) * w >= row_bytes) {
        pk_set_error("row_bytes %u does not fit %u ranks of %u bytes in a stride of %u", row_bytes, n_ranks, w, row_stride);
        return PK_EINVAL;
    }
    for (uint32_t i = 0; i < n_segs; i++)
        if (segs[i].src_row + segs[i].n_rows > plane_rows) { pk_set_error("segment %u leaves the planes (%llu rows)", i, (unsigned long long)plane_rows); return PK_EINVAL; }
    if (!n_segs) return PK_OK;
    int rc = set_device(e); if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    bool same = e->d_segs && e->seg_w == w && e->seg_host.size() == n_segs;
    for (uint32_t i = 0; same && i < n_segs; i++)
        same = e->seg_host[i].src_row == segs[i].src_row && e->seg_host[i].n_rows == segs[i].n_rows && e->seg_host[i].dst_row == segs[i].dst_row;
    if (!same) {
        // a new list: rare (once per anchor and slice); the upload is synchronous so that seg_host may be rewritten
        e->seg_host.resize(n_segs);
        for (uint32_t i = 0; i < n_segs; i++) e->seg_host[i] = PkgSeg{segs[i].src_row, segs[i].n_rows, segs[i].dst_row, 0};
        e->seg_chunks = pkg_plan_segments(e->seg_host.data(), n_segs, w);
        e->seg_w = w;
        if (e->d_segs_cap < n_segs) {
            CU(cudaDeviceSynchronize());
            cudaFree(e->d_segs); e->d_segs = nullptr; e->d_segs_cap = 0;
            CU(cudaMalloc(&e->d_segs, sizeof(PkgSeg) * n_segs));
            e->d_segs_cap = n_segs;
        } else {
            CU(cudaDeviceSynchronize());        // an earlier launch may still read the old list
        }
        CU(cudaMemcpy(e->d_segs, e->seg_host.data(), sizeof(PkgSeg) * n_segs, cudaMemcpyHostToDevice));
    }
    uint64_t total_rows = 0;
    if (e->gather_dst_mode && pkg_dst_mode_ok(e->seg_host.data(), n_segs, n_ranks, w, row_stride, row_bytes, d_rows, &total_rows)) {
        if (pk_launch_gather_slice_dst(d_planes, n_ranks, plane_rows, w, e->d_segs, n_segs, total_rows, (uint8_t *)d_rows, s)) {
            pk_set_error("gather_slice launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return PK_ECUDA;
        }
        return PK_OK;
    }
    if (pk_launch_gather_slice(d_planes, n_ranks, plane_rows, w, e->d_segs, n_segs, e->seg_chunks, (uint8_t *)d_rows, row_stride, row_bytes, s)) {
        pk_set_error("gather_slice launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return PK_ECUDA;
    }
    return PK_OK;
}

// ------------------------------------------------------------------ probe dispatch
static void free_scratch(pk_engine *e) {
    cudaFree(e->sc.buf1); cudaFree(e->sc.buf2); cudaFree(e->sc.spill);
    cudaFree(e->sc.cursor1); cudaFree(e->sc.cursor2); cudaFree(e->sc.spill_cursor); cudaFree(e->sc.err);
    cudaFree(e->sc.out_list); cudaFree(e->sc.out_cursor);
    e->sc = PkPartScratch{};
    e->sc.tune = &e->tune; e->sc.last_window = &e->tune.last_window;
    e->sc_plan = PkPartPlan{};
}

static int ensure_scratch(pk_engine *e, const PkPartPlan &pl) {
    const PkPartPlan &have = e->sc_plan;
    const uint32_t n_groups = (e->n_local + 31) / 32;
    const uint64_t out_bytes = e->unpermute ? pl.out_bytes : 0;
    if (e->sc.buf1 && have.buf1_items >= pl.buf1_items && have.buf2_items >= pl.buf2_items &&
        have.spill_items >= pl.spill_items && have.n_regions1 >= pl.n_regions1 && have.n_regions2 >= pl.n_regions2 &&
        e->sc.out_bytes >= out_bytes)
        return PK_OK;
    CU(cudaDeviceSynchronize());
    free_scratch(e);
    CU(cudaMalloc(&e->sc.buf1, std::max<uint64_t>(pl.buf1_items, 1) * 8));
    CU(cudaMalloc(&e->sc.buf2, std::max<uint64_t>(pl.buf2_items, 1) * 8));
    CU(cudaMalloc(&e->sc.spill, std::max<uint64_t>(pl.spill_items, 1) * 8));
    CU(cudaMalloc(&e->sc.cursor1, sizeof(uint32_t) * std::max(pl.n_regions1, 1u)));
    CU(cudaMalloc(&e->sc.cursor2, sizeof(uint32_t) * std::max(pl.n_regions2, 1u)));
    CU(cudaMalloc(&e->sc.spill_cursor, sizeof(unsigned long long)));
    CU(cudaMalloc(&e->sc.err, sizeof(uint32_t)));
    CU(cudaMemset(e->sc.err, 0, sizeof(uint32_t)));
    if (out_bytes) {
        CU(cudaMalloc(&e->sc.out_list, out_bytes));
        CU(cudaMalloc(&e->sc.out_cursor, sizeof(uint32_t) * n_groups * pk_part_ocursor_words()));
    }
    e->sc.out_bytes = out_bytes;
    e->sc_plan = pl;
    return PK_OK;
}

// one-byte rows written contiguously out of group tables: 4-byte result items in fine position bins, un-permuted through
// shared-memory slices of the bitmap (pk_partition.cu)
static int fine_out_ok(const pk_engine *e, uint32_t row_stride, uint32_t col_offset) {
    if (!(e->unpermute && !e->h_utables.empty() && e->n_local <= 8 && row_stride == 1 && col_offset == 0)) return 0;
    return e->gfmt == 1 ? 2 : 1;
}

// rows for positions [p0, p0+n): the partitioned path for large batches, the direct kernel otherwise
static int probe_any(pk_engine *e, const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n,
                     uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, cudaStream_t s) {
    const uint32_t mode = e->cfg.probe_mode;
    const uint64_t sub = e->cfg.chunk_positions ? e->cfg.chunk_positions : PK_PART_MAX_N;
    for (uint64_t o = 0; o < n; o += sub) {
        const uint64_t m = std::min(sub, n - o);
        const bool part = mode == 2 || (mode == 0 && m >= (1ull << 20));
        if (part) {
            PkPartPlan pl;
            pk_part_plan(m, e->tune, &pl, e->n_local, fine_out_ok(e, row_stride, col_offset), e->cfg.k);
            int rc = ensure_scratch(e, pl); if (rc) return rc;
            if (pk_launch_probe_partitioned(d_words, d_mask, p0 + o, m, probe_ks(e), e->d_tables, e->h_tables.data(),
                                            e->h_utables.empty() ? nullptr : e->h_utables.data(), e->h_utables.empty() ? nullptr : e->d_utables, e->n_local,
                                            d_rows + o * row_stride, row_stride, col_offset, pl, e->sc, e->l2_prefetch, s, e->pev)) {
                pk_set_error("partitioned probe launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return PK_ECUDA;
            }
            e->stats.kernel_launches += 3 + (pl.pb2 ? 1 : 0) + (e->unpermute ? 1 : 0);
            e->stats.probe_launches += 1;
            e->pev_valid = true;
        } else {
            if (!e->h_utables.empty())
                pk_launch_probe_group(d_words, d_mask, p0 + o, m, probe_ks(e), e->d_utables, e->n_local, d_rows + o * row_stride, row_stride, col_offset, s);
            else
                pk_launch_probe(d_words, d_mask, p0 + o, m, e->ks, e->d_tables, e->n_local, d_rows + o * row_stride,
                                row_stride, col_offset, s);
            e->stats.kernel_launches += 1;
            e->stats.probe_launches += 1;
        }
        CU(cudaGetLastError());
    }
    return PK_OK;
}

static int check_part_error(pk_engine *e) {
    if (!e->sc.err) return PK_OK;
    uint32_t err = 0;
    CU(cudaMemcpy(&err, e->sc.err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) { pk_set_error("partitioned probe: spill list overflow (internal error)"); return PK_ECUDA; }
    return PK_OK;
}

extern "C" int pk_probe_device(pk_engine *e, const void *d_words, const void *d_mask, uint64_t p0, uint64_t n,
                               void *d_rows, uint32_t row_stride, uint32_t col_offset, void *stream) {
    NEED_FINAL(e);
    if (!d_words || !d_mask || !d_rows) { pk_set_error("null argument"); return PK_EINVAL; }
    if (col_offset + e->row_bytes > row_stride) { pk_set_error("row_stride %u too small for %u bytes at offset %u", row_stride, e->row_bytes, col_offset); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    return probe_any(e, (const uint64_t *)d_words, (const uint32_t *)d_mask, p0, n, (uint8_t *)d_rows, row_stride,
                     col_offset, stream ? (cudaStream_t)stream : e->stream);
}

// ------------------------------------------------------------------ host-level hot path
template <typename T> static int grow(T *&p, uint64_t &cap, uint64_t need) {
    if (cap >= need) return PK_OK;
    CU(cudaDeviceSynchronize());
    cudaFree(p); p = nullptr; cap = 0;
    CU(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return PK_OK;
}

// file images requested from anchor_genome_impl (pk_anchor_genome_bgzf): [0] step 1, [1] low-res
struct BgzfOut {
    uint8_t *const *gz; const uint64_t *gz_cap; uint8_t *const *gzi; const uint64_t *gzi_cap; uint64_t *sizes;
};
static int ensure_bgzf_tables(pk_engine *e) {
    if (e->z_tables) return PK_OK;
    std::vector<uint32_t> h(PK_BGZF_TABLE_WORDS);
    pk_bgzf_tables_host(h.data());
    CU(cudaMalloc(&e->z_tables, h.size() * 4));
    CU(cudaMemcpy(e->z_tables, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&e->z_totals, 4 * sizeof(unsigned long long)));
    return PK_OK;
}

// rank-local half of the genome-sharded path (pk_anchor_genome_plane): rows go to a caller-owned device plane in
// the concatenated numbering, nothing is reduced or copied back
struct PlaneOut { uint8_t *d_plane; uint64_t plane_rows; uint32_t row_stride; };

// destroys the per-chromosome events of one anchor_genome_impl call on every way out
struct EventBag {
    std::vector<cudaEvent_t> ev;
    ~EventBag() { for (auto x : ev) if (x) cudaEventDestroy(x); }
};

// Layout of an anchor in the concatenated ("cat") numbering all device buffers of one call use: chromosome c
// starts at cat_off[c], a multiple of 32 bases (it owns whole packed words), with at least one 'N' before it,
// so no window spans two chromosomes. Returns the total length (= rows a plane needs).
static uint64_t anchor_layout(uint32_t n_chroms, const uint64_t *lens, uint64_t *cat_off) {
    uint64_t ltot = 0;
    for (uint32_t c = 0; c < n_chroms; c++) {
        if (cat_off) cat_off[c] = ltot;
        ltot = (ltot + lens[c] + 1 + 31) & ~31ull;
    }
    return ltot;
}
extern "C" uint64_t pk_anchor_layout(uint32_t n_chroms, const uint64_t *lens, uint64_t *cat_off) {
    if (n_chroms && !lens) return 0;
    return anchor_layout(n_chroms, lens, cat_off);
}

static int anchor_genome_impl(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                              uint8_t *const *bitmap1, uint8_t *const *bitmap_low, uint64_t *const *bin_hist,
                              uint64_t *col_sums, uint64_t *nkmers_out, const BgzfOut *z, const PlaneOut *plane = nullptr) {
    NEED_FINAL(e);
    if (n_chroms && (!seqs || !lens)) { pk_set_error("null argument"); return PK_EINVAL; }
    const uint32_t k = e->cfg.k, step = e->cfg.lowres_step, rb = e->row_bytes, N = e->n_local;
    memset(&e->stats, 0, sizeof e->stats);
    // layout of the concatenated sequence: chromosome c starts at off[c], a multiple of 32 bases (so it owns
    // whole packed words), with at least one 'N' before it: no window spans two chromosomes
    std::vector<uint64_t> off(n_chroms), nk(n_chroms), binlen(n_chroms), nbins(n_chroms), lowoff(n_chroms), histoff(n_chroms);
    uint64_t ltot = 0, lowtot = 0, histtot = 0, postot = 0;
    for (uint32_t c = 0; c < n_chroms; c++) {
        if (!seqs[c] && lens[c]) { pk_set_error("null sequence %u", c); return PK_EINVAL; }
        off[c] = ltot;
        ltot = (ltot + lens[c] + 1 + 31) & ~31ull;      // == anchor_layout()
        nk[c] = lens[c] >= k ? lens[c] - k + 1 : 0;        // len < k: nothing (kmc_file.cpp:878-882)
        binlen[c] = nk[c] ? pk_bin_len(&e->cfg, nk[c]) : 0;
        const bool want_hist = bin_hist && bin_hist[c] && nk[c];
        if (want_hist && binlen[c] == 0) {
            pk_set_error("chromosome %u with %llu k-mers (< min_bin_count %u) has no defined bins", c,
                         (unsigned long long)nk[c], e->cfg.min_bin_count);
            return PK_EINVAL;
        }
        nbins[c] = want_hist ? (nk[c] + binlen[c] - 1) / binlen[c] : 0;
        lowoff[c] = lowtot; lowtot += (nk[c] + step - 1) / step;
        histoff[c] = histtot; histtot += nbins[c] * (N + 1);
        postot += nk[c];
        if (nkmers_out) nkmers_out[c] = nk[c];
    }
    if (z) {
        if (!z->gz || !z->gzi || !z->gz_cap || !z->gzi_cap || !z->sizes) { pk_set_error("null argument"); return PK_EINVAL; }
        const uint64_t nbytes[2] = {postot * rb, lowtot * rb};
        for (int i = 0; i < 2; i++) {
            if (!z->gz[i] || !z->gzi[i]) { pk_set_error("null argument"); return PK_EINVAL; }
            if (z->gz_cap[i] < pk_bgzf_bound_impl(nbytes[i]) || z->gzi_cap[i] < pk_bgzf_gzi_bound_impl(nbytes[i])) {
                pk_set_error("BGZF output %d: capacity %llu / %llu below pk_bgzf_bound %llu / pk_bgzf_gzi_bound %llu", i,
                             (unsigned long long)z->gz_cap[i], (unsigned long long)z->gzi_cap[i],
                             (unsigned long long)pk_bgzf_bound_impl(nbytes[i]), (unsigned long long)pk_bgzf_gzi_bound_impl(nbytes[i]));
                return PK_EINVAL;
            }
        }
    }
    if (postot == 0) {
        if (z) {       // nothing to write: both files are the bare EOF member, both indexes hold zero entries
            static const uint8_t eof[28] = {0x1F, 0x8B, 8, 4, 0, 0, 0, 0, 0, 0xFF, 6, 0, 0x42, 0x43, 2, 0, 0x1B, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 2; i++) {
                memcpy(z->gz[i], eof, sizeof eof);
                memset(z->gzi[i], 0, 8);
                z->sizes[2 * i] = sizeof eof; z->sizes[2 * i + 1] = 8;
            }
        }
        return PK_OK;
    }
    int rc = set_device(e); if (rc) return rc;
    if (z) {
        rc = ensure_bgzf_tables(e); if (rc) return rc;
        const uint64_t nbytes[2] = {postot * rb, lowtot * rb};
        rc = grow(e->z_cat, e->z_cat_cap, nbytes[0] + 16); if (rc) return rc;
        rc = grow(e->z_scratch, e->z_scratch_cap, pk_bgzf_scratch_bytes(nbytes[0])); if (rc) return rc;
        for (int i = 0; i < 2; i++) {
            rc = grow(e->z_gz[i], e->z_gz_cap[i], pk_bgzf_bound_impl(nbytes[i])); if (rc) return rc;
            rc = grow(e->z_gzi[i], e->z_gzi_cap[i], pk_bgzf_gzi_bound_impl(nbytes[i]) / 8 + 1); if (rc) return rc;
        }
    }
    const uint64_t nw = pk_packed_words(ltot);
    const uint64_t npos = ltot >= k ? ltot - k + 1 : 0;
    rc = grow(e->g_ascii, e->g_ascii_cap, ltot + 64); if (rc) return rc;
    rc = grow(e->g_words, e->g_words_cap, nw); if (rc) return rc;
    rc = grow(e->g_mask, e->g_mask_cap, nw); if (rc) return rc;
    if (plane) {
        if (!plane->d_plane || plane->plane_rows < ltot) { pk_set_error("plane of %llu rows is too small for %llu", (unsigned long long)plane->plane_rows, (unsigned long long)ltot); return PK_EINVAL; }
    } else {
        rc = grow(e->g_rows, e->g_rows_cap, ltot * rb); if (rc) return rc;
    }
    uint8_t *const rows_all = plane ? plane->d_plane : e->g_rows;
    const uint32_t rs = plane ? plane->row_stride : rb;         // bytes between rows of rows_all
    rc = grow(e->g_low, e->g_low_cap, (lowtot + 1) * rb); if (rc) return rc;
    rc = grow(e->d_hist, e->hist_cap, histtot + 1); if (rc) return rc;
    if (!e->ev[0]) for (auto &ev : e->ev) CU(cudaEventCreate(&ev));
    if (!e->copy_stream) CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    if (!e->in_stream) CU(cudaStreamCreateWithFlags(&e->in_stream, cudaStreamNonBlocking));
    cudaStream_t s = e->stream, cs = e->copy_stream, hs = e->in_stream;
    const uint32_t mode = e->cfg.probe_mode;
    const uint64_t sub = e->cfg.chunk_positions ? e->cfg.chunk_positions : PK_PART_MAX_N;
    bool pipelined = mode == 2 || (mode == 0 && npos >= (1ull << 20));
    EventBag bag_in, bag_done;
    bag_in.ev.assign(n_chroms, nullptr); bag_done.ev.assign(n_chroms, nullptr);
    std::vector<cudaEvent_t> &in_done = bag_in.ev, &done = bag_done.ev;
    CU(cudaEventRecord(e->ev[0], s));
    CU(cudaStreamWaitEvent(hs, e->ev[0], 0));
    // ---- H2D on the input stream, chromosome by chromosome
    CU(cudaMemsetAsync(e->g_ascii, 'N', ltot + 64, hs));
    for (uint32_t c = 0; c < n_chroms; c++) {
        if (lens[c]) CU(cudaMemcpyAsync(e->g_ascii + off[c], seqs[c], lens[c], cudaMemcpyHostToDevice, hs));
        CU(cudaEventCreateWithFlags(&in_done[c], cudaEventDisableTiming));
        CU(cudaEventRecord(in_done[c], hs));
    }
    CU(cudaMemsetAsync(e->d_hist, 0, (histtot + 1) * 8, s));
    CU(cudaMemsetAsync(e->d_colsums, 0, N * 8, s));
    // ---- batches. The partitioned probe streams every table through L2 once per batch, so few batches are
    // best for the kernels; but with ONE batch nothing overlaps the first H2D and the last D2H (2.7 + 1.3 ms of
    // an 11.1 ms call on configs[1], profiles/r1f_bench.json). Two batches of whole chromosomes let the copy
    // engines run under the other batch's kernels while the tables are still read only twice.
    // A batch is one partitioned launch: at most `sub` positions (PK_PART_MAX_N, or the chunk_positions knob). A
    // genome beyond that gets more batches; a single chromosome beyond it takes the sub-launch loop of probe_any.
    struct Batch { uint32_t c0, c1; uint64_t base, npos; PkPartPlan pl; };
    std::vector<Batch> batches;
    uint32_t nb = 1;
    if (pipelined && e->e2e_batches > 1 && npos >= e->e2e_batch_min) {
        // every batch streams the tables once more (~5 TB/s) and hides about half of the PCIe copies (~52 GB/s):
        // worth it while table bytes < ~48 x copied bytes (configs[1]: 2.6 GB vs 13 GB; 64 genomes at k=31: 114 vs 65 GB)
        double table_bytes = 0;
        if (!e->h_utables.empty()) for (const PkTable &t : e->h_utables) table_bytes += 32.0 * t.n_buckets;
        else for (const HostTable &t : e->tabs) table_bytes += 32.0 * t.dev.n_buckets;
        const double copied = (double)ltot + (plane ? 0.0 : (double)postot * rb);
        if (table_bytes < 48.0 * copied || e->e2e_batch_force) nb = (uint32_t)e->e2e_batches;
    }
    if (pipelined && npos > sub) nb = std::max<uint32_t>(nb, (uint32_t)((npos + sub - 1) / sub));
    for (int attempt = 0; attempt < 4; attempt++) {
        batches.clear();
        uint32_t c = 0;
        for (uint32_t b = 0; b < nb && c < n_chroms; b++) {
            // cut after the chromosome at which the running length first reaches (b+1)/nb of the total
            const uint64_t target = ltot * (b + 1) / nb;
            Batch bt{};
            bt.c0 = c;
            // ... e2e_front_small: a chromosome that straddles the target goes to the LATER batch unless 3/4 of it lie before
            // the target (the first batch's H2D is exposed in full, the later batches' copies run under kernels)
            while (c < n_chroms && (b + 1 == nb || c == bt.c0 ||
                                    off[c] + (e->e2e_front_small ? lens[c] - lens[c] / 4 : lens[c] / 2) <= target)) c++;
            bt.c1 = c;
            bt.base = off[bt.c0];
            const uint64_t end = (bt.c1 < n_chroms ? off[bt.c1] : ltot);
            bt.npos = end > bt.base + k - 1 ? end - bt.base - (k - 1) : 0;
            batches.push_back(bt);
        }
        if (!batches.empty()) batches.back().c1 = n_chroms;
        uint64_t mxb = 0;
        for (const Batch &bt : batches) mxb = std::max(mxb, bt.npos);
        if (!pipelined || mxb <= sub) break;
        if (attempt == 3 || nb >= n_chroms) { pipelined = false; nb = 1; attempt = 2; continue; }   // one last pass lays out the single batch
        nb = std::min<uint32_t>(n_chroms, nb * 2);
    }
    if (pipelined) {
        PkPartPlan mx{};
        for (auto &bt : batches) {
            pk_part_plan(bt.npos ? bt.npos : 1, e->tune, &bt.pl, N, fine_out_ok(e, rs, 0), k);
            mx.buf1_items = std::max(mx.buf1_items, bt.pl.buf1_items); mx.buf2_items = std::max(mx.buf2_items, bt.pl.buf2_items);
            mx.spill_items = std::max(mx.spill_items, bt.pl.spill_items);
            mx.n_regions1 = std::max(mx.n_regions1, bt.pl.n_regions1); mx.n_regions2 = std::max(mx.n_regions2, bt.pl.n_regions2);
            mx.out_shift = std::max(mx.out_shift, bt.pl.out_shift);
            mx.out_bytes = std::max(mx.out_bytes, bt.pl.out_bytes);
        }
        rc = ensure_scratch(e, mx); if (rc) return rc;
    }
    bool first = true;
    uint64_t z_rows_done = 0, z_members_done = 0;       // BGZF mode: rows in the output stream so far, members deflated so far
    for (const Batch &bt : batches) {
        const PkPartPlan &pl = bt.pl;
        uint8_t *rows_b = rows_all + bt.base * rs;
        if (pipelined) pk_part_begin(N, pl, e->sc, s);
        // ---- pack (+ K1 of the partitioned probe) per chromosome as soon as its bytes are on the device
        for (uint32_t c = bt.c0; c < bt.c1; c++) {
            CU(cudaStreamWaitEvent(s, in_done[c], 0));
            const uint64_t w0 = off[c] / 32, w1 = c + 1 < n_chroms ? off[c + 1] / 32 : nw;
            pk_launch_pack(e->g_ascii + off[c], ltot + 64 - off[c], w1 - w0, e->g_words + w0, e->g_mask + w0, s);
            e->stats.kernel_launches += 1;
            if (pipelined && nk[c]) {
                pk_part_append(e->g_words, e->g_mask, bt.base, off[c] - bt.base, nk[c], probe_ks(e), N, rows_b, rs, 0, pl, e->sc, s);
                e->stats.kernel_launches += 1;
            }
        }
        if (first) { CU(cudaEventRecord(e->ev[1], s)); CU(cudaEventRecord(e->ev[2], s)); }
        if (pipelined) {
            pk_part_probe(e->g_words, e->g_mask, bt.base, probe_ks(e), e->h_tables.data(), e->h_utables.empty() ? nullptr : e->h_utables.data(),
                          e->h_utables.empty() ? nullptr : e->d_utables, N, rows_b, rs, 0, pl, e->sc, e->l2_prefetch, s, nullptr);
            e->stats.kernel_launches += (pl.pb2 ? 1 : 0) + 2 * ((N + 31) / 32);
            e->stats.probe_launches += 1;
        } else {
            rc = probe_any(e, e->g_words, e->g_mask, 0, npos, rows_all, rs, 0, s); if (rc) return rc;
        }
        if (first) CU(cudaEventRecord(e->ev[3], s));
        first = false;
        // ---- per chromosome: un-permute its position bins, reduce, and send rows home on the copy stream
        uint32_t next_bin = 0;
        for (uint32_t c = bt.c0; c < bt.c1; c++) {
            if (!nk[c]) continue;
            if (pipelined && e->sc.out_list) {
                const uint32_t b1 = (uint32_t)((off[c] - bt.base + nk[c] - 1) >> pl.out_shift) + 1;
                if (b1 > next_bin) {
                    pk_part_unpermute(next_bin, b1, N, rows_b, rs, 0, pl, e->sc, s);
                    e->stats.kernel_launches += 1;
                    next_bin = b1;
                }
            }
            if (plane) continue;                    // the caller exchanges, reduces and stores (genome-sharded path)
            const uint8_t *rows_c = rows_all + off[c] * rb;
            uint8_t *low_c = e->g_low + lowoff[c] * rb;
            const bool want_low = (bitmap_low && bitmap_low[c]) || z;
            if (nbins[c] || col_sums || want_low) {
                pk_launch_reduce(rows_c, rb, N, 0, nk[c], nbins[c] ? binlen[c] : 0, nbins[c] ? e->d_hist + histoff[c] : nullptr,
                                 col_sums ? e->d_colsums : nullptr, want_low ? low_c : nullptr, step, s);
                e->stats.kernel_launches += 1 + (want_low ? 1 : 0);
            }
            if (z) {     // one stream per anchor, chromosomes back to back (cpp/anchor.cpp:167: bgzf_write appends chunk after chunk)
                CU(cudaMemcpyAsync(e->z_cat + z_rows_done * rb, rows_c, nk[c] * rb, cudaMemcpyDeviceToDevice, s));
                z_rows_done += nk[c];
            }
            CU(cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
            CU(cudaEventRecord(done[c], s));
            CU(cudaStreamWaitEvent(cs, done[c], 0));
            if (bitmap1 && bitmap1[c]) CU(cudaMemcpyAsync(bitmap1[c], rows_c, nk[c] * rb, cudaMemcpyDeviceToHost, cs));
            if (bitmap_low && bitmap_low[c]) CU(cudaMemcpyAsync(bitmap_low[c], low_c, ((nk[c] + step - 1) / step) * rb, cudaMemcpyDeviceToHost, cs));
        }
        if (z && !plane) {
            // the BGZF members this batch completed are deflated on the copy stream (idle in this mode: the rows stay on
            // the device) while the next batch is partitioned and probed on the compute stream
            const bool last = &bt == &batches.back();
            const uint64_t m1 = last ? pk_bgzf_blocks_impl(postot * rb) : z_rows_done * rb / PK_BGZF_PAYLOAD;
            if (m1 > z_members_done) {
                pk_launch_bgzf_encode(e->z_cat, postot * rb, rb, z_members_done, m1, e->z_scratch, e->z_tables, cs);
                e->stats.kernel_launches += 1;
                z_members_done = m1;
            }
        }
    }
    if (z) {
        // cs is ordered after every chromosome's rows, low-res rows and reductions (done[c])
        pk_launch_bgzf_finish(e->z_cat, postot * rb, e->z_gz[0], e->z_gzi[0], e->z_totals, e->z_scratch, cs);
        pk_launch_bgzf(e->g_low, lowtot * rb, rb, e->z_gz[1], e->z_gzi[1], e->z_totals + 2, e->z_scratch, e->z_tables, cs);
        e->stats.kernel_launches += 5;
        CU(cudaGetLastError());
        unsigned long long tot[4];
        CU(cudaMemcpyAsync(tot, e->z_totals, sizeof tot, cudaMemcpyDeviceToHost, cs));
        CU(cudaStreamSynchronize(cs));
        for (int i = 0; i < 2; i++) {
            if (tot[2 * i] > z->gz_cap[i] || tot[2 * i + 1] > z->gzi_cap[i]) { pk_set_error("BGZF image larger than its bound (internal error)"); return PK_ECUDA; }
            CU(cudaMemcpyAsync(z->gz[i], e->z_gz[i], tot[2 * i], cudaMemcpyDeviceToHost, cs));
            CU(cudaMemcpyAsync(z->gzi[i], e->z_gzi[i], tot[2 * i + 1], cudaMemcpyDeviceToHost, cs));
            z->sizes[2 * i] = tot[2 * i]; z->sizes[2 * i + 1] = tot[2 * i + 1];
        }
    }
    CU(cudaEventRecord(e->ev[4], s));
    std::vector<unsigned long long> hist_h, col_h;
    if (histtot) {
        hist_h.resize(histtot);
        CU(cudaMemcpyAsync(hist_h.data(), e->d_hist, histtot * 8, cudaMemcpyDeviceToHost, s));
    }
    if (col_sums) {
        col_h.resize(N);
        CU(cudaMemcpyAsync(col_h.data(), e->d_colsums, N * 8, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    CU(cudaStreamSynchronize(cs));
    CU(cudaEventRecord(e->ev[5], s));
    CU(cudaEventSynchronize(e->ev[5]));
    rc = check_part_error(e); if (rc) return rc;
    for (uint32_t c = 0; c < n_chroms; c++)
        if (nbins[c]) memcpy(bin_hist[c], hist_h.data() + histoff[c], nbins[c] * (N + 1) * 8);
    if (col_sums) for (uint32_t g = 0; g < N; g++) col_sums[g] += col_h[g];
    CU(cudaEventElapsedTime(&e->stats.h2d_ms, e->ev[0], e->ev[1]));      // H2D overlapped with pack + K1
    e->stats.pack_ms = 0.f;
    CU(cudaEventElapsedTime(&e->stats.probe_ms, e->ev[2], e->ev[3]));
    CU(cudaEventElapsedTime(&e->stats.reduce_ms, e->ev[3], e->ev[4]));
    CU(cudaEventElapsedTime(&e->stats.d2h_ms, e->ev[4], e->ev[5]));
    CU(cudaEventElapsedTime(&e->stats.total_ms, e->ev[0], e->ev[5]));
    e->stats.positions = postot;
    e->stats.probes = postot * N;
    e->last_lowtot = (!plane && ((bitmap_low != nullptr) || z)) ? lowtot : 0;
    return PK_OK;
}

extern "C" int pk_anchor_genome(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                                uint8_t *const *bitmap1, uint8_t *const *bitmap_low, uint64_t *const *bin_hist,
                                uint64_t *col_sums, uint64_t *nkmers_out) {
    return anchor_genome_impl(e, n_chroms, seqs, lens, bitmap1, bitmap_low, bin_hist, col_sums, nkmers_out, nullptr);
}

extern "C" int pk_anchor_genome_bgzf(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                                     uint8_t *const *gz, const uint64_t *gz_cap, uint8_t *const *gzi, const uint64_t *gzi_cap,
                                     uint64_t *sizes, uint64_t *const *bin_hist, uint64_t *col_sums, uint64_t *nkmers_out) {
    const BgzfOut z{gz, gz_cap, gzi, gzi_cap, sizes};
    return anchor_genome_impl(e, n_chroms, seqs, lens, nullptr, nullptr, bin_hist, col_sums, nkmers_out, &z);
}

extern "C" int pk_anchor_genome_plane(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                                      void *d_plane, uint64_t plane_rows, uint32_t row_stride, uint64_t *nkmers_out) {
    if (!e || !d_plane) { pk_set_error("null argument"); return PK_EINVAL; }
    if (row_stride < e->row_bytes) { pk_set_error("row_stride %u below the shard's %u row bytes", row_stride, e->row_bytes); return PK_EINVAL; }
    const PlaneOut p{(uint8_t *)d_plane, plane_rows, row_stride};
    return anchor_genome_impl(e, n_chroms, seqs, lens, nullptr, nullptr, nullptr, nullptr, nkmers_out, nullptr, &p);
}

extern "C" int pk_paircount_bins_device(pk_engine *e, const void *d_rows_low, uint32_t row_stride, uint32_t n_cols, uint64_t n_rows,
                                        uint32_t rows_per_bin, void *d_counts, void *stream) {
    if (!e || !d_rows_low || !d_counts) { pk_set_error("null argument"); return PK_EINVAL; }
    if (n_cols == 0 || n_cols > 8 * row_stride || n_cols > 8192 || rows_per_bin == 0) { pk_set_error("bad n_cols %u / rows_per_bin %u", n_cols, rows_per_bin); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    pk_launch_paircount_bins((const uint8_t *)d_rows_low, row_stride, n_cols, n_rows, rows_per_bin, (uint32_t *)d_counts,
                             stream ? (pk_stream_t)stream : e->stream);
    CU(cudaGetLastError());
    return PK_OK;
}

// Pair-count bins of the anchor the last pk_anchor_genome / pk_anchor_genome_bgzf call processed, from its low-res rows
// (still resident on the device): counts[sum_c ceil(n_low_c / rows_per_bin)][N_local], chromosome after chromosome.
extern "C" int pk_anchor_paircount_bins(pk_engine *e, uint32_t n_chroms, const uint64_t *nkmers, uint32_t bin_positions, uint32_t *counts) {
    NEED_FINAL(e);
    if (!nkmers || !counts) { pk_set_error("null argument"); return PK_EINVAL; }
    const uint32_t step = e->cfg.lowres_step, rb = e->row_bytes, N = e->n_local;
    if (bin_positions < step) { pk_set_error("bin of %u positions is shorter than the low-res step %u", bin_positions, step); return PK_EINVAL; }
    const uint32_t rpb = (bin_positions + step - 1) / step;
    uint64_t lowtot = 0, bins = 0;
    for (uint32_t c = 0; c < n_chroms; c++) { const uint64_t nl = (nkmers[c] + step - 1) / step; lowtot += nl; bins += (nl + rpb - 1) / rpb; }
    if (!e->g_low || e->last_lowtot != lowtot) { pk_set_error("the last anchoring call left %llu low-res rows, the chromosome list describes %llu", (unsigned long long)e->last_lowtot, (unsigned long long)lowtot); return PK_ESTATE; }
    if (!bins) return PK_OK;
    int rc = set_device(e); if (rc) return rc;
    uint32_t *d_counts = nullptr;
    CU(cudaMalloc(&d_counts, bins * N * sizeof(uint32_t)));
    struct Freer { void *p; ~Freer() { cudaFree(p); } } f{d_counts};
    uint64_t lo = 0, bo = 0;
    for (uint32_t c = 0; c < n_chroms; c++) {
        const uint64_t nl = (nkmers[c] + step - 1) / step;
        pk_launch_paircount_bins(e->g_low + lo * rb, rb, N, nl, rpb, d_counts + bo * N, e->stream);
        lo += nl; bo += (nl + rpb - 1) / rpb;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts, d_counts, bins * N * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PK_OK;
}

extern "C" uint64_t pk_bgzf_bound(uint64_t n_bytes) { return pk_bgzf_bound_impl(n_bytes); }
extern "C" uint64_t pk_bgzf_gzi_bound(uint64_t n_bytes) { return pk_bgzf_gzi_bound_impl(n_bytes); }
extern "C" int pk_bgzf_compress_device(pk_engine *e, const void *d_in, uint64_t n_bytes, uint32_t match_dist, void *d_gz,
                                       void *d_gzi, void *d_totals, void *stream) {
    if (!e || (!d_in && n_bytes) || !d_gz || !d_gzi || !d_totals) { pk_set_error("null argument"); return PK_EINVAL; }
    if (match_dist < 1 || match_dist > 32768) { pk_set_error("match_dist %u out of 1..32768", match_dist); return PK_EINVAL; }
    if ((uintptr_t)d_gzi & 7 || (uintptr_t)d_totals & 7) { pk_set_error("d_gzi / d_totals must be 8-byte aligned"); return PK_EINVAL; }
    int rc = set_device(e); if (rc) return rc;
    rc = ensure_bgzf_tables(e); if (rc) return rc;
    rc = grow(e->z_scratch, e->z_scratch_cap, pk_bgzf_scratch_bytes(n_bytes)); if (rc) return rc;
    pk_launch_bgzf((const uint8_t *)d_in, n_bytes, match_dist, (uint8_t *)d_gz, (unsigned long long *)d_gzi,
                   (unsigned long long *)d_totals, e->z_scratch, e->z_tables, stream ? (pk_stream_t)stream : e->stream);
    CU(cudaGetLastError());
    return PK_OK;
}

extern "C" int pk_anchor_chrom(pk_engine *e, const char *ascii, uint64_t len, uint8_t *bitmap1, uint8_t *bitmap_low,
                               uint64_t *bin_hist, uint64_t *col_sums, uint64_t *nkmers_out) {
    const char *seqs[1] = {ascii};
    uint8_t *b1[1] = {bitmap1}, *bl[1] = {bitmap_low};
    uint64_t *bh[1] = {bin_hist};
    uint64_t nk = 0;
    int rc = pk_anchor_genome(e, 1, seqs, &len, b1, bl, bh, col_sums, &nk);
    if (nkmers_out) *nkmers_out = rc == PK_OK ? nk : 0;
    return rc;
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
    const uint64_t nw = pk_packed_words(len);
    rc = grow(e->g_ascii, e->g_ascii_cap, len + 64); if (rc) return rc;
    rc = grow(e->g_words, e->g_words_cap, nw); if (rc) return rc;
    rc = grow(e->g_mask, e->g_mask_cap, nw); if (rc) return rc;
    rc = grow(e->g_rows, e->g_rows_cap, (len + 1) * rb); if (rc) return rc;
    rc = grow(e->g_u32, e->g_u32_cap, nk); if (rc) return rc;
    cudaStream_t s = e->stream;
    memset(&e->stats, 0, sizeof e->stats);
    CU(cudaMemcpyAsync(e->g_ascii, read, len, cudaMemcpyHostToDevice, s));
    pk_launch_pack(e->g_ascii, len, nw, e->g_words, e->g_mask, s);
    rc = probe_any(e, e->g_words, e->g_mask, 0, nk, e->g_rows, rb, 0, s); if (rc) return rc;
    pk_launch_rows_to_u32(e->g_rows, rb, byte_off, nbytes, mask, nk, e->g_u32, s);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counters, e->g_u32, nk * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    rc = check_part_error(e); if (rc) return rc;
    if (shift) for (uint64_t i = 0; i < nk; i++) counters[i] <<= shift;
    *n_out = nk;
    return PK_OK;
}

extern "C" int pk_engine_tune(pk_engine *e, const char *name, int value) {
    if (!e || !name) { pk_set_error("null argument"); return PK_EINVAL; }
    const std::string n(name);
    if (n == "k3_window") { e->tune.window = value; return PK_OK; }
    else if (n == "k3w_variant") { if (value >= -1 && value < pk_part_n_wvariants()) e->tune.wvariant = value; return PK_OK; }
    else if (n == "k3w_group") { if (value != 0 && value != 1 && value != 2 && value != 4) { pk_set_error("k3w_group %d: must be 0 (auto), 1, 2 or 4", value); return PK_EINVAL; } e->tune.wgroup = value; return PK_OK; }
    else if (n == "k3_rank_atomic") { e->tune.rank_atomic = value < 0 ? 0 : value > 2 ? 2 : value; return PK_OK; }   // 2: no output (timing ablation)
    else if (n == "gather_dst_mode") { e->gather_dst_mode = value ? 1 : 0; return PK_OK; }
    else if (n == "k3w_big") { e->tune.wbig = value; return PK_OK; }
    else if (n == "compact_items") { e->tune.compact = value ? 1 : 0; return PK_OK; }
    else if (n == "k3_lean") { if (value >= 0 && value <= pk_part_n_lvariants()) e->tune.lean = value; return PK_OK; }
    else if (n == "k3_l2") { if (value >= 0 && value <= pk_part_n_gvariants()) e->tune.k3_l2 = value; return PK_OK; }
    else if (n == "k1_roll") { e->tune.k1_roll = value ? 1 : 0; return PK_OK; }
    else if (n == "fine_out") { e->tune.fine_out = value ? 1 : 0; return PK_OK; }
    else if (n == "fine_shift") { e->tune.fine_shift = value < 0 ? 0 : value > 24 ? 24 : value; return PK_OK; }
    else if (n == "k3_variant") { if (value >= -1 && value < pk_part_n_variants()) e->tune.variant = value; return PK_OK; }
    else if (n == "l2_prefetch") { e->l2_prefetch = value; return PK_OK; }
    else if (n == "group_tables") {        // 0: per-genome tables only; takes effect at the next pk_engine_finalize
        if (!value)
            for (auto &t : e->tabs)
                if (t.sealed) { pk_set_error("per-genome tables were freed (group_only): the group tables cannot be dropped"); return PK_ESTATE; }
        e->union_tables = value ? 1 : 0;
        if (!value) drop_group_tables(e);       // the per-genome tables answer from here on
        return PK_OK;
    }
    else if (n == "group_g32") {           // 0: 64-bit group slots whatever k; 1: 32-bit slots where k and the table size allow; 2: wherever k
        e->allow_g32 = value < 0 ? 0 : value > 2 ? 2 : value;   // allows, small tables padded to 2^(2k-20) buckets (tests). Takes effect when the group tables are next built from scratch
        return PK_OK;
    }
    else if (n == "group_only") {          // free per-genome tables once their group table is built (finalize / seal_group)
        e->group_only = value ? 1 : 0;
        return PK_OK;
    }
    else if (n == "e2e_front_small") { e->e2e_front_small = value ? 1 : 0; return PK_OK; }
    else if (n == "e2e_batch_min") { e->e2e_batch_min = value < 0 ? 0 : (uint64_t)value; return PK_OK; }
    else if (n == "e2e_batches") { if (value < 1 || value > 8) { pk_set_error("e2e_batches %d out of 1..8", value); return PK_EINVAL; } e->e2e_batches = value; e->e2e_batch_force = 1; return PK_OK; }
    else if (n == "unpermute") {
        if (e->unpermute != value) {       // the scratch layout depends on it: drop it, the next launch re-allocates
            int rc = set_device(e); if (rc) return rc;
            CU(cudaDeviceSynchronize());
            free_scratch(e);
            e->unpermute = value;
        }
        return PK_OK;
    }
    else { pk_set_error("unknown tuning knob '%s'", name); return PK_EINVAL; }
    return PK_OK;
}

extern "C" int pk_engine_stats(const pk_engine *e, pk_stats *out) {
    if (!e || !out) { pk_set_error("null argument"); return PK_EINVAL; }
    *out = e->stats;
    out->k_partition_ms = out->k_fine_ms = out->k_probe_ms = out->k_spill_ms = out->k_unpermute_ms = 0.f;
    out->k_probe_window = (float)e->tune.last_window;
    if (e->pev_valid) {      // kernels of the last partitioned launch, timed on their own stream
        CU(cudaEventSynchronize(e->pev[5]));
        CU(cudaEventElapsedTime(&out->k_unpermute_ms, e->pev[4], e->pev[5]));
        CU(cudaEventElapsedTime(&out->k_partition_ms, e->pev[0], e->pev[1]));
        CU(cudaEventElapsedTime(&out->k_fine_ms, e->pev[1], e->pev[2]));
        CU(cudaEventElapsedTime(&out->k_probe_ms, e->pev[2], e->pev[3]));
        CU(cudaEventElapsedTime(&out->k_spill_ms, e->pev[3], e->pev[4]));
    }
    return PK_OK;
}
