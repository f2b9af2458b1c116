/* pk_anchor.h — C ABI of libpkanchor.so, the B200-native pan-kmer anchoring engine.
 *
 * This is the drop-in boundary for the `panagram index` anchoring path. Each
 * entry point names the reference interface it replaces (paths relative to the
 * kjenike/panagram tree). Plain pointers and sizes only; no C++ or torch types;
 * no exceptions cross the boundary. All functions return PK_OK (0) or a negative
 * pk_status; pk_last_error() gives a thread-local message for the last failure.
 *
 * Semantics, bit-exact to the reference:
 *   row[p] bit g = 1  iff  canon(seq[p:p+k]) is in K_g
 *   canon(x) = min(x, revcomp(x)) as a 2k-bit integer, A=0 C=1 G=2 T=3, first base
 *   most significant (kmer_api.h:373-386); only ACGTacgt are symbols
 *   (kmer_api.h:264-275): a window containing any other byte gives an all-zero row
 *   (kmc_file.cpp:972-993). Genome g lives in byte g/8, bit g%8 of a row
 *   (cpp/anchor.cpp:155-159). k <= 32.
 *
 * Threading: an engine is bound to one CUDA device; calls on one engine must be
 * serialised by the caller. Multi-GPU = one engine (one process) per GPU, each
 * owning a contiguous genome shard [genome_begin, genome_end).
 */
#ifndef PK_ANCHOR_H
#define PK_ANCHOR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 1

typedef enum pk_status {
    PK_OK = 0,
    PK_EINVAL = -1,       /* bad argument */
    PK_EIO = -2,          /* file missing / bad KMC marker (OpenForRA returning false, kmc_file.cpp:30-39) */
    PK_ECUDA = -3,        /* CUDA runtime error */
    PK_ENOMEM = -4,       /* host or device allocation failed / table full */
    PK_EUNSUPPORTED = -5, /* k > 32, Quake-mode DB, non-canonical DB */
    PK_ESTATE = -6        /* call out of order (e.g. probe before finalize) */
} pk_status;

int pk_abi_version(void);
const char *pk_last_error(void);
/* number of CUDA devices visible; 0 when none (never an error) */
int pk_device_count(void);

/* ---- KMC database reader -------------------------------------------------
 * Replaces CKMCFile::OpenForRA + ReadParamsFrom_prefix_file_buf
 * (KMC/kmc_api/kmc_file.cpp:25-53,178-325; bound as KMCFile.OpenForRA /
 * KMCFile.Info in KMC/py_kmc_api/py_kmc_api.cpp:85,98). Host only. Accepts both
 * KMC1 (kmc_tools output: .onehot, bitvec{i}) and KMC2 (kmc output: .count). */
typedef struct pk_kmcdb pk_kmcdb;
typedef struct pk_kmcdb_info {
    uint32_t kmer_length, mode, counter_size, lut_prefix_length, signature_len;
    uint32_t kmc_version;       /* 0 = KMC1, 0x200 = KMC2 */
    uint32_t both_strands;
    uint32_t _pad;
    uint64_t min_count, max_count, total_kmers;
} pk_kmcdb_info;
int pk_kmcdb_open(const char *prefix, pk_kmcdb **out);
int pk_kmcdb_info_get(const pk_kmcdb *db, pk_kmcdb_info *info);
void pk_kmcdb_close(pk_kmcdb *db);

/* ---- FASTA reader ---------------------------------------------------------
 * Replaces the getline loop of KMCdb::anchor_fasta (cpp/anchor.cpp:74-100) / Genome.iter_fasta
 * (panagram/index.py:922-930) in front of the engine. Host only. Record name = header up to the first ' '
 * (cpp/anchor.cpp:84,97); sequence = the lines concatenated verbatim ('\r' stays, as with getline). strip_cr != 0
 * applies what `kmc -fm` / Biopython see instead: a trailing '\r' is cut from every line, bytes < 32 are dropped
 * from the sequence, the name is the header's first whitespace-separated token. Sequences live in page-locked
 * memory (plain memory when no CUDA driver is present) owned by the handle; pointers stay valid until
 * pk_fasta_close. gzip / bgzip files (recognised by their magic, as Genome.iter_fasta recognises them by suffix,
 * index.py:922-930) are inflated with zlib, every member of the file. */
typedef struct pk_fasta pk_fasta;
int pk_fasta_open(const char *path, int strip_cr, pk_fasta **out);
uint32_t pk_fasta_n_records(const pk_fasta *fa);
int pk_fasta_record(const pk_fasta *fa, uint32_t i, const char **name, const uint8_t **seq, uint64_t *len);
void pk_fasta_close(pk_fasta *fa);

/* ---- engine ---------------------------------------------------------------
 * Replaces KMCdb::KMCdb (cpp/anchor.cpp:21-35) / Genome._load_kmc
 * (panagram/index.py:847-863): instead of ceil(N/32) merged "bitvec" databases
 * searched by LUT + binary search, the engine keeps one bucketed hash table per
 * genome in HBM. */
typedef struct pk_engine pk_engine;
typedef struct pk_config {
    uint32_t k;              /* k-mer length, 1..32 */
    uint32_t n_genomes;      /* N: total genomes = bit columns of a full row */
    uint32_t genome_begin;   /* this engine's shard [genome_begin, genome_end) ... */
    uint32_t genome_end;     /* ... multiples of 8 unless genome_end == n_genomes */
    int32_t device;          /* CUDA ordinal */
    uint32_t lowres_step;    /* 100 (cpp/anchor.cpp:170; index.py lowres_step) */
    uint32_t max_bin_len;    /* 200000 (cpp/anchor.cpp:114; max_bin_kbp*1000) */
    uint32_t min_bin_count;  /* 100 (cpp/anchor.cpp:116; min_bin_count) */
    float load_factor;       /* table fill target, 0 < f <= 0.9; 0 selects the default 0.5 */
    uint32_t chunk_positions;/* max positions per probe launch (0 = default 512 Mi); a tuning/testing knob */
    uint32_t probe_mode;     /* 0 auto (partitioned for launches >= 1 Mi positions), 1 direct, 2 partitioned */
} pk_config;
int pk_engine_create(const pk_config *cfg, pk_engine **out);
void pk_engine_destroy(pk_engine *e);

/* Table construction. genome ids are GLOBAL (0..N-1); ids outside this engine's
 * shard are accepted and ignored (so every rank can run the same loop).
 *
 * pk_engine_reserve: size genome g's table for up to max_keys distinct k-mers.
 * pk_engine_add_kmc: all k-mers of a per-genome KMC database with
 *   min_count <= counter <= max_count (the filter of kmc_file.cpp:1396), i.e. the
 *   set K_g that rule kmc_count produces (workflow/Snakefile:81-110). Reserves
 *   automatically.
 * pk_engine_add_bitvec: a merged bitvec database (rule kmc_bitvec,
 *   workflow/Snakefile:54-69): counter bit j set => k-mer belongs to genome
 *   first_genome + j (index.py:407-415). Reserves automatically.
 * pk_engine_add_keys: canonical k-mer integers from host memory.
 * pk_engine_add_sequence: every canonical k-mer of an ASCII sequence (what
 *   `kmc -ci1 -fm` would count for a FASTA record; kmc_core/splitter.cpp:44-47).
 *   Needs a prior pk_engine_reserve. */
int pk_engine_reserve(pk_engine *e, uint32_t genome, uint64_t max_keys);
int pk_engine_add_kmc(pk_engine *e, uint32_t genome, const char *kmc_prefix);
int pk_engine_add_bitvec(pk_engine *e, uint32_t first_genome, const char *kmc_prefix);
int pk_engine_add_keys(pk_engine *e, uint32_t genome, const uint64_t *canon_kmers, uint64_t n);
int pk_engine_add_sequence(pk_engine *e, uint32_t genome, const char *ascii, uint64_t len);
/* as pk_engine_add_sequence, from ASCII already resident in device memory */
int pk_engine_add_sequence_device(pk_engine *e, uint32_t genome, const void *d_ascii, uint64_t len);
int pk_engine_finalize(pk_engine *e);
/* pk_engine_seal_group: build the group table (below) of local genomes [8*group, 8*group+8) NOW, from their
 * per-genome tables as they stand. With the "group_only" knob set (pk_engine_tune) those per-genome tables are
 * freed at once — they are build intermediates then, and every lookup is answered from the group tables — so
 * a host that adds genomes group by group never holds more than one group's per-genome tables (64 genomes of
 * 150 Mbp at k=31: ~110 GB of group tables instead of 154 GB + 110 GB). pk_engine_finalize seals what is left.
 * Adding k-mers to a sealed genome is PK_ESTATE. */
int pk_engine_seal_group(pk_engine *e, uint32_t group);

typedef struct pk_table_stats {
    uint64_t n_keys;        /* distinct k-mers stored */
    uint64_t n_buckets;     /* 32-byte buckets (4 x uint64 slots) */
    uint64_t n_overflow;    /* keys not in their home bucket */
    uint64_t bytes;         /* device bytes */
} pk_table_stats;
int pk_engine_table_stats(const pk_engine *e, uint32_t genome, pk_table_stats *out);
/* Group tables: pk_engine_finalize also merges the per-genome tables of every 8 consecutive local genomes into
 * ONE table whose slots carry an 8-bit membership mask (the structure of the reference's merged "bitvec"
 * database, workflow/Snakefile:54-69 / index.py:407-426, as a hash table), so that the partitioned probe
 * answers 8 genomes with one bucket read. n_keys = distinct k-mers of the group; PK_ESTATE when group tables are
 * switched off (pk_engine_tune "group_tables" 0) or could not be built. group = local genome index / 8. */
int pk_engine_group_stats(const pk_engine *e, uint32_t group, pk_table_stats *out);

/* pk_engine_sample_kmers: a uniform sample of the engine's k-mers with their genomes — every canonical k-mer x of any
 * local genome with hash32(x) < hmax (hmax = 0: every k-mer), each once per group of 8 local genomes:
 * keys[i] = x, tags[i] = (local_group << 8) | membership mask of the group's 8 genomes. The hash depends on x
 * alone, so the engines of a sharded run sample the same k-mers: a FracMinHash sketch whose pairwise
 * intersection / union counts give the Jaccard indices and Mash distances of genome_dist.tsv — what rules
 * mash_sample / mash_triangle compute with `mash sketch -s 10000` + `mash triangle -C -E`
 * (panagram/workflow/Snakefile:124-149; read by figs.py:50-59). HOST buffers of `cap` entries; *n_out = entries the
 * sample holds (PK_ENOMEM when that exceeds cap: call again with more room). Reads the group tables. */
int pk_engine_sample_kmers(pk_engine *e, uint32_t hmax, uint64_t *keys, uint32_t *tags, uint64_t cap, uint64_t *n_out);

/* ---- the hot path, host buffers ---------------------------------------------
 * pk_bin_len: the bin-length rule of KMCdb::write_bits (cpp/anchor.cpp:114-118) /
 *   Genome.bin_bitsum (index.py:1169-1172). 0 when nkmers < min_bin_count.
 * pk_anchor_chrom: replaces KMCdb::write_bits (cpp/anchor.cpp:112-195) and
 *   Genome._write_bitmap + the per-chromosome reductions of Genome.run_anchor
 *   (index.py:949-969,1044-1051) for ONE chromosome, with the
 *   CKMCFile::GetCountersForRead calls (kmc_file.cpp:873-1027) inside it.
 *   Equivalent to pk_anchor_genome with one chromosome.
 *     ascii      [len]                          chromosome bytes, any case
 *     bitmap1    [nkmers * row_bytes]           step-1 rows, nkmers = len-k+1
 *     bitmap_low [ceil(nkmers/step) * row_bytes] rows with p % lowres_step == 0
 *     bin_hist   [nbins * (N_local+1)]          per-bin histogram of popcount(row),
 *                                               nbins = ceil(nkmers/binlen)
 *     col_sums   [N_local]  (+=)                per-genome set-bit counts (index.py:1051)
 *   row_bytes = ceil(N_local/8) where N_local = genome_end - genome_begin.
 *   Any output pointer may be NULL to skip it. len < k: returns PK_OK with
 *   *nkmers_out = 0 and writes nothing (GetCountersForRead clears and returns
 *   false, kmc_file.cpp:878-882). nkmers < min_bin_count (binlen 0, a division
 *   by zero in the reference, cpp/anchor.cpp:116-120): bitmaps and col_sums are
 *   produced, bin_hist must be NULL or PK_EINVAL is returned. */
uint64_t pk_bin_len(const pk_config *cfg, uint64_t nkmers);
/* pk_anchor_genome: all chromosomes of one anchor in ONE batch (what KMCdb::anchor_fasta,
 *   cpp/anchor.cpp:37-109, does chromosome by chromosome). Batching matters: every probe
 *   launch streams each genome table through L2 once, so its cost is amortised over all
 *   positions of the batch. Arrays are indexed by chromosome; per-chromosome outputs have
 *   the sizes listed above; any array (or entry) may be NULL. col_sums accumulates over the
 *   whole genome; nkmers_out[c] receives len-k+1 (0 when lens[c] < k). */
int pk_anchor_genome(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                     uint8_t *const *bitmap1, uint8_t *const *bitmap_low, uint64_t *const *bin_hist,
                     uint64_t *col_sums, uint64_t *nkmers_out);
int pk_anchor_chrom(pk_engine *e, const char *ascii, uint64_t len,
                    uint8_t *bitmap1, uint8_t *bitmap_low,
                    uint64_t *bin_hist, uint64_t *col_sums, uint64_t *nkmers_out);

/* The rank-local half of the genome-sharded path (cpp/anchor.cpp has no counterpart: it holds every
 * genome in one process). pk_anchor_layout: the concatenated ("cat") row numbering one call uses — chromosome c's
 * k-mer p is row cat_off[c] + p; returns the number of rows a plane must hold. pk_anchor_genome_plane: H2D + pack +
 * probe of all chromosomes, pipelined exactly as pk_anchor_genome, with the shard's row bytes written to the
 * caller-owned DEVICE plane [plane_rows][row_stride] (e.g. from pk_device_alloc, IPC-exported to the peers;
 * row_stride >= ceil(N_local/8): the plane width all ranks share, a narrow last shard leaves its padding bytes
 * untouched); rows between chromosomes are zero or unwritten. Complete when the call returns. */
uint64_t pk_anchor_layout(uint32_t n_chroms, const uint64_t *lens, uint64_t *cat_off);
int pk_anchor_genome_plane(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                           void *d_plane, uint64_t plane_rows, uint32_t row_stride, uint64_t *nkmers_out);

/* Replaces CKMCFile::GetCountersForRead as the anchoring path calls it on
 * bitvec database `dbi` (cpp/anchor.cpp:148; index.py:932-938):
 * counters[p] bit j = presence in genome 32*dbi + j. Returns PK_OK and
 * *n_out = len-k+1, or *n_out = 0 when len < k. Genomes outside this engine's
 * shard read as 0. */
int pk_get_counters_for_read(pk_engine *e, uint32_t dbi, const char *read, uint64_t len,
                             uint32_t *counters, uint64_t *n_out);

/* pinned host memory for the buffers above (optional; pageable memory works,
 * but copies then do not overlap with kernels) */
int pk_host_alloc(void **out, size_t bytes);
int pk_host_free(void *p);

/* ---- the hot path, device buffers (building blocks for multi-GPU hosts) ------
 * All pointers are device pointers on the engine's device; `stream` is a
 * cudaStream_t passed as void* (NULL = the engine's own stream). Nothing
 * synchronises; the caller orders work on `stream`.
 *
 * pk_pack_device:   ASCII -> 2-bit words (32 bases per uint64, first base in the
 *                   two most significant bits) + invalid-base mask (1 bit per base,
 *                   LSB first, uint32 per 32 bases). d_words/d_mask need
 *                   pk_packed_words(len) uint64 / uint32 elements.
 * pk_probe_device:  rows for positions [p0, p0+n) of the packed sequence into
 *                   d_rows + (p - p0) * row_stride + col_offset, writing
 *                   ceil(N_local/8) bytes per row.
 * pk_reduce_device: per-bin popcount histogram (+=, N_cols+1 uint64 per bin,
 *                   bin = (p_first + i) / binlen), per-column sums (+=) and the
 *                   low-res rows (rows with p = p_first+i, p % step == 0, written at
 *                   index p/step - ceil(p_first/step)) of n full rows.
 * pk_interleave_device: [R][n][w] column planes (an all-gather of per-rank rows)
 *                   -> [n][row_stride] rows, plane r at byte offset r*w. */
uint64_t pk_packed_words(uint64_t len);
int pk_pack_device(pk_engine *e, const void *d_ascii, uint64_t len, void *d_words, void *d_mask, void *stream);
int pk_probe_device(pk_engine *e, const void *d_words, const void *d_mask, uint64_t p0, uint64_t n,
                    void *d_rows, uint32_t row_stride, uint32_t col_offset, void *stream);
int pk_reduce_device(pk_engine *e, const void *d_rows, uint32_t row_stride, uint32_t n_cols, uint64_t p_first,
                     uint64_t n, uint64_t binlen, void *d_bin_hist, void *d_col_sums,
                     void *d_rows_low, uint32_t lowres_step, void *stream);
int pk_interleave_device(pk_engine *e, const void *d_planes, uint32_t n_ranks, uint64_t n, uint32_t w,
                         void *d_rows, uint32_t row_stride, void *stream);

/* ---- peer memory: the fused exchange step of the genome-sharded path ------------
 * One process per GPU. Each rank allocates its plane ([n][w] bytes: its shard's columns of every row) with
 * pk_device_alloc, exports it (cudaIpc handle, 64 bytes, exchanged by the host over any channel), and maps
 * its peers' planes with pk_ipc_open. pk_gather_interleave_device then assembles full rows
 *   rows[i][r*w .. r*w+w) = planes[r][i][0..w)
 * in ONE kernel that reads the peers' planes in place over NVLink (coalesced 16-byte loads), transposes
 * through shared memory and writes whole rows: no NCCL all-gather, no [R][n][w] staging buffer, no separate
 * interleave pass. d_planes is a HOST array of n_ranks device pointers (own plane included, any order the
 * caller wants as column order). The caller orders the ranks (a stream-ordered barrier before the call,
 * another before the planes are overwritten). n_ranks <= 16. */
int pk_device_alloc(pk_engine *e, void **d_ptr, size_t bytes);   /* zero-filled */
int pk_device_free(pk_engine *e, void *d_ptr);
int pk_ipc_export(pk_engine *e, const void *d_ptr, uint8_t handle[64]);
int pk_ipc_open(pk_engine *e, const uint8_t handle[64], void **d_ptr);
int pk_ipc_close(pk_engine *e, void *d_ptr);
int pk_gather_interleave_device(pk_engine *e, const void *const *d_planes, uint32_t n_ranks, uint64_t n, uint32_t w,
                                void *d_rows, uint32_t row_stride, void *stream);

/* Position-split exchange: instead of every rank assembling every row (an all-gather: R times the NVLink and
 * HBM-write volume anyone needs), rank r assembles only ITS slice of the output rows — the slice it then reduces,
 * compresses and stores — reading that slice of all R planes in place over NVLink:
 *   rows[seg.dst_row + i][q*w .. q*w+w) = planes[q][seg.src_row + i][0..w)     0 <= i < seg.n_rows, for every segment
 * A slice is a list of segments because the planes are numbered like the concatenated anchor (chromosomes
 * separated by gap rows) while the output is the bitmap stream (chromosomes back to back, cpp/anchor.cpp:167).
 * plane_rows = rows allocated in every plane; row_bytes = bytes of a full row (ceil(N/8): a narrow last shard is
 * clipped). One kernel, no shared memory, every byte crosses NVLink once. Segment lists are cached by content:
 * repeating the previous call's list costs no upload. */
typedef struct pk_segment { uint64_t src_row, n_rows, dst_row; } pk_segment;
int pk_gather_slice_device(pk_engine *e, const void *const *d_planes, uint32_t n_ranks, uint64_t plane_rows, uint32_t w,
                           const pk_segment *segs, uint32_t n_segs, void *d_rows, uint32_t row_stride, uint32_t row_bytes,
                           void *stream);

/* ---- BGZF output on the GPU ------------------------------------------------------
 * Replaces bgzf_open/bgzf_index_build_init/bgzf_write/bgzf_index_dump/bgzf_close as KMCdb::anchor_fasta and
 * write_bits use them (cpp/anchor.cpp:46-54,102-106,167,177; Python path: bgzip.BGZipWriter + `bgzip -rI`,
 * panagram/index.py:1035-1037,1089-1094): device bytes in, the complete image of the .gz (BGZF: gzip members of
 * <= 0xff00 payload bytes + the 28-byte EOF member) and of its .gzi (uint64 n, then n x (compressed offset,
 * uncompressed offset) of every member after the first — what Genome.load_bgz_blocks reads, index.py:793-799)
 * out. The deflate streams differ from zlib's (one warp per member, fixed Huffman codes, matches at
 * `match_dist` = the row width only); the DECOMPRESSED bytes, which is what parity is defined on, are identical.
 *   pk_bgzf_bound / pk_bgzf_gzi_bound: capacity the two images need for n_bytes of payload.
 *   pk_bgzf_compress_device: d_gz / d_gzi device buffers of at least those sizes; d_totals: 2 x uint64 on
 *     the device = bytes of the .gz image, bytes of the .gzi image. Asynchronous on `stream`.
 *   pk_anchor_genome_bgzf: pk_anchor_genome with the step-1 and low-res bitmaps delivered as file images
 *     (all chromosomes of the anchor form ONE stream, as in the reference) into caller-owned HOST buffers
 *     gz[2] / gzi[2] (index 0: step 1, index 1: lowres_step) of capacities gz_cap[2] / gzi_cap[2]
 *     (>= the bounds above for sum(nkmers) * row_bytes and sum(ceil(nkmers/step)) * row_bytes);
 *     sizes[4] receives gz bytes, gzi bytes (step 1), gz bytes, gzi bytes (low-res). */
uint64_t pk_bgzf_bound(uint64_t n_bytes);
uint64_t pk_bgzf_gzi_bound(uint64_t n_bytes);
int pk_bgzf_compress_device(pk_engine *e, const void *d_in, uint64_t n_bytes, uint32_t match_dist, void *d_gz,
                            void *d_gzi, void *d_totals, void *stream);
int pk_anchor_genome_bgzf(pk_engine *e, uint32_t n_chroms, const char *const *seqs, const uint64_t *lens,
                          uint8_t *const *gz, const uint64_t *gz_cap, uint8_t *const *gzi, const uint64_t *gzi_cap,
                          uint64_t *sizes, uint64_t *const *bin_hist, uint64_t *col_sums, uint64_t *nkmers_out);

/* Pair-count bins: per bin of bin_positions consecutive positions of one chromosome (rows_per_bin =
 * ceil(bin_positions / lowres_step) low-res rows), the number of low-res rows in which each genome's bit is set —
 * Index.bitmap_to_paircount_bins (panagram/index.py:454-459), the input of chrom_umaps.csv / genome_umap.csv
 * (Genome.write_umaps, index.py:1107-1131). pk_paircount_bins_device: any device rows, counts[bins][n_cols] uint32 on the
 * device. pk_anchor_paircount_bins: the anchor the last pk_anchor_genome / pk_anchor_genome_bgzf call processed, from
 * its low-res rows still resident on the device, into HOST counts[sum_c bins_c][N_local], chromosome after chromosome
 * (nkmers = that call's nkmers_out; PK_ESTATE when it does not match). */
int pk_paircount_bins_device(pk_engine *e, const void *d_rows_low, uint32_t row_stride, uint32_t n_cols, uint64_t n_rows,
                             uint32_t rows_per_bin, void *d_counts, void *stream);
int pk_anchor_paircount_bins(pk_engine *e, uint32_t n_chroms, const uint64_t *nkmers, uint32_t bin_positions, uint32_t *counts);

/* Tuning knobs of the partitioned probe (per engine; no reference counterpart). Results never depend on
 * them; tests run the parity suite under several settings. name = "k3_window" (1: probe out of table
 * windows staged in shared memory by TMA bulk copies when they fit, 0: always probe through L1/L2),
 * "k3w_variant" (-1 auto, or a kernel variant index), "k3w_group" (0 = by window size, or 1, 2, 4 genomes per window group; two groups of windows are staged per block),
 * "k3_variant" (-1 auto; variant of the L1/L2 kernel), "l2_prefetch" (0/1), "group_tables" (0/1;
 * takes effect at the next pk_engine_finalize), "group_only" (0/1: free a group's per-genome tables once its
 * group table is built; see pk_engine_seal_group), "group_g32" (slot format of the group tables: 0 = 64-bit slots,
 * 1 = 32-bit slots [membership mask 8 | key remainder 20 | displacement 4] when k <= 23 and the table has at least
 * 2^(2k-20) buckets, which any genome-scale table has — default; 2 = 32-bit slots wherever k allows, small tables
 * padded to that size), "unpermute" (0/1: applies to
 * scratch allocated afterwards), "e2e_batches" (1..8: batches of whole chromosomes per pk_anchor_genome call;
 * copies of one batch overlap the kernels of the other), "e2e_batch_min" (positions from which a genome is
 * split into batches; default 32 Mi), "e2e_front_small" (0/1: which batch an evenly split chromosome joins),
 * "k3_l2" (one 32-bit-slot group table, one-byte rows: 0 = second partition pass + window kernels, v >= 1 = no second
 * pass, the coarse regions are probed through L2 by block shape v; default 7 = 256 threads x 4 items), "k3_lean" (the
 * same launches with k3_l2 = 0: 0 = general window kernel, v >= 1 = lean window kernel variant v), "k3w_big",
 * "compact_items", "k1_roll", "fine_out", "fine_shift", "k3_rank_atomic" (experiment knobs of those kernels; see
 * csrc/pk_internal.h). Unknown names return PK_EINVAL. */
int pk_engine_tune(pk_engine *e, const char *name, int value);

/* timing / accounting of the last pk_anchor_chrom or pk_get_counters_for_read */
typedef struct pk_stats {
    float h2d_ms, pack_ms, probe_ms, reduce_ms, d2h_ms, total_ms;
    /* CUDA-event durations of the kernels of the last partitioned probe launch, on the stream
     * they ran on: K1 partition_seq, K2 partition_fine, K3 probe_part, K3 over the spill list,
     * K4 unpermute */
    float k_partition_ms, k_fine_ms, k_probe_ms, k_spill_ms, k_unpermute_ms;
    float k_probe_window;    /* K3 of the last launch: 5 probe_g32l2_kernel (coarse regions through L2, no K2), 4 probe_g32c_kernel (lean form, one 32-bit-slot group table), 2 probe_win_kernel on group tables, 1 on per-genome tables, 3 items_group_kernel, 0 probe_part_kernel */
    uint64_t positions, probes, probe_launches, kernel_launches;
} pk_stats;
int pk_engine_stats(const pk_engine *e, pk_stats *out);

#ifdef __cplusplus
}
#endif
#endif
