// Internal declarations shared by the translation units of libpkanchor.so.
#ifndef PK_INTERNAL_H
#define PK_INTERNAL_H
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pk_anchor.h"

void pk_set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

// ---- host KMC reader (pk_kmcdb.cpp) ----
struct pk_kmcdb {
    std::string prefix;
    pk_kmcdb_info info;
    uint64_t single_lut = 0;        // 4^lut_prefix_length
    std::vector<uint64_t> lut;      // flattened LUT(s) + guard = total_kmers
    uint32_t suf_size = 0, rec_size = 0;
    FILE *suf = nullptr;
};
int pk_kmcdb_open_impl(const char *prefix, pk_kmcdb **out);
bool pk_kmcdb_read_records(pk_kmcdb *db, uint64_t first, uint64_t count, uint8_t *dst);
void pk_kmcdb_close_impl(pk_kmcdb *db);

// ---- device-side table descriptor ----
// One bucketed open-addressing table per genome: n_buckets buckets of 4 x uint64
// slots (32 B = one DRAM sector). Empty slot = ~0 (never a canonical k-mer for
// k <= 32: T^k canonicalises to A^k = 0).
struct PkTable {
    unsigned long long *slots;
    uint32_t n_buckets;
    uint32_t fmt;               // group tables: 0 = 64-bit slots (G64), 1 = 32-bit slots (G32, pk_device.cuh); 0 for per-genome tables
};

#define PK_EMPTY 0xFFFFFFFFFFFFFFFFull

// Slot formats. S64: 4 x uint64 whole k-mers per bucket (any k <= 32). S32: 8 x uint32 per bucket, each
// (low 28 key bits << 4) | displacement-from-home (0..14; 15 only occurs in the all-ones EMPTY slot); the
// other 2k-28 key bits are implied by the home bucket (pk_key_hash). Used when 2k-28 <= 20 (k <= 24):
// half the bytes per key, twice the slots per sector. A key that finds 15 consecutive full buckets goes to
// a small engine-wide stash (S64-style set of (genome << 48 | k-mer)), consulted only at the end of a
// maximal walk-on.
#define PK_FMT_S64 0
#define PK_FMT_S32 1
#define PK_EMPTY32 0xFFFFFFFFu
#define PK_S32_REM_BITS 28
#define PK_S32_REM_MASK 0x0FFFFFFFu
#define PK_S32_DISP_BITS 4
#define PK_S32_MAX_DISP 14u
#define PK_S32_MAX_EB 20u
#define PK_STASH_SLOTS 65536u
struct PkKeySpec {
    uint32_t k;
    uint32_t fmt;       // PK_FMT_*
    uint32_t eb;        // S32: key bits embedded in the bucket index = max(0, 2k - 28)
    uint32_t ghash;     // != 0: this launch hashes positions with pk_g32_hash (its group tables are G32)
    unsigned long long *stash;      // [PK_STASH_SLOTS], EMPTY-initialised (S32 only)
    unsigned int *stash_n;          // number of stashed keys
};

// ---- kernel launchers (pk_kernels.cu); all asynchronous on `stream` ----
typedef struct CUstream_st *pk_stream_t;
void pk_launch_fill_empty(unsigned long long *slots, uint64_t n_slots, pk_stream_t s);
void pk_launch_pack(const uint8_t *d_ascii, uint64_t len, uint64_t n_words, uint64_t *d_words, uint32_t *d_mask, pk_stream_t s);
// insert every valid canonical k-mer of positions [0, n) of a packed sequence into table t
void pk_launch_insert_seq(const uint64_t *d_words, const uint32_t *d_mask, uint64_t n, PkKeySpec ks, PkTable t, uint32_t g_local,
                          unsigned long long *d_counters /*[3]: inserted, overflow, fail*/, pk_stream_t s);
void pk_launch_insert_keys(const uint64_t *d_keys, uint64_t n, PkKeySpec ks, PkTable t, uint32_t g_local,
                           unsigned long long *d_counters, pk_stream_t s);
// decode KMC suffix records [rec0, rec0+n) and insert. tables: one PkTable* array on device
// (local genomes); mode 0: all into tables[local_genome]; mode 1 (bitvec): counter bit j ->
// global genome first_genome + j, inserted when inside [gbegin, gend).
struct PkDecodeArgs {
    const uint8_t *d_recs;      // n records, rec_size bytes each
    uint64_t rec0, n;
    const uint64_t *d_lut;      // n_lut_slots + 1 entries
    uint64_t n_lut_slots, single_lut;
    uint32_t suf_size, counter_size, rec_size;
    PkKeySpec ks;
    uint64_t min_count, max_count;
    int bitvec;
    uint32_t first_genome, gbegin, gend;
    const PkTable *d_tables;    // indexed by local genome
    uint32_t local_genome;      // mode 0
    unsigned long long *d_counters;   // [3 * n_local]: inserted, overflow, fail per local genome
};
void pk_launch_decode_insert(const PkDecodeArgs &a, pk_stream_t s);
void pk_launch_count_bits(const uint8_t *d_recs, uint64_t n, uint32_t rec_size, uint32_t suf_size, uint32_t counter_size, uint64_t min_count,
                          uint64_t max_count, unsigned long long *d_bits, pk_stream_t s);
// merge a per-genome table into its group table (genome bit `bit`); see pk_kernels.cu
void pk_launch_union_merge(PkTable src, uint32_t n_src_buckets, uint32_t hshift, PkKeySpec ks, PkTable dst, uint32_t bit, uint32_t g_local,
                           int use_stash, unsigned long long *d_counters /*[4]*/, pk_stream_t s);
void pk_launch_union_merge_stash(PkKeySpec ks, PkTable dst, uint32_t g0, uint32_t ng, unsigned long long *d_counters, pk_stream_t s);
void pk_launch_sample_group(PkTable t, PkKeySpec ks, uint32_t group, uint32_t hmax, int take_all, unsigned long long *d_keys, uint32_t *d_tags,
                            uint64_t cap, unsigned long long *d_n, pk_stream_t s);
void pk_launch_sample_stash(PkKeySpec ks, int g32, uint32_t hmax, int take_all, unsigned long long *d_keys, uint32_t *d_tags, uint64_t cap,
                            unsigned long long *d_n, pk_stream_t s);
void pk_launch_probe(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, PkKeySpec ks,
                     const PkTable *d_tables, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride,
                     uint32_t col_offset, pk_stream_t s);
void pk_launch_probe_group(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, PkKeySpec ks,
                           const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, pk_stream_t s);
// list4 != NULL (fine un-permute lists, one-byte rows): results are appended as 4-byte items instead of written as rows
void pk_launch_items_group(const void *d_buf, const uint32_t *d_counts, const unsigned long long *d_flat_total, uint32_t n_regions, uint64_t cap,
                           const uint64_t *d_words, uint64_t p0, PkKeySpec ks, const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows,
                           uint32_t row_stride, uint32_t col_offset, uint32_t *list4, uint32_t *list_cursor, uint32_t cursor_stride,
                           uint32_t out_shift, int compact, pk_stream_t s);
void pk_launch_reduce(const uint8_t *d_rows, uint32_t row_stride, uint32_t n_cols, uint64_t p_first, uint64_t n,
                      uint64_t binlen, unsigned long long *d_bin_hist, unsigned long long *d_col_sums,
                      uint8_t *d_rows_low, uint32_t step, pk_stream_t s);
void pk_launch_paircount_bins(const uint8_t *d_rows, uint32_t row_stride, uint32_t n_cols, uint64_t n_rows, uint32_t rows_per_bin,
                              uint32_t *d_counts, pk_stream_t s);
void pk_launch_interleave(const uint8_t *d_planes, uint32_t n_ranks, uint64_t n, uint32_t w, uint8_t *d_rows,
                          uint32_t row_stride, pk_stream_t s);
int pk_launch_gather_interleave(const void *const *planes, uint32_t n_ranks, uint64_t n, uint32_t w, uint8_t *d_rows,
                                uint32_t row_stride, pk_stream_t s);
int pk_launch_gather_slice(const void *const *planes, uint32_t n_ranks, uint64_t plane_rows, uint32_t w, const void *d_segs,
                           uint32_t n_segs, uint64_t n_chunks, uint8_t *d_rows, uint32_t row_stride, uint32_t row_bytes, pk_stream_t s);
int pk_launch_gather_slice_dst(const void *const *planes, uint32_t n_ranks, uint64_t plane_rows, uint32_t w, const void *d_segs,
                               uint32_t n_segs, uint64_t total_rows, uint8_t *d_rows, pk_stream_t s);
void pk_launch_rows_to_u32(const uint8_t *d_rows, uint32_t row_stride, uint32_t byte_off, uint32_t n_bytes,
                           uint32_t bit_mask, uint64_t n, uint32_t *d_out, pk_stream_t s);

// ---- on-GPU BGZF writer (pk_bgzf.cu) ----
#define PK_BGZF_TABLE_WORDS (1024 + 17 * 32 + 256)      // CRC slicing-by-4 tables + "append 2^j zero bytes" operators + literal codes
#define PK_BGZF_PAYLOAD 0xFF00ull              // payload bytes per BGZF member (PKZ_PAYLOAD)
uint64_t pk_bgzf_blocks_impl(uint64_t n);
uint64_t pk_bgzf_bound_impl(uint64_t n);
uint64_t pk_bgzf_gzi_bound_impl(uint64_t n);
uint64_t pk_bgzf_scratch_bytes(uint64_t n);
void pk_bgzf_tables_host(uint32_t *dst /*[PK_BGZF_TABLE_WORDS]*/);
void pk_launch_bgzf_encode(const uint8_t *d_in, uint64_t n, uint32_t dist, uint64_t m0, uint64_t m1, uint8_t *d_scratch,
                           const uint32_t *d_tables, pk_stream_t s);
void pk_launch_bgzf_finish(const uint8_t *d_in, uint64_t n, uint8_t *d_out, unsigned long long *d_gzi, unsigned long long *d_totals,
                           uint8_t *d_scratch, pk_stream_t s);
void pk_launch_bgzf(const uint8_t *d_in, uint64_t n, uint32_t dist, uint8_t *d_out, unsigned long long *d_gzi,
                    unsigned long long *d_totals, uint8_t *d_scratch, const uint32_t *d_tables, pk_stream_t s);

// ---- partitioned probe (pk_partition.cu) ----
// positions per partitioned launch: 2^18 fine partitions (9 + 9 radix bits, PT_MAXB = 512 digits per pass) of a
// mean fill of 640 = 5/6 of the K3 block capacity of 768. A longer launch would overflow the fixed-capacity
// regions en masse into the spill list, so longer batches are cut into sub-launches.
#define PK_PART_MAX_N (640ull << 18)
struct PkPartPlan {
    uint32_t pb1, pb2, cap1, cap2, n_regions1, n_regions2;
    uint64_t buf1_items, buf2_items, spill_items;   // 8-byte (hash, pos) items
    uint32_t out_shift, out_bins;                   // un-permute lists: out_bins bins of 2^out_shift positions
    uint64_t out_bytes, fine_rows;                  // bytes of the un-permute lists; rows of the launch (fine mode)
    uint32_t compact;                               // 1: K1/K2 emit compact items that carry hash + key remainder (pk_partition.cu)
    int wbig;                                       // >= 0: the window-kernel variant (large blocks) this plan's capacity belongs to
    uint32_t out_fine;                              // 1: one-byte rows — 4-byte (position in bin, bits) items in <= 512 bins, un-permuted
                                                    //    through shared-memory slices of the bitmap (unpermute_slice_kernel)
};
// tuning state of the partitioned probe: per engine (two engines of one process — thread per GPU — do not share it)
struct PkPartTune {
    int variant = -1;       // L1/L2 kernel variant, -1 auto
    int window = 1;         // 0: never use the window kernels
    int wvariant = -1;      // window kernel variant, -1 auto
    int wgroup = 0;         // genomes per window group (2 * group stage buffers); 0 = by window size
    int rank_atomic = 1;    // window kernel: shared-memory-atomic ranking of the results (1) or warp match_any (0)
    int wbig = 6;           // window-kernel variant for one-byte rows out of 32-bit-slot group tables (>= 4: the large-block variants)
    int compact = 1;        // compact items on the one-byte-row / 32-bit-slot path (K3 does not read the sequence)
    int lean = 1;           // K3 on ONE 32-bit-slot group table with compact items and fine bins: the lean kernel, variant lean - 1 (0: the general window kernel)
    int k3_l2 = 7;          // > 0: no K2; K3 probes the coarse regions through L2 (probe_g32l2_kernel variant k3_l2 - 1) where one
                            // 32-bit-slot group table, compact items and fine bins allow it; 0: K2 + the window kernels
    int k1_roll = 1;        // K1: rolling k-mers over 16 consecutive positions per thread (0: re-extract every window)
    int fine_shift = 0;     // fine mode: log2 of the positions per bin, 0 = the smallest that gives <= 512 bins
    int fine_out = 1;       // one-byte rows out of group tables: fine position bins + shared-memory-slice un-permute
    int last_window = 0;    // K3 of the last launch: 2 window kernel on group tables, 1 on per-genome tables, 3 L1/L2 on group tables, 4 lean window kernel on one 32-bit-slot group table, 5 coarse regions through L2 (no K2), 0 L1/L2 kernel
};
struct PkPartScratch {
    const PkPartTune *tune;                         // never NULL on the engine's paths
    int *last_window;
    void *buf1, *buf2, *spill;
    uint32_t *cursor1, *cursor2;
    unsigned long long *spill_cursor;
    uint32_t *err;
    void *out_list;                                 // NULL: probe_part scatters row bytes itself
    uint32_t *out_cursor;
    uint64_t out_bytes;
};
uint32_t pk_part_obins(void);
uint32_t pk_part_ocursor_words(void);
int pk_part_n_variants(void);
int pk_part_n_wvariants(void);
int pk_part_n_lvariants(void);
int pk_part_n_gvariants(void);
// fine_out: 0 no, 1 one-byte rows out of group tables (fine bins + slice un-permute), 2 ... with 32-bit slots (larger K3 blocks)
void pk_part_plan(uint64_t n, const PkPartTune &tune, PkPartPlan *pl, uint32_t n_local = 1, int fine_out = 0, uint32_t k = 0);
uint32_t pk_part_max_fine_bins(void);
void pk_part_begin(uint32_t n_local, const PkPartPlan &pl, const PkPartScratch &sc, pk_stream_t s);
void pk_part_append(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t off, uint64_t n, PkKeySpec ks,
                    uint32_t n_local, uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl,
                    const PkPartScratch &sc, pk_stream_t s);
void pk_part_probe(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, PkKeySpec ks, const PkTable *h_tables,
                   const PkTable *h_utables /*group tables, one per 8 local genomes, or NULL*/, const PkTable *d_utables /*the same on the device*/, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl,
                   const PkPartScratch &sc, int prefetch, pk_stream_t s, struct CUevent_st **evs);
void pk_part_unpermute(uint32_t bin0, uint32_t bin1, uint32_t n_local, uint8_t *d_rows, uint32_t row_stride,
                       uint32_t col_offset, const PkPartPlan &pl, const PkPartScratch &sc, pk_stream_t s);
int pk_launch_probe_partitioned(const uint64_t *d_words, const uint32_t *d_mask, uint64_t p0, uint64_t n, PkKeySpec ks,
                                const PkTable *d_tables, const PkTable *h_tables, const PkTable *h_utables, const PkTable *d_utables, uint32_t n_local, uint8_t *d_rows,
                                uint32_t row_stride, uint32_t col_offset, const PkPartPlan &pl, const PkPartScratch &sc,
                                int prefetch, pk_stream_t s, struct CUevent_st **evs /*6 events or NULL*/);
#endif
