"""On-disk layout of an anchor directory, byte-compatible with what the
reference writes and reads:

  bitmap.{1,100}.gz   BGZF of raw rows            (cpp/anchor.cpp:43-54,167,177; index.py:960-963)
  bitmap.{1,100}.gzi  uint64 n + n x (coffset, uoffset)  (cpp/anchor.cpp:102-105; read at index.py:793-799)
  chrs.tsv            name id size gene_count     (cpp/anchor.cpp:66-69,84-85)
  bitsum.bins.tsv     chr start 0..N              (cpp/anchor.cpp:57-63,184-189)
  total_paircounts.csv name,count,frac            (index.py:1068-1074; required by index.py:639-641)

The BGZF writer compresses independent 0xff00-byte blocks on a thread pool
(zlib releases the GIL); parity with the reference is defined on the
decompressed bytes.
"""
from __future__ import annotations

import gzip
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

BGZF_PAYLOAD = 0xFF00
_BGZF_HDR = bytes([0x1F, 0x8B, 8, 4, 0, 0, 0, 0, 0, 0xFF, 6, 0, ord("B"), ord("C"), 2, 0])
BGZF_EOF = bytes([0x1F, 0x8B, 8, 4, 0, 0, 0, 0, 0, 0xFF, 6, 0, 0x42, 0x43, 2, 0, 0x1B, 0, 3, 0,
                  0, 0, 0, 0, 0, 0, 0, 0])


def _bgzf_block(payload: bytes, level: int) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(payload) + co.flush()
    bsize = len(body) + 26
    if bsize > 0x10000:                     # incompressible: stored block
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        body = co.compress(payload) + co.flush()
        bsize = len(body) + 26
    return b"".join((_BGZF_HDR, struct.pack("<H", bsize - 1), body,
                     struct.pack("<II", zlib.crc32(payload), len(payload))))


class BgzfWriter:
    """Streaming BGZF + .gzi writer (bgzf_open/bgzf_index_build_init/bgzf_write/
    bgzf_index_dump/bgzf_close of cpp/anchor.cpp:46-54,102-106)."""

    def __init__(self, path, level: int = 6, threads: int = 8):
        self.path = Path(path)
        self.fh = open(self.path, "wb")
        self.level = level
        self.pool = ThreadPoolExecutor(max_workers=max(1, threads))
        self.pending = bytearray()
        self.coff = 0
        self.uoff = 0
        self.index: list[tuple[int, int]] = []   # start of every block after the first (htslib's write-mode index,
                                                 # cpp/anchor.cpp:47,102: no entry for the EOF member)

    def _emit(self, view: memoryview):
        chunks = [bytes(view[o:o + BGZF_PAYLOAD]) for o in range(0, len(view), BGZF_PAYLOAD)]
        for payload, blk in zip(chunks, self.pool.map(lambda c: _bgzf_block(c, self.level), chunks)):
            if self.coff:
                self.index.append((self.coff, self.uoff))
            self.fh.write(blk)
            self.coff += len(blk)
            self.uoff += len(payload)

    def write(self, data):
        mv = memoryview(data).cast("B")
        if self.pending:
            need = BGZF_PAYLOAD - len(self.pending)
            self.pending += mv[:need]
            mv = mv[need:]
            if len(self.pending) < BGZF_PAYLOAD:
                return
            self._emit(memoryview(bytes(self.pending)))
            self.pending = bytearray()
        full = len(mv) // BGZF_PAYLOAD * BGZF_PAYLOAD
        if full:
            self._emit(mv[:full])
        self.pending += mv[full:]

    def close(self, gzi_path=None):
        if self.pending:
            self._emit(memoryview(bytes(self.pending)))
            self.pending = bytearray()
        self.fh.write(BGZF_EOF)
        self.fh.close()
        self.pool.shutdown()
        if gzi_path is not None:
            with open(gzi_path, "wb") as fh:
                fh.write(struct.pack("<Q", len(self.index)))
                fh.write(np.asarray(self.index, dtype="<u8").tobytes())


def load_bgz_blocks(gzi_path) -> np.ndarray:
    """index.py:793-799: [(compressed, uncompressed)] block starts with (0,0) prepended."""
    raw = Path(gzi_path).read_bytes()
    (n,) = struct.unpack_from("<Q", raw, 0)
    blocks = np.frombuffer(raw, dtype="<u8", count=2 * n, offset=8).reshape(n, 2)
    return np.vstack([np.zeros((1, 2), dtype="<u8"), blocks])


def query_bytes(gz_path, gzi_path, byte_start: int, length: int) -> bytes:
    """The access pattern of Genome._query_bytes (index.py:827-845): locate the BGZF
    block holding an uncompressed offset through the .gzi, decompress from there."""
    blocks = load_bgz_blocks(gzi_path)
    bi = int(np.searchsorted(blocks[:, 1], byte_start, side="right")) - 1
    coff, uoff = int(blocks[bi, 0]), int(blocks[bi, 1])
    out = bytearray()
    with open(gz_path, "rb") as fh:
        fh.seek(coff)
        skip = byte_start - uoff
        while len(out) < length:
            hdr = fh.read(18)
            if len(hdr) < 18:
                break
            (bsize,) = struct.unpack_from("<H", hdr, 16)
            body = fh.read(bsize + 1 - 18)
            data = zlib.decompress(body[:-8], -15)
            out += data[skip:]
            skip = 0
    return bytes(out[:length])


def read_bgzf(path) -> bytes:
    with gzip.open(path, "rb") as fh:
        return fh.read()


def chrs_tsv(chroms: list[tuple[str, int]]) -> str:
    """cpp/anchor.cpp:66-69,84-85: name (header up to the first space), id, size = len-k+1, 0."""
    return "name\tid\tsize\tgene_count\n" + "".join(
        f"{name}\t{i}\t{size}\t0\n" for i, (name, size) in enumerate(chroms))


def bins_tsv(n_genomes: int, per_chrom: list[tuple[int, np.ndarray]]) -> str:
    """cpp/anchor.cpp:57-63,184-189. per_chrom: [(binlen, hist[nbins, N+1])] in chromosome order."""
    lines = ["chr\tstart" + "".join(f"\t{i}" for i in range(n_genomes + 1)) + "\n"]
    for cid, (binlen, hist) in enumerate(per_chrom):
        for b, row in enumerate(hist):
            lines.append(f"{cid}\t{b * binlen}" + "".join(f"\t{int(c)}" for c in row) + "\n")
    return "".join(lines)


def paircounts_csv(names: list[str], counts: np.ndarray, anchor: str) -> str:
    """index.py:1068-1074: DataFrame{count, frac = count / count[anchor]}.to_csv() with the
    samples index named 'name'."""
    denom = float(counts[names.index(anchor)])
    lines = ["name,count,frac\n"]
    for n, c in zip(names, counts):
        c = int(c)
        if denom:
            frac = repr(c / denom)
        else:
            frac = "" if c == 0 else "inf"      # pandas: 0/0 -> NaN (empty field), c/0 -> inf
        lines.append(f"{n},{c},{frac}\n")
    return "".join(lines)


# ---- UMAP inputs / files ----------------------------------------------------------
def paircount_bins(rows_low: np.ndarray, n_genomes: int, step: int, bin_size: int) -> tuple[np.ndarray, np.ndarray]:
    """Index.bitmap_to_paircount_bins (index.py:454-459) on the low-res rows of ONE chromosome
    (row i = position i * step): per bin of `bin_size` positions the per-genome count of set bits, divided by
    the bin's maximum over genomes (0/0 -> NaN, which the caller's .fillna(0) turns into 0).
    Returns (bin starts [nbins], fractions [nbins, N])."""
    rows_low = np.asarray(rows_low, dtype=np.uint8).reshape(len(rows_low), -1)
    bits = np.unpackbits(rows_low, axis=1, bitorder="little")[:, :n_genomes]
    pos = np.arange(bits.shape[0], dtype=np.int64) * step
    bins = pos // bin_size
    first = np.flatnonzero(np.r_[True, bins[1:] != bins[:-1]]) if bits.shape[0] else np.zeros(0, dtype=np.int64)
    sums = np.add.reduceat(bits.astype(np.int64), first, axis=0) if bits.shape[0] else np.zeros((0, n_genomes), np.int64)
    mx = sums.max(axis=1, keepdims=True) if sums.size else sums
    with np.errstate(invalid="ignore", divide="ignore"):
        frac = np.where(mx > 0, sums / np.maximum(mx, 1), 0.0)
    return bins[first] * bin_size, frac


def paircount_frac(counts: np.ndarray, bin_size: int) -> tuple[np.ndarray, np.ndarray]:
    """(bin starts, fractions) from per-bin per-genome counts [nbins, N] (pk_anchor_paircount_bins): every bin divided
    by its maximum over genomes, 0/0 -> 0 (Index.bitmap_to_paircount_bins + the caller's fillna(0))."""
    counts = np.asarray(counts, dtype=np.int64).reshape(len(counts), -1)
    mx = counts.max(axis=1, keepdims=True) if counts.size else counts
    with np.errstate(invalid="ignore", divide="ignore"):
        frac = np.where(mx > 0, counts / np.maximum(mx, 1), 0.0)
    return np.arange(counts.shape[0], dtype=np.int64) * bin_size, frac


def umap_rows(chrom: str, starts: np.ndarray, frac: np.ndarray, bin_size: int, neighbors: int = 4, dist: float = 0.0,
              eps: float = 1.0, samples: int = 1) -> list[tuple]:
    """Genome.run_umap (index.py:1133-1156): UMAP(n_neighbors, min_dist, n_components=2, random_state=42) +
    DBSCAN(eps, min_samples) of the binned pair-counts -> (chrom, start, end, umap1, umap2, cluster) rows. When
    umap-learn is not installed (or the fit fails) the reference's own fallback is written: zeros (index.py:1149-1152)."""
    emb = None
    try:
        import umap  # type: ignore
        from sklearn.cluster import DBSCAN
        emb = umap.UMAP(n_neighbors=neighbors, min_dist=dist, n_components=2, random_state=42).fit_transform(frac)
        clusters = DBSCAN(eps=eps, min_samples=samples).fit_predict(emb)
    except Exception:
        emb = None
    if emb is None:
        return [(chrom, int(s), int(s) + bin_size, 0, 0, 0) for s in starts]
    return [(chrom, int(s), int(s) + bin_size, float(a), float(b), int(c)) for s, (a, b), c in zip(starts, emb, clusters)]


def umaps_csv(rows: list[tuple]) -> str:
    """chrom_umaps.csv (DataFrame.set_index("chrom").to_csv()) and genome_umap.csv (to_csv(index=False)) have the
    same text: header chrom,start,end,umap1,umap2,cluster (index.py:1128-1131)."""
    return "chrom,start,end,umap1,umap2,cluster\n" + "".join(",".join(map(str, r)) + "\n" for r in rows)


# ---- genome_dist.tsv ----------------------------------------------------------------
def kmer_sample_hash(keys: np.ndarray) -> np.ndarray:
    """hash32 of pk_engine_sample_kmers' sampling rule (pk_hash64 in csrc/pk_device.cuh), restated for the tests and for
    host-side sub-sampling."""
    x = np.asarray(keys, dtype=np.uint64).copy()
    x ^= x >> np.uint64(32)
    x *= np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(29)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    return (x >> np.uint64(32)).astype(np.uint32)


_BITS8 = ((np.arange(256)[:, None] >> np.arange(8)[None, :]) & 1).astype(np.float64)       # [mask value, bit]


def pair_counts(records: list[tuple[np.ndarray, np.ndarray, int]], n_genomes: int) -> np.ndarray:
    """[N, N] intersection sizes (diagonal: set sizes) of the genomes' k-mer samples. records = one
    (keys uint64 [n], tags uint32 [n], genome_begin) per engine as pk_engine_sample_kmers returns them: tag =
    (local group << 8) | membership mask of genomes genome_begin + 8 * group + 0..7.
    Per group of 8 genomes the keys are sorted once and equal keys merged (a k-mer can sit in the table and in the
    stash); pairs inside a group then come from the histogram of the 256 mask values, pairs across two groups from
    the 256 x 256 histogram of the mask pairs of their common keys (a sorted-merge join)."""
    groups = {}
    for keys, tags, base in records:
        keys = np.asarray(keys, dtype=np.uint64)
        tags = np.asarray(tags, dtype=np.uint32)
        for grp in np.unique(tags >> 8) if tags.size else []:
            sel = (tags >> 8) == grp
            g0 = int(base) + 8 * int(grp)
            k, m = keys[sel], (tags[sel] & 0xFF).astype(np.uint8)
            if g0 in groups:
                k, m = np.concatenate([groups[g0][0], k]), np.concatenate([groups[g0][1], m])
            groups[g0] = (k, m)
    inter = np.zeros((n_genomes, n_genomes), dtype=np.float64)
    merged = {}
    for g0, (k, m) in groups.items():
        order = np.argsort(k, kind="stable")
        k, m = k[order], m[order]
        if k.size:
            first = np.flatnonzero(np.r_[True, k[1:] != k[:-1]])
            if first.size != k.size:
                m = np.bitwise_or.reduceat(m, first)
                k = k[first]
        merged[g0] = (k, m)
        ng = min(8, n_genomes - g0)
        if ng <= 0:
            continue
        hist = np.bincount(m, minlength=256).astype(np.float64)
        inter[g0:g0 + ng, g0:g0 + ng] = ((_BITS8 * hist[:, None]).T @ _BITS8)[:ng, :ng]
    starts = sorted(merged)
    for ia, ga in enumerate(starts):
        for gb in starts[ia + 1:]:
            na, nb = min(8, n_genomes - ga), min(8, n_genomes - gb)
            if na <= 0 or nb <= 0:
                continue
            (ka, ma), (kb, mb) = merged[ga], merged[gb]
            _, ja, jb = np.intersect1d(ka, kb, assume_unique=True, return_indices=True)
            h2 = np.bincount(ma[ja].astype(np.int64) * 256 + mb[jb], minlength=65536).astype(np.float64).reshape(256, 256)
            blk = (_BITS8.T @ h2 @ _BITS8)[:na, :nb]
            inter[ga:ga + na, gb:gb + nb] = blk
            inter[gb:gb + nb, ga:ga + na] = blk.T
    return inter


def mash_distance(jaccard: float, k: int) -> float:
    """Mash distance of a Jaccard index j: -1/k * ln(2j / (1 + j)) (1 when j = 0) — Ondov et al. 2016, what `mash dist`
    and `mash triangle` print."""
    if jaccard <= 0:
        return 1.0
    if jaccard >= 1:
        return 0.0
    return max(0.0, -np.log(2.0 * jaccard / (1.0 + jaccard)) / k)


def genome_dist_tsv(names: list[str], inter: np.ndarray, k: int) -> str:
    """The edge list `mash triangle -C -E` writes (workflow/Snakefile:139-149) and make_all_genome_dend reads
    (figs.py:50-59: f, t, d, p, x per line; only d is used): for every pair i > j one line
    name_i, name_j, Mash distance, p-value, shared/total — here from the intersection / union sizes of the genomes'
    k-mer samples instead of 10000-hash MinHash sketches. p-value: the probability of >= `shared` common k-mers among
    `total` by chance — mash tests P[Binomial(total, r) >= shared] with r the expected Jaccard of two random k-mer sets
    of these sizes; here its Chernoff bound exp(-total * KL(shared/total || r)) — 0 for any pair of related genomes."""
    n = len(names)
    size = np.diag(inter)
    space = 4.0 ** k
    lines = []
    for i in range(1, n):
        for j in range(i):
            shared = inter[i, j]
            total = size[i] + size[j] - shared
            jac = shared / total if total > 0 else 0.0
            d = mash_distance(jac, k)
            px, py = 1.0 / (1.0 + space / max(size[i], 1.0)), 1.0 / (1.0 + space / max(size[j], 1.0))
            r = px * py / (px + py - px * py)
            # Chernoff bound of the binomial tail: exp(-total * KL(shared/total || r)); 1 when the overlap is at chance level
            a = shared / total if total > 0 else 0.0
            if shared <= 0 or a <= r:
                p = 1.0
            elif a >= 1.0:
                p = float(r ** total) if total < 1e4 else 0.0
            else:
                kl = a * np.log(a / r) + (1.0 - a) * np.log((1.0 - a) / (1.0 - r))
                p = float(np.exp(-min(total * kl, 745.0))) if total * kl < 745.0 else 0.0
            lines.append(f"{names[i]}\t{names[j]}\t{d:.6g}\t{p:.6g}\t{int(shared)}/{int(total)}\n")
    return "".join(lines)
