"""TEST INFRASTRUCTURE ONLY — NOT PART OF THE PRODUCT.

Python face of the CPU oracle: ctypes bindings to ``oracle/libpkoracle.so``
(the plain-C restatement in ``anchor_oracle.c``; parity pinning described in its
header) plus a brute-force, set-based restatement for tiny cases in the style of
``KMC/tests/py_kmc_api/test_py_kmc_file.py:174-197``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module; ``panagram_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libpkoracle.so"
    src = _HERE / "anchor_oracle.c"
    if force or not so.exists() or (src.exists() and so.stat().st_mtime < src.stat().st_mtime):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall",
                               "-o", str(so), str(src)])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        vp, u64p, u32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        L.pko_db_open.restype = vp
        L.pko_db_open.argtypes = [C.c_char_p]
        L.pko_db_close.argtypes = [vp]
        L.pko_db_from_sorted.restype = vp
        L.pko_db_from_sorted.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, C.c_uint64]
        L.pko_db_info.argtypes = [vp, u64p]
        L.pko_get_counters_for_read.restype = C.c_int
        L.pko_get_counters_for_read.argtypes = [vp, C.c_char_p, C.c_uint64, u32p]
        L.pko_db_list.restype = C.c_uint64
        L.pko_db_list.argtypes = [vp, u64p, u32p]
        L.pko_kmers_of_seq.restype = C.c_uint64
        L.pko_kmers_of_seq.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, u64p]
        L.pko_binlen.restype = C.c_uint64
        L.pko_binlen.argtypes = [C.c_uint64]
        L.pko_anchor_chrom.restype = C.c_uint64
        L.pko_anchor_chrom.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_char_p, C.c_uint64,
                                       u8p, u8p, u64p, u64p, u64p, u64p, u64p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleDB:
    """One KMC database opened the way CKMCFile::OpenForRA does (kmc_file.cpp:25-53)."""

    def __init__(self, handle):
        if not handle:
            raise OSError("oracle: cannot open KMC database")
        self.h = C.c_void_p(handle)
        info = np.zeros(10, dtype=np.uint64)
        lib().pko_db_info(self.h, _p(info, C.c_uint64))
        (self.k, self.counter_size, self.lut_prefix_len, self.signature_len, self.min_count,
         self.max_count, self.total_kmers, both, self.version, self.mode) = (int(x) for x in info)
        self.both_strands = bool(both)

    @classmethod
    def open(cls, prefix) -> "OracleDB":
        return cls(lib().pko_db_open(str(prefix).encode()))

    @classmethod
    def from_kmers(cls, k: int, kmers, counters=None, lut_prefix_len: int | None = None,
                   counter_size: int = 4) -> "OracleDB":
        """In-memory KMC1 canonical DB from canonical k-mer integers."""
        kmers = np.asarray(kmers, dtype=np.uint64)
        if counters is None:
            counters = np.ones(len(kmers), dtype=np.uint32)
        counters = np.asarray(counters, dtype=np.uint32)
        order = np.argsort(kmers, kind="stable")
        kmers, counters = np.ascontiguousarray(kmers[order]), np.ascontiguousarray(counters[order])
        if lut_prefix_len is None:
            lut_prefix_len = next(l for l in range(1, 5) if (k - l) % 4 == 0)
        h = lib().pko_db_from_sorted(k, lut_prefix_len, counter_size, _p(kmers, C.c_uint64),
                                     _p(counters, C.c_uint32), len(kmers))
        return cls(h)

    def close(self):
        if self.h:
            lib().pko_db_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_counters_for_read(self, read: bytes) -> np.ndarray | None:
        """GetCountersForRead (kmc_file.cpp:873-900): None when len < k."""
        if isinstance(read, str):
            read = read.encode()
        n = len(read) - self.k + 1
        if n <= 0:
            return None
        out = np.empty(n, dtype=np.uint32)
        ok = lib().pko_get_counters_for_read(self.h, read, len(read), _p(out, C.c_uint32))
        return out if ok else None

    def list(self) -> tuple[np.ndarray, np.ndarray]:
        kmers = np.empty(self.total_kmers, dtype=np.uint64)
        counts = np.empty(self.total_kmers, dtype=np.uint32)
        n = lib().pko_db_list(self.h, _p(kmers, C.c_uint64), _p(counts, C.c_uint32))
        return kmers[:n], counts[:n]


def canonical_kmers(seqs, k: int) -> np.ndarray:
    """Sorted distinct canonical k-mer integers of a list of sequences (bytes / uint8 arrays): the set K_g that
    `kmc -k -ci1 -fm` produces for a FASTA (pko_kmers_of_seq in C, then sort + unique). ctypes releases the GIL."""
    parts = []
    for s in seqs:
        b = s.tobytes() if isinstance(s, np.ndarray) else bytes(s)
        if len(b) < k:
            continue
        out = np.empty(len(b) - k + 1, dtype=np.uint64)
        n = lib().pko_kmers_of_seq(b, len(b), k, _p(out, C.c_uint64))
        parts.append(out[:n])
    if not parts:
        return np.zeros(0, dtype=np.uint64)
    x = np.sort(np.concatenate(parts))            # np.unique's hash path is ~50x slower than sort + compare here
    if x.size == 0:
        return x
    keep = np.empty(x.size, dtype=bool)
    keep[0] = True
    np.not_equal(x[1:], x[:-1], out=keep[1:])
    return x[keep]


def binlen(nkmers: int) -> int:
    return int(lib().pko_binlen(nkmers))


def anchor_chrom(dbs: list[OracleDB], n_genomes: int, seq: bytes, lowres_step: int = 100) -> dict:
    """cpp/anchor.cpp:112-195 for one chromosome (+ index.py:1051 column sums)."""
    assert lowres_step == 100, "the C++ reference hard-codes 100 (cpp/anchor.cpp:170)"
    k = dbs[0].k
    nbytes = (n_genomes + 7) // 8
    nk = len(seq) - k + 1
    if nk < 100:
        raise ValueError("nkmers < 100 is undefined in the reference (cpp/anchor.cpp:116-120)")
    bl = binlen(nk)
    nch = nk // bl + (nk % bl != 0)
    b1 = np.zeros(nk * nbytes, dtype=np.uint8)
    b100 = np.zeros(((nk + 99) // 100 + 1) * nbytes, dtype=np.uint8)
    starts = np.zeros(nch, dtype=np.uint64)
    hist = np.zeros(nch * (n_genomes + 1), dtype=np.uint64)
    pc = np.zeros(n_genomes, dtype=np.uint64)
    n100 = C.c_uint64(0)
    nbins = C.c_uint64(0)
    arr = (C.c_void_p * len(dbs))(*[d.h for d in dbs])
    r = lib().pko_anchor_chrom(arr, len(dbs), n_genomes, seq, len(seq), _p(b1, C.c_uint8),
                               _p(b100, C.c_uint8), C.byref(n100), _p(starts, C.c_uint64),
                               _p(hist, C.c_uint64), C.byref(nbins), _p(pc, C.c_uint64))
    if r == 2 ** 64 - 1:
        raise RuntimeError("oracle anchor failed")
    return {"nkmers": int(r), "bitmap1": b1.reshape(nk, nbytes),
            "bitmap100": b100[:n100.value * nbytes].reshape(n100.value, nbytes),
            "bin_start": starts, "bin_hist": hist.reshape(nch, n_genomes + 1), "paircounts": pc}


def parse_fasta(path) -> list[tuple[str, bytes]]:
    """cpp/anchor.cpp:74-100: name = header up to the first space; lines concatenated
    verbatim (no \\r stripping)."""
    recs, name, parts = [], None, []
    with open(path, "rb") as fh:
        for line in fh.read().split(b"\n"):
            if line[:1] == b">":
                if name is not None:
                    recs.append((name, b"".join(parts)))
                name, parts = line[1:].split(b" ")[0].decode(), []
            else:
                parts.append(line)
    if name is not None:
        recs.append((name, b"".join(parts)))
    return recs


def anchor_fasta(dbs: list[OracleDB], n_genomes: int, fasta) -> dict:
    """cpp/anchor.cpp:37-109 KMCdb::anchor_fasta: the decompressed payload of
    bitmap.1 / bitmap.100 and the exact text of chrs.tsv / bitsum.bins.tsv."""
    b1, b100 = [], []
    chrs = ["name\tid\tsize\tgene_count\n"]
    bins = ["chr\tstart" + "".join(f"\t{i}" for i in range(n_genomes + 1)) + "\n"]
    pc = np.zeros(n_genomes, dtype=np.uint64)
    for cid, (name, seq) in enumerate(parse_fasta(fasta)):
        r = anchor_chrom(dbs, n_genomes, seq)
        b1.append(r["bitmap1"].tobytes())
        b100.append(r["bitmap100"].tobytes())
        pc += r["paircounts"]
        chrs.append(f"{name}\t{cid}\t{r['nkmers']}\t0\n")
        for st, row in zip(r["bin_start"], r["bin_hist"]):
            bins.append(f"{cid}\t{int(st)}" + "".join(f"\t{int(c)}" for c in row) + "\n")
    return {"bitmap.1": b"".join(b1), "bitmap.100": b"".join(b100), "chrs.tsv": "".join(chrs),
            "bitsum.bins.tsv": "".join(bins), "paircounts": pc}


# ---- brute force (pure Python; tiny inputs only) ----------------------------

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def canon_str(kmer: bytes) -> bytes:
    """min(kmer, revcomp) as strings — test_py_kmc_file.py:181-185."""
    rc = kmer.translate(_COMP)[::-1]
    return rc if rc < kmer else kmer


def kmer_int(kmer: bytes) -> int:
    v = 0
    for ch in kmer:
        v = (v << 2) | b"ACGT".index(ch)
    return v


def kmer_set(seqs, k: int) -> set[bytes]:
    """Canonical k-mer set of sequences (K_g of SURVEY §0): windows with any byte
    outside ACGTacgt are skipped (kmc_core/splitter.cpp:44-47)."""
    out = set()
    for s in seqs:
        s = s.upper() if s.isascii() else s
        for i in range(len(s) - k + 1):
            w = s[i:i + k]
            if all(c in b"ACGT" for c in w):
                out.add(canon_str(w))
    return out


def brute_rows(seq: bytes, k: int, sets: list[set[bytes]]) -> np.ndarray:
    """bitmap rows [len-k+1, ceil(N/8)] by direct set membership."""
    n = len(seq) - k + 1
    nbytes = (len(sets) + 7) // 8
    rows = np.zeros((max(n, 0), nbytes), dtype=np.uint8)
    for p in range(n):
        w = seq[p:p + k]
        if not all(c in b"ACGTacgt" for c in w):
            continue
        c = canon_str(w.upper())
        for g, s in enumerate(sets):
            if c in s:
                rows[p, g // 8] |= 1 << (g % 8)
    return rows
