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
        self.index: list[tuple[int, int]] = []   # start of every block after the first

    def _emit(self, view: memoryview):
        chunks = [bytes(view[o:o + BGZF_PAYLOAD]) for o in range(0, len(view), BGZF_PAYLOAD)]
        for payload, blk in zip(chunks, self.pool.map(lambda c: _bgzf_block(c, self.level), chunks)):
            self.fh.write(blk)
            self.coff += len(blk)
            self.uoff += len(payload)
            self.index.append((self.coff, self.uoff))

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
