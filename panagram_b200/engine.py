"""Python host side of the anchoring engine: a thin object layer over the C ABI
(include/pk_anchor.h). Mirrors the roles of ``KMCdb`` in ``cpp/anchor.cpp:16-35``
(open the k-mer sets once, anchor many FASTAs) and of ``Genome._load_kmc`` /
``Genome._write_bitmap`` in ``panagram/index.py:847-863,949-969``.

All compute happens in libpkanchor.so on the GPU; this module only moves
buffers. numpy is used for host arrays, torch (optional) only to hand device
pointers of tensors to the device-level entry points.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import PkConfig, PkStats, PkTableStats, check


def _u8(seq) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        if seq.dtype != np.uint8:
            raise TypeError("sequence arrays must be uint8")
        return np.ascontiguousarray(seq)
    if isinstance(seq, str):
        seq = seq.encode()
    return np.frombuffer(seq, dtype=np.uint8)


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy array over page-locked host memory (pk_host_alloc)."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(_lib.lib().pk_host_alloc(C.byref(p), max(nbytes, 1)))
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    weakref.finalize(buf, _lib.lib().pk_host_free, p)
    return arr


class Engine:
    """Per-genome k-mer tables resident in HBM + the anchoring calls."""

    def __init__(self, k: int, n_genomes: int, genome_begin: int = 0, genome_end: int | None = None,
                 device: int = 0, lowres_step: int = 100, max_bin_kbp: int = 200,
                 min_bin_count: int = 100, load_factor: float = 0.5, chunk_positions: int = 0,
                 probe_mode: str | int = "auto"):
        self._L = _lib.lib()
        genome_end = n_genomes if genome_end is None else genome_end
        mode = {"auto": 0, "direct": 1, "partitioned": 2}.get(probe_mode, probe_mode)
        self.cfg = PkConfig(k, n_genomes, genome_begin, genome_end, device, lowres_step,
                            max_bin_kbp * 1000, min_bin_count, load_factor, chunk_positions, mode)
        self.k, self.n_genomes = k, n_genomes
        self.genome_begin, self.genome_end = genome_begin, genome_end
        self.n_local = genome_end - genome_begin
        self.row_bytes = (self.n_local + 7) // 8
        self.lowres_step = lowres_step
        h = C.c_void_p()
        check(self._L.pk_engine_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, self._L.pk_engine_destroy, h)

    def close(self):
        self._fin()

    # ---- table construction -------------------------------------------------
    def reserve(self, genome: int, max_keys: int):
        check(self._L.pk_engine_reserve(self._h, genome, max_keys))

    def add_kmc(self, genome: int, prefix):
        """K_g from a per-genome KMC database (kmc/{s}.count or .onehot)."""
        check(self._L.pk_engine_add_kmc(self._h, genome, str(prefix).encode()))

    def add_bitvec(self, first_genome: int, prefix):
        """A merged bitvec database (kmc/bitvec{i}): counter bit j -> genome first+j."""
        check(self._L.pk_engine_add_bitvec(self._h, first_genome, str(prefix).encode()))

    def add_keys(self, genome: int, canon_kmers):
        a = np.ascontiguousarray(canon_kmers, dtype=np.uint64)
        check(self._L.pk_engine_add_keys(self._h, genome, a.ctypes.data, a.size))

    def add_sequence(self, genome: int, seq):
        a = _u8(seq)
        check(self._L.pk_engine_add_sequence(self._h, genome, a.ctypes.data, a.size))

    def add_sequence_device(self, genome: int, d_ptr: int, length: int):
        check(self._L.pk_engine_add_sequence_device(self._h, genome, d_ptr, length))

    def finalize(self):
        check(self._L.pk_engine_finalize(self._h))

    def seal_group(self, group: int):
        """Build the group table of local genomes [8*group, 8*group+8) now; with tune(group_only=1) their
        per-genome tables are freed (pk_engine_seal_group)."""
        check(self._L.pk_engine_seal_group(self._h, group))

    def table_stats(self, genome: int) -> dict:
        s = PkTableStats()
        check(self._L.pk_engine_table_stats(self._h, genome, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def group_stats(self, group: int) -> dict | None:
        """The group table of local genomes [8*group, 8*group+8) (pk_engine_group_stats); None when group tables
        are off or were not built."""
        s = PkTableStats()
        if self._L.pk_engine_group_stats(self._h, group, C.byref(s)) != 0:
            return None
        return {f: getattr(s, f) for f, _ in s._fields_}

    def sample_kmers(self, frac: float = 1.0) -> tuple[np.ndarray, np.ndarray]:
        """(keys uint64 [n], genome bit matrix rows as (local_group, mask8) tags uint32 [n]): the k-mers x of the local
        genomes with hash32(x) < frac * 2^32 (frac >= 1: all of them), each once per group of 8 genomes
        (pk_engine_sample_kmers)."""
        hmax = 0 if frac >= 1.0 else max(1, int(frac * 2.0 ** 32))
        total = sum((self.group_stats(u) or {"n_keys": 0})["n_keys"] for u in range((self.n_local + 7) // 8))
        cap = int(total * min(frac, 1.0) * 1.1) + 70_000
        for _ in range(3):
            keys = np.empty(cap, dtype=np.uint64)
            tags = np.empty(cap, dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self._L.pk_engine_sample_kmers(self._h, hmax, keys.ctypes.data, tags.ctypes.data, cap, C.byref(n))
            if rc == -4 and n.value > cap:              # PK_ENOMEM: the sample is larger than expected
                cap = int(n.value) + 1024
                continue
            check(rc)
            return keys[:n.value], tags[:n.value]
        raise RuntimeError("pk_engine_sample_kmers: sample size kept growing")

    # ---- hot path, host buffers --------------------------------------------
    def bin_len(self, nkmers: int) -> int:
        return int(self._L.pk_bin_len(C.byref(self.cfg), nkmers))

    def anchor_chrom(self, seq, bitmap1=True, low=True, hist=True, colsums=True, out=None) -> dict:
        """One chromosome -> rows / low-res rows / per-bin popcount histogram / column sums.

        `out` may carry preallocated (e.g. pinned) arrays under the keys
        'bitmap1' and 'low'. Returns a dict with nkmers and the requested arrays.
        """
        a = _u8(seq)
        nk = max(a.size - self.k + 1, 0)
        rb, step = self.row_bytes, self.lowres_step
        out = dict(out or {})
        res = {"nkmers": nk}
        if nk == 0:
            return res
        binlen = self.bin_len(nk)
        b1 = lo = hi = cs = None
        if bitmap1:
            b1 = out.get("bitmap1")
            if b1 is None:
                b1 = np.empty((nk, rb), dtype=np.uint8)
        if low:
            nlow = (nk + step - 1) // step
            lo = out.get("low")
            if lo is None:
                lo = np.empty((nlow, rb), dtype=np.uint8)
        if hist and binlen > 0:
            nbins = (nk + binlen - 1) // binlen
            hi = np.zeros((nbins, self.n_local + 1), dtype=np.uint64)
        if colsums:
            cs = np.zeros(self.n_local, dtype=np.uint64)
        nko = C.c_uint64(0)
        check(self._L.pk_anchor_chrom(self._h, a.ctypes.data, a.size,
                                      None if b1 is None else b1.ctypes.data,
                                      None if lo is None else lo.ctypes.data,
                                      None if hi is None else hi.ctypes.data,
                                      None if cs is None else cs.ctypes.data, C.byref(nko)))
        assert nko.value == nk
        res.update(bitmap1=None if b1 is None else b1[:nk], low=lo, bin_hist=hi, col_sums=cs,
                   binlen=binlen)
        return res

    def anchor_genome(self, seqs, bitmap1=True, low=True, hist=True, colsums=True, pinned=False,
                      out: dict | None = None) -> dict:
        """All chromosomes of one anchor in one batch (pk_anchor_genome).

        Returns {'chroms': [per-chromosome dicts as anchor_chrom], 'col_sums': [N_local]}.
        Chromosomes with fewer than min_bin_count k-mers get no histogram. `out` = a previous
        result for the same chromosome lengths whose (pinned) arrays are reused.
        """
        arrs = [_u8(s) for s in seqs]
        n = len(arrs)
        rb, step = self.row_bytes, self.lowres_step
        alloc = pinned_empty if pinned else (lambda shape, dtype=np.uint8: np.empty(shape, dtype=dtype))
        nks = [max(a.size - self.k + 1, 0) for a in arrs]
        bl = [self.bin_len(nk) if nk else 0 for nk in nks]
        if out is not None:
            assert [c["nkmers"] for c in out["chroms"]] == nks, "out= must come from the same chromosome lengths"
            b1 = [c["bitmap1"] for c in out["chroms"]]
            lo = [c["low"] for c in out["chroms"]]
            hi = [c["bin_hist"] for c in out["chroms"]]
            cs = out["col_sums"]
            if cs is not None:
                cs[:] = 0
        else:
            b1 = [alloc((nk, rb)) if (bitmap1 and nk) else None for nk in nks]
            lo = [alloc(((nk + step - 1) // step, rb)) if (low and nk) else None for nk in nks]
            hi = [np.zeros(((nk + b - 1) // b, self.n_local + 1), dtype=np.uint64) if (hist and nk and b) else None
                  for nk, b in zip(nks, bl)]
            cs = np.zeros(self.n_local, dtype=np.uint64) if colsums else None

        def ptrs(lst):
            return (C.c_void_p * n)(*[None if x is None else x.ctypes.data for x in lst])

        lens = (C.c_uint64 * n)(*[a.size for a in arrs])
        nko = (C.c_uint64 * n)()
        check(self._L.pk_anchor_genome(self._h, n, ptrs(arrs), lens, ptrs(b1), ptrs(lo), ptrs(hi),
                                       None if cs is None else cs.ctypes.data, nko))
        assert list(nko) == nks
        chroms = [{"nkmers": nk, "bitmap1": x, "low": y, "bin_hist": z, "binlen": b}
                  for nk, x, y, z, b in zip(nks, b1, lo, hi, bl)]
        return {"chroms": chroms, "col_sums": cs}

    def anchor_genome_bgzf(self, seqs, hist=True, colsums=True, out: dict | None = None) -> dict:
        """All chromosomes of one anchor in one batch, the two bitmaps delivered as FILE IMAGES
        (pk_anchor_genome_bgzf): 'gz'/'gzi' hold the bytes of bitmap.1.gz / bitmap.1.gzi and
        'gz_low'/'gzi_low' those of bitmap.<step>.gz/.gzi, compressed on the GPU. `out` = a previous
        result for the same chromosome lengths (its pinned buffers are reused)."""
        arrs = [_u8(s) for s in seqs]
        n = len(arrs)
        rb, step = self.row_bytes, self.lowres_step
        nks = [max(a.size - self.k + 1, 0) for a in arrs]
        bl = [self.bin_len(nk) if nk else 0 for nk in nks]
        nbytes = [sum(nks) * rb, sum((nk + step - 1) // step for nk in nks) * rb]
        if out is not None:
            assert [c["nkmers"] for c in out["chroms"]] == nks, "out= must come from the same chromosome lengths"
            bufs, hi, cs = out["_bufs"], [c["bin_hist"] for c in out["chroms"]], out["col_sums"]
            if cs is not None:
                cs[:] = 0
        else:
            bufs = {"gz": [pinned_empty(int(self._L.pk_bgzf_bound(b))) for b in nbytes],
                    "gzi": [pinned_empty(int(self._L.pk_bgzf_gzi_bound(b))) for b in nbytes]}
            hi = [np.zeros(((nk + b - 1) // b, self.n_local + 1), dtype=np.uint64) if (hist and nk and b) else None
                  for nk, b in zip(nks, bl)]
            cs = np.zeros(self.n_local, dtype=np.uint64) if colsums else None

        def ptrs(lst, m):
            return (C.c_void_p * m)(*[None if x is None else x.ctypes.data for x in lst])

        lens = (C.c_uint64 * n)(*[a.size for a in arrs])
        nko = (C.c_uint64 * n)()
        caps_gz = (C.c_uint64 * 2)(*[b.size for b in bufs["gz"]])
        caps_gzi = (C.c_uint64 * 2)(*[b.size for b in bufs["gzi"]])
        sizes = (C.c_uint64 * 4)()
        check(self._L.pk_anchor_genome_bgzf(self._h, n, ptrs(arrs, n), lens, ptrs(bufs["gz"], 2), caps_gz,
                                            ptrs(bufs["gzi"], 2), caps_gzi, sizes, ptrs(hi, n),
                                            None if cs is None else cs.ctypes.data, nko))
        assert list(nko) == nks
        chroms = [{"nkmers": nk, "bin_hist": z, "binlen": b} for nk, z, b in zip(nks, hi, bl)]
        return {"chroms": chroms, "col_sums": cs, "gz": bufs["gz"][0][:sizes[0]], "gzi": bufs["gzi"][0][:sizes[1]],
                "gz_low": bufs["gz"][1][:sizes[2]], "gzi_low": bufs["gzi"][1][:sizes[3]], "_bufs": bufs}

    def anchor_layout(self, lens) -> tuple[list[int], int]:
        """The concatenated row numbering pk_anchor_genome_plane uses: ([first row of every chromosome], rows a
        plane needs) (pk_anchor_layout)."""
        n = len(lens)
        arr = (C.c_uint64 * n)(*[int(l) for l in lens])
        off = (C.c_uint64 * n)()
        total = int(self._L.pk_anchor_layout(n, arr, off))
        return [int(x) for x in off], total

    def anchor_genome_plane(self, seqs, d_plane: int, plane_rows: int, row_stride: int | None = None) -> list[int]:
        """H2D + pack + probe of all chromosomes of one anchor into a caller-owned device plane
        [plane_rows][row_stride] (pk_anchor_genome_plane: the rank-local half of the genome-sharded path;
        row_stride defaults to this shard's row bytes). Returns nkmers per chromosome; complete on return."""
        arrs = [_u8(s) for s in seqs]
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lens = (C.c_uint64 * n)(*[a.size for a in arrs])
        nko = (C.c_uint64 * n)()
        check(self._L.pk_anchor_genome_plane(self._h, n, ptrs, lens, d_plane, plane_rows, row_stride or self.row_bytes, nko))
        return [int(x) for x in nko]

    def anchor_paircount_bins(self, nkmers: list[int], bin_positions: int) -> list[np.ndarray]:
        """Per chromosome the [bins, N_local] counts of low-res rows with each genome's bit set, per bin of
        `bin_positions` positions, for the anchor the last anchor_genome / anchor_genome_bgzf call processed — reduced
        on the device from its low-res rows (pk_anchor_paircount_bins)."""
        step = self.lowres_step
        rpb = (bin_positions + step - 1) // step
        nb = [(((nk + step - 1) // step) + rpb - 1) // rpb for nk in nkmers]
        out = np.zeros((max(sum(nb), 1), self.n_local), dtype=np.uint32)
        arr = (C.c_uint64 * len(nkmers))(*[int(x) for x in nkmers])
        check(self._L.pk_anchor_paircount_bins(self._h, len(nkmers), arr, bin_positions, out.ctypes.data))
        res, o = [], 0
        for n in nb:
            res.append(out[o:o + n])
            o += n
        return res

    def paircount_bins_device(self, d_rows_low: int, row_stride: int, n_cols: int, n_rows: int, rows_per_bin: int,
                              d_counts: int, stream: int | None = None):
        check(self._L.pk_paircount_bins_device(self._h, d_rows_low, row_stride, n_cols, n_rows, rows_per_bin, d_counts,
                                               self._st(stream)))

    def bgzf_bound(self, nbytes: int) -> tuple[int, int]:
        return int(self._L.pk_bgzf_bound(nbytes)), int(self._L.pk_bgzf_gzi_bound(nbytes))

    def bgzf_compress_device(self, d_in: int, nbytes: int, match_dist: int, d_gz: int, d_gzi: int, d_totals: int,
                             stream: int | None = None):
        """Device bytes -> BGZF .gz image + .gzi image + totals (2 x uint64), all on the device."""
        check(self._L.pk_bgzf_compress_device(self._h, d_in or None, nbytes, match_dist, d_gz, d_gzi, d_totals,
                                              self._st(stream)))

    def get_counters_for_read(self, dbi: int, read) -> np.ndarray | None:
        """uint32 counters of bitvec database `dbi` for every k-mer of `read`
        (CKMCFile::GetCountersForRead); None when len(read) < k."""
        a = _u8(read)
        if a.size < self.k:
            return None
        outv = np.empty(a.size - self.k + 1, dtype=np.uint32)
        n = C.c_uint64(0)
        check(self._L.pk_get_counters_for_read(self._h, dbi, a.ctypes.data, a.size, outv.ctypes.data,
                                               C.byref(n)))
        return outv[:n.value]

    def stats(self) -> dict:
        s = PkStats()
        check(self._L.pk_engine_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def tune(self, **knobs):
        """Tuning knobs of the partitioned probe (pk_engine_tune); results never depend on them."""
        for name, value in knobs.items():
            check(self._L.pk_engine_tune(self._h, name.encode(), int(value)))

    # ---- hot path, device pointers (ints; e.g. torch.Tensor.data_ptr()) ------
    # `stream`: a cudaStream_t handle as an int (torch.cuda.Stream.cuda_stream). None = the engine's own stream.
    # 0 is the handle of CUDA's legacy default stream — torch's default stream — and is passed on as such
    # (cudaStreamLegacy), never silently replaced by the engine's stream: work a caller orders on torch's default
    # stream stays ordered with the library's kernels.
    @staticmethod
    def _st(stream):
        if stream is None:
            return None
        return 1 if stream == 0 else stream          # cudaStreamLegacy == (cudaStream_t)0x1

    def packed_words(self, length: int) -> int:
        return int(self._L.pk_packed_words(length))

    def pack_device(self, d_ascii: int, length: int, d_words: int, d_mask: int, stream: int | None = None):
        check(self._L.pk_pack_device(self._h, d_ascii, length, d_words, d_mask, self._st(stream)))

    def probe_device(self, d_words: int, d_mask: int, p0: int, n: int, d_rows: int, row_stride: int,
                     col_offset: int = 0, stream: int | None = None):
        check(self._L.pk_probe_device(self._h, d_words, d_mask, p0, n, d_rows, row_stride, col_offset,
                                      self._st(stream)))

    def reduce_device(self, d_rows: int, row_stride: int, n_cols: int, p_first: int, n: int, binlen: int,
                      d_hist: int = 0, d_colsums: int = 0, d_low: int = 0, stream: int | None = None):
        check(self._L.pk_reduce_device(self._h, d_rows, row_stride, n_cols, p_first, n, binlen,
                                       d_hist or None, d_colsums or None, d_low or None,
                                       self.lowres_step, self._st(stream)))

    # ---- peer memory (genome-sharded exchange without NCCL) ---------------------
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self._L.pk_device_alloc(self._h, C.byref(p), nbytes))
        return p.value

    def device_free(self, d_ptr: int):
        check(self._L.pk_device_free(self._h, d_ptr))

    def ipc_export(self, d_ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        check(self._L.pk_ipc_export(self._h, d_ptr, buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        check(self._L.pk_ipc_open(self._h, buf, C.byref(p)))
        return p.value

    def ipc_close(self, d_ptr: int):
        check(self._L.pk_ipc_close(self._h, d_ptr))

    def gather_interleave_device(self, plane_ptrs: list[int], n: int, w: int, d_rows: int, row_stride: int,
                                 stream: int | None = None):
        arr = (C.c_void_p * len(plane_ptrs))(*plane_ptrs)
        check(self._L.pk_gather_interleave_device(self._h, arr, len(plane_ptrs), n, w, d_rows, row_stride,
                                                  self._st(stream)))

    def gather_slice_device(self, plane_ptrs: list[int], plane_rows: int, w: int, segments, d_rows: int,
                            row_stride: int, row_bytes: int, stream: int | None = None):
        """This rank's slice of the rows out of every rank's plane (pk_gather_slice_device).
        segments: [(src_row, n_rows, dst_row)]."""
        arr = (C.c_void_p * len(plane_ptrs))(*plane_ptrs)
        segs = (_lib.PkSegment * len(segments))(*[_lib.PkSegment(int(a), int(b), int(c)) for a, b, c in segments])
        check(self._L.pk_gather_slice_device(self._h, arr, len(plane_ptrs), plane_rows, w, segs, len(segments),
                                             d_rows, row_stride, row_bytes, self._st(stream)))

    def interleave_device(self, d_planes: int, n_ranks: int, n: int, w: int, d_rows: int, row_stride: int,
                          stream: int | None = None):
        check(self._L.pk_interleave_device(self._h, d_planes, n_ranks, n, w, d_rows, row_stride,
                                           self._st(stream)))
