"""Genome-sharded anchoring across the GPUs of one node (one process per GPU).

The bitmap shards by genome = by column: rank g of a GENOME GROUP of Rg ranks owns the tables of a contiguous,
byte-aligned genome range and probes EVERY anchor position against them into its *plane* (its bytes of every
row). One exchange step then assembles full rows — position-split: rank g builds only ITS slice of the output
stream, reading that slice of all Rg planes in place over NVLink (pk_gather_slice_device), and goes on to reduce,
compress (BGZF on the GPU) and store exactly those rows. Slices are cut on BGZF-member boundaries (multiples of
0xff00 payload bytes), so the members the ranks produce concatenate into the very file a single GPU writes.
Small per-anchor reductions (bin histograms, column sums, the low-res rows) are summed with an all-reduce.

When every rank can hold more than 1/world of the tables, the world is a GRID of Rp genome groups (world =
Rg x Rp): the groups hold replicas of the tables and work on different anchors at the same time — no exchange
between groups at all.

torch.distributed is the plumbing (NCCL on GPUs; the host logic below runs on gloo/CPU tensors in the tests);
all compute is in libpkanchor.so. cpp/anchor.cpp has no counterpart for any of this (one process holds every
genome, `#pragma omp parallel for` over anchors, cpp/anchor.cpp:217): what must hold is that the directory
written here equals the directory a single engine writes.
"""
from __future__ import annotations

import math
import os
import struct
from pathlib import Path

import numpy as np

BGZF_PAYLOAD = 0xFF00
BGZF_EOF_LEN = 28


# ------------------------------------------------------------------ pure host logic (tested on the CPU)
def shard_bounds(n_genomes: int, world: int) -> list[tuple[int, int]]:
    """Contiguous genome ranges on 8-genome (byte) boundaries, as equal as possible.
    Every rank gets the same byte width except possibly trailing ranks, which may be
    narrower or empty when N is small."""
    nbytes = (n_genomes + 7) // 8
    per = (nbytes + world - 1) // world            # bytes per rank
    out = []
    for r in range(world):
        b0 = min(r * per, nbytes)
        b1 = min((r + 1) * per, nbytes)
        out.append((min(8 * b0, n_genomes), min(8 * b1, n_genomes)))
    return out


def plane_width(n_genomes: int, world: int) -> int:
    """Bytes per row every rank contributes (narrow last shards are padded)."""
    return ((n_genomes + 7) // 8 + world - 1) // world


def slice_bounds(total_rows: int, row_bytes: int, world: int) -> list[int]:
    """world+1 row boundaries of the position slices: every inner boundary falls on a BGZF member boundary of the
    row stream (a multiple of 0xff00 bytes that is also a whole number of rows), as equal as possible."""
    unit = BGZF_PAYLOAD * row_bytes // math.gcd(BGZF_PAYLOAD, row_bytes) // row_bytes     # rows per lcm(0xff00, row_bytes) bytes
    units = (total_rows + unit - 1) // unit
    b = [min(total_rows, (units * r // world) * unit) for r in range(world + 1)]
    b[-1] = total_rows
    return b


def stream_segments(cat_off, nks, s0: int, s1: int) -> list[tuple[int, int, int]]:
    """Rows [s0, s1) of the bitmap stream (chromosomes back to back) as runs of plane rows:
    [(src_row in the cat numbering, n_rows, dst_row relative to s0)]."""
    segs, so = [], 0
    for off, nk in zip(cat_off, nks):
        a, b = max(s0, so), min(s1, so + nk)
        if b > a:
            segs.append((int(off) + a - so, b - a, a - s0))
        so += nk
    return segs


def slice_pieces(nks, s0: int, s1: int) -> list[tuple[int, int, int, int]]:
    """The chromosome pieces inside stream rows [s0, s1): [(chromosome, first position in it, n, row in the slice)]."""
    out, so = [], 0
    for c, nk in enumerate(nks):
        a, b = max(s0, so), min(s1, so + nk)
        if b > a:
            out.append((c, a - so, b - a, a - s0))
        so += nk
    return out


def merge_bgzf_parts(sizes: list[tuple[int, int]], uoffs: list[int]) -> tuple[list[int], int]:
    """Where every rank's members go in the final .gz: part r is a complete BGZF image (members + the 28-byte EOF
    member) of `sizes[r][0]` bytes for the stream bytes starting at uoffs[r]; every part but the last is written
    without its EOF member. Returns (file offset per part, total file size)."""
    offs, o = [], 0
    last = len(sizes) - 1
    for r, (gz, _) in enumerate(sizes):
        offs.append(o)
        o += gz if r == last else gz - BGZF_EOF_LEN
    return offs, o


def merge_gzi(parts: list[bytes], coffs: list[int], uoffs: list[int], nbytes: list[int]) -> bytes:
    """The .gzi of the concatenated file from the per-part .gzi images (uint64 n, then n x (compressed, uncompressed)
    offsets of every member AFTER the part's first). A non-empty part that is not the first of the file also
    contributes the start of its first member. Layout read by Genome.load_bgz_blocks (index.py:793-799)."""
    ent = []
    first = True
    for raw, co, uo, nb in zip(parts, coffs, uoffs, nbytes):
        if nb == 0:
            continue
        (n,) = struct.unpack_from("<Q", raw, 0)
        e = np.frombuffer(raw, dtype="<u8", count=2 * n, offset=8).reshape(n, 2).astype(np.uint64)
        if not first:
            ent.append(np.array([[co, uo]], dtype=np.uint64))
        ent.append(e + np.array([co, uo], dtype=np.uint64))
        first = False
    allent = np.concatenate(ent) if ent else np.zeros((0, 2), dtype=np.uint64)
    return struct.pack("<Q", len(allent)) + allent.astype("<u8").tobytes()


def interleave_planes_host(planes: np.ndarray, n_genomes: int) -> np.ndarray:
    """Host restatement of the exchange for tests: [R, n, w] -> [n, ceil(N/8)]."""
    r, n, w = planes.shape
    rows = planes.transpose(1, 0, 2).reshape(n, r * w)
    return np.ascontiguousarray(rows[:, : (n_genomes + 7) // 8])


def gather_planes(local, world: int, group=None):
    """all-gather of per-rank row planes: local [n, w] uint8 -> [world, n, w] (the NCCL form of the exchange)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def grid_shape(world: int, genome_ranks: int | None) -> tuple[int, int]:
    """(Rg, Rp): ranks per genome group, number of groups."""
    rg = genome_ranks or world
    if rg < 1 or world % rg:
        raise ValueError(f"genome_ranks={rg} does not divide the world size {world}")
    return rg, world // rg


# ------------------------------------------------------------------ the rank-local engine + the exchange
class ShardedAnchorer:
    """Rank-local engine + the exchange step.

    Streams: every device-level call of this class (probe, gather, reduce, BGZF) and every collective is issued on
    ONE torch stream, `self.stream`, made current for the duration of the call — the collectives of
    torch.distributed are ordered on torch's current stream, so kernels and barriers are ordered with each other
    by construction (the library itself never synchronises device-level calls)."""

    def __init__(self, k: int, n_genomes: int, rank: int, world: int, device: int, genome_ranks: int | None = None,
                 **engine_kw):
        import torch
        import torch.distributed as dist
        from .engine import Engine
        self.rank, self.world, self.n_genomes = rank, world, n_genomes
        self.rg, self.rp = grid_shape(world, genome_ranks)
        self.gi, self.pi = rank % self.rg, rank // self.rg           # place in the genome group, group index
        self.bounds = shard_bounds(n_genomes, self.rg)
        self.begin, self.end = self.bounds[self.gi]
        if self.end <= self.begin:
            raise ValueError(f"rank {rank} owns no genomes: use at most {(n_genomes + 7) // 8} ranks per genome group")
        self.w = plane_width(n_genomes, self.rg)
        self.row_bytes = (n_genomes + 7) // 8
        self.group = None
        if self.rp > 1:                                                # new_group is collective over the WORLD, in one order
            for p in range(self.rp):
                g = dist.new_group(ranks=[p * self.rg + j for j in range(self.rg)])
                if p == self.pi:
                    self.group = g
        self.engine = Engine(k, n_genomes, self.begin, self.end, device=device, **engine_kw)
        self.k, self.step = k, self.engine.lowres_step
        self.dev = torch.device(f"cuda:{device}")
        self.stream = torch.cuda.Stream(device=self.dev)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._gather_stream = None   # side stream of probe_exchange(overlap=True)
        self._gather_done = None     # event: the overlapped gather of the previous step
        self._planes = None          # [2] ping-pong: {"rows", "own", "peers"}
        self._host_rows = None       # page-locked staging of anchor_genome(rows_to_host=True), grow-only
        self._host_gz = None         # page-locked staging of the BGZF image of the slice
        self._i = 0
        self.umap_bin_size = 100000  # positions per pair-count bin (Genome.write_umaps, index.py:1107)
        self.last = {}               # timings of the last anchor_genome call (ms)

    def owns(self, genome: int) -> bool:
        return self.begin <= genome < self.end

    # ---- peer memory ----
    def _ensure_planes(self, plane_rows: int):
        """Two planes of >= plane_rows rows, IPC-mapped into every rank of the genome group. Collective (the growth
        decision depends only on plane_rows, which is the same on every rank of the group)."""
        import torch.distributed as dist
        if self._planes and self._planes[0]["rows"] >= plane_rows:
            return
        self.close_p2p()
        eng = self.engine
        rows = (plane_rows + 4095) // 4096 * 4096
        self._planes = []
        for _ in range(2):
            own = eng.device_alloc(rows * self.w)
            handles = [None] * self.rg
            if self.rg > 1:
                dist.all_gather_object(handles, eng.ipc_export(own), group=self.group)
            peers = [own if j == self.gi else eng.ipc_open(handles[j]) for j in range(self.rg)]
            self._planes.append({"rows": rows, "own": own, "peers": peers})

    def close_p2p(self):
        import torch
        import torch.distributed as dist
        if not self._planes:
            return
        torch.cuda.synchronize(self.dev)
        if self.rg > 1:
            dist.barrier(group=self.group)
        for pl in self._planes:
            for j, p in enumerate(pl["peers"]):
                if j != self.gi:
                    self.engine.ipc_close(p)
        if self.rg > 1:
            dist.barrier(group=self.group)
        for pl in self._planes:
            self.engine.device_free(pl["own"])
        self._planes = None

    def barrier(self):
        """Stream-ordered barrier over the genome group (a 1-element all-reduce on self.stream)."""
        import torch.distributed as dist
        if self.rg > 1:
            dist.all_reduce(self._flag, group=self.group)

    # ---- device-level step (bench.py and the parity tests drive this) ----
    def finish_exchange(self):
        """After probe_exchange(..., overlap=True): self.stream waits for the gather still running on the side stream."""
        if self._gather_stream is not None:
            self.stream.wait_stream(self._gather_stream)

    def probe_exchange(self, d_words: int, d_mask: int, npos: int, segments, d_rows, timing=None, overlap: bool = False):
        """Probe all `npos` positions against the local shard into this rank's plane, barrier, then assemble THIS
        rank's slice (`segments`, from stream_segments) of the full rows into d_rows [(rows of the slice), row_bytes].
        Call under `with torch.cuda.stream(self.stream)`. One barrier per step: planes alternate, and a rank
        reaches the barrier of step i only after its own gather of step i-1, so when the barrier of step i
        releases, nobody still reads the plane that step i+1 overwrites. `timing` = 3 CUDA events recorded
        before the probe, after the barrier and after the gather.
        overlap=True: the gather runs on a side stream, under the probe of the NEXT step (which writes the other plane;
        the gather kernel uses no shared memory and few registers, it fits beside the probe kernels). The invariant
        above is kept by making this stream wait for the rank's own previous gather before it enters the barrier.
        d_rows is complete after finish_exchange() (or the next step's barrier)."""
        import torch
        eng, st = self.engine, self.stream.cuda_stream
        pl = self._planes[self._i & 1]
        self._i += 1
        if timing:
            timing[0].record(self.stream)
        eng.probe_device(d_words, d_mask, 0, npos, pl["own"], self.w, 0, st)
        if self._gather_done is not None:
            self.stream.wait_event(self._gather_done)       # no-op unless an overlapped gather is still in flight
            self._gather_done = None
        self.barrier()
        if timing:
            timing[1].record(self.stream)
        gs = self.stream
        if overlap:
            if self._gather_stream is None:
                self._gather_stream = torch.cuda.Stream(device=self.dev)
            gs = self._gather_stream
            gs.wait_stream(self.stream)                     # the barrier (and through it every peer's probe) comes first
        if segments:
            eng.gather_slice_device(pl["peers"], pl["rows"], self.w, segments, d_rows.data_ptr(), d_rows.shape[1],
                                    self.row_bytes, gs.cuda_stream)
        if overlap:
            self._gather_done = torch.cuda.Event()
            self._gather_done.record(gs)
        if timing:
            timing[2].record(gs)
        return d_rows

    def probe_allgather(self, d_words: int, d_mask: int, npos: int, d_local, d_planes, d_rows):
        """The NCCL form of the exchange, kept for comparison (bench.py --exchange nccl): local probe ->
        all_gather_into_tensor -> interleave kernel; every rank ends up with every row. d_local [npos, w],
        d_planes [Rg, npos, w], d_rows [npos, Rg*w] are uint8 tensors on this rank's device. Call under
        `with torch.cuda.stream(self.stream)`."""
        import torch.distributed as dist
        eng, st = self.engine, self.stream.cuda_stream
        eng.probe_device(d_words, d_mask, 0, npos, d_local.data_ptr(), self.w, 0, st)
        if self.rg == 1:
            return d_local
        dist.all_gather_into_tensor(d_planes.view(-1), d_local.view(-1), group=self.group)
        eng.interleave_device(d_planes.data_ptr(), self.rg, npos, self.w, d_rows.data_ptr(), self.rg * self.w, st)
        return d_rows

    def _load_anchor_shared(self, arrs, cat_off, ltot: int):
        """The packed anchor (2-bit words + invalid-base mask, concatenated numbering of pk_anchor_layout) on every rank
        of the genome group, each rank having copied and packed only its 1/Rg share. Call on self.stream."""
        import torch
        import torch.distributed as dist
        eng, rg, gi = self.engine, self.rg, self.gi
        words32 = (ltot + 31) // 32
        cw = ((words32 + rg - 1) // rg + 1) & ~1                      # packed words per rank, even (the mask is read as uint64)
        a0, a1 = gi * cw * 32, min(ltot, (gi + 1) * cw * 32)          # this rank's bases
        share = torch.full((cw * 32 + 64,), ord("N"), dtype=torch.uint8, device=self.dev)
        for off, a in zip(cat_off, arrs):
            lo, hi = max(a0, off), min(a1, off + a.size)
            if hi > lo:
                share[lo - a0:hi - a0].copy_(torch.from_numpy(a[lo - off:hi - off]), non_blocking=True)
        nwt = eng.packed_words(cw * 32)
        tw = torch.empty(nwt, dtype=torch.int64, device=self.dev)
        tm = torch.empty(nwt, dtype=torch.int32, device=self.dev)
        eng.pack_device(share.data_ptr(), cw * 32, tw.data_ptr(), tm.data_ptr(), self.stream.cuda_stream)
        total = max(rg * cw, eng.packed_words(ltot)) + 8
        d_words = torch.zeros(total, dtype=torch.int64, device=self.dev)
        d_mask = torch.full((total,), -1, dtype=torch.int32, device=self.dev)      # beyond the anchor: invalid bases
        dist.all_gather_into_tensor(d_words[: rg * cw], tw[:cw], group=self.group)
        dist.all_gather_into_tensor(d_mask[: rg * cw], tm[:cw], group=self.group)
        return d_words, d_mask

    # ---- the product call: one anchor genome -> this rank's share of its results ----
    def anchor_genome(self, seqs, bgzf: bool = True, rows_to_host: bool = False) -> dict:
        """All chromosomes of one anchor. Collective over the genome group. Every rank returns
          nkmers [C], slice (s0, s1) in stream rows, rows (device tensor [s1-s0, row_bytes], this rank's slice),
          hist [per chromosome: bins x (N+1)], col_sums [N], low (host, all low-res rows) — the three reduced over
          the group — and, with bgzf, gz / gzi: the BGZF image of the slice compressed on this rank's GPU; with
          rows_to_host, rows_host: the slice's rows in (reused, page-locked) host memory."""
        import torch
        import torch.distributed as dist
        from .engine import _u8
        eng, k, step, rb, N = self.engine, self.k, self.step, self.row_bytes, self.n_genomes
        arrs = [_u8(s) for s in seqs]
        lens = [a.size for a in arrs]
        cat_off, plane_rows = eng.anchor_layout(lens)
        nks = [max(l - k + 1, 0) for l in lens]
        total = sum(nks)
        self._ensure_planes(max(plane_rows, 1))
        sb = slice_bounds(total, rb, self.rg)
        s0, s1 = sb[self.gi], sb[self.gi + 1]
        pl = self._planes[self._i & 1]
        self._i += 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        with torch.cuda.stream(self.stream):
            ev[0].record(self.stream)
            if self.rg == 1:
                # H2D + pack + probe, pipelined inside the library on its own streams; complete on return
                eng.anchor_genome_plane(arrs, pl["own"], pl["rows"], self.w)
            else:
                # every rank needs the whole anchor, but not over its own PCIe link: rank g copies and packs 1/Rg of
                # it, an all-gather over NVLink hands everyone the packed sequence (0.375 byte per base)
                d_words, d_mask = self._load_anchor_shared(arrs, cat_off, plane_rows)
                eng.probe_device(d_words.data_ptr(), d_mask.data_ptr(), 0, max(plane_rows - k + 1, 0), pl["own"], self.w, 0,
                                 self.stream.cuda_stream)
            ev[1].record(self.stream)
            self.barrier()
            rows = torch.empty((max(s1 - s0, 1), rb), dtype=torch.uint8, device=self.dev)
            segs = stream_segments(cat_off, nks, s0, s1)
            if segs:
                eng.gather_slice_device(pl["peers"], pl["rows"], self.w, segs, rows.data_ptr(), rb, rb, self.stream.cuda_stream)
            ev[2].record(self.stream)
            if rows_to_host:
                # the slice's rows go home on the copy stream while this stream reduces them
                nb = (s1 - s0) * rb
                if self._host_rows is None or self._host_rows.numel() < nb:
                    self._host_rows = torch.empty(max(nb, 1), dtype=torch.uint8, pin_memory=True)
                self.copy_stream.wait_stream(self.stream)
                with torch.cuda.stream(self.copy_stream):
                    self._host_rows[:nb].copy_(rows.view(-1)[:nb], non_blocking=True)
                rows.record_stream(self.copy_stream)
            # reductions over this rank's rows, then summed over the group
            binlen = [eng.bin_len(nk) if nk else 0 for nk in nks]
            nbins = [(nk + b - 1) // b if b else 0 for nk, b in zip(nks, binlen)]
            hoff = np.concatenate(([0], np.cumsum([nb * (N + 1) for nb in nbins]))).astype(np.int64)
            nlow = [(nk + step - 1) // step for nk in nks]
            loff = np.concatenate(([0], np.cumsum(nlow))).astype(np.int64)
            red = torch.zeros(int(hoff[-1]) + N, dtype=torch.int64, device=self.dev)       # histograms, then column sums
            low = torch.zeros((max(int(loff[-1]), 1), rb), dtype=torch.uint8, device=self.dev)
            for c, p_first, n, r0 in slice_pieces(nks, s0, s1):
                l0 = (p_first + step - 1) // step
                eng.reduce_device(rows.data_ptr() + r0 * rb, rb, N, p_first, n, binlen[c],
                                  red.data_ptr() + 8 * int(hoff[c]) if nbins[c] else 0,
                                  red.data_ptr() + 8 * int(hoff[-1]), low.data_ptr() + (int(loff[c]) + l0) * rb,
                                  self.stream.cuda_stream)
            if self.rg > 1:
                dist.all_reduce(red, group=self.group)
                dist.all_reduce(low, group=self.group)           # disjoint supports: the sum is the union
            # pair-count bins (the UMAP input) from the complete low-res rows, on the device
            rpb = (self.umap_bin_size + step - 1) // step
            pcn = [(nl + rpb - 1) // rpb for nl in nlow]
            pc = torch.zeros((max(sum(pcn), 1), N), dtype=torch.int32, device=self.dev)
            o = 0
            for c, nl in enumerate(nlow):
                if nl:
                    eng.paircount_bins_device(low.data_ptr() + int(loff[c]) * rb, rb, N, nl, rpb, pc.data_ptr() + 4 * o * N,
                                              self.stream.cuda_stream)
                o += pcn[c]
            ev[3].record(self.stream)
            out = {"nkmers": nks, "slice": (s0, s1), "rows": rows[: s1 - s0], "binlen": binlen}
            if bgzf:
                nb = (s1 - s0) * rb
                cap_gz, cap_gzi = eng.bgzf_bound(nb)
                gz = torch.empty(cap_gz, dtype=torch.uint8, device=self.dev)
                gzi = torch.empty(cap_gzi // 8 + 1, dtype=torch.int64, device=self.dev)
                tot = torch.zeros(2, dtype=torch.int64, device=self.dev)
                eng.bgzf_compress_device(rows.data_ptr(), nb, rb, gz.data_ptr(), gzi.data_ptr(), tot.data_ptr(),
                                         self.stream.cuda_stream)
                t = tot.cpu()                                     # synchronises self.stream
                ngz, ngzi = int(t[0]), int(t[1])
                if self._host_gz is None or self._host_gz.numel() < ngz + ngzi:
                    self._host_gz = torch.empty(cap_gz + cap_gzi + 16, dtype=torch.uint8, pin_memory=True)
                self._host_gz[:ngz].copy_(gz[:ngz], non_blocking=True)
                self._host_gz[ngz:ngz + ngzi].copy_(gzi.view(torch.uint8)[:ngzi], non_blocking=True)
                self.stream.synchronize()
                out["gz"] = self._host_gz[:ngz].numpy()           # page-locked staging, valid until the next call
                out["gzi"] = self._host_gz[ngz:ngz + ngzi].numpy()
            if rows_to_host:
                self.stream.wait_stream(self.copy_stream)
                out["rows_host"] = self._host_rows[:(s1 - s0) * rb].numpy().reshape(s1 - s0, rb)
            ev[4].record(self.stream)
            red_h = red.cpu().numpy().astype(np.uint64)
            out["hist"] = [red_h[int(hoff[c]):int(hoff[c + 1])].reshape(nbins[c], N + 1) if nbins[c] else None
                           for c in range(len(nks))]
            out["col_sums"] = red_h[int(hoff[-1]):]
            out["low"] = low[: int(loff[-1])].cpu().numpy()
            pc_h = pc.cpu().numpy().astype(np.uint32)
            out["pc_counts"] = [pc_h[sum(pcn[:c]):sum(pcn[:c + 1])] for c in range(len(nks))]
        torch.cuda.synchronize(self.dev)
        self.last = {"probe_ms": ev[0].elapsed_time(ev[1]), "exchange_ms": ev[1].elapsed_time(ev[2]),
                     "reduce_ms": ev[2].elapsed_time(ev[3]), "bgzf_d2h_ms": ev[3].elapsed_time(ev[4])}
        return out


def assemble_bitmap(gi: int, rg: int, group, gz_path, gzi_path, gz: np.ndarray, gzi: np.ndarray, uoff: int, nbytes: int):
    """All ranks of a genome group write ONE BGZF file: rank gi holds the complete BGZF image `gz` (+ `.gzi` image)
    of the `nbytes` stream bytes that start at `uoff`. The sizes are exchanged, rank 0 sizes the file, every rank
    writes its members at its offset (all but the last non-empty part without the EOF member) and rank 0 writes
    the merged index. Collective over the group (any backend: only object collectives and barriers)."""
    import torch.distributed as dist
    mine = (int(gz.size), int(gzi.size), int(uoff), int(nbytes))
    allm = [None] * rg
    if rg > 1:
        dist.all_gather_object(allm, mine, group=group)          # also orders the caller's mkdir before the writes
    else:
        allm = [mine]
    live = [r for r in range(rg) if allm[r][3] > 0] or [0]       # parts that hold rows, in stream order
    offs, total = merge_bgzf_parts([(allm[r][0], allm[r][1]) for r in live], [allm[r][2] for r in live])
    if gi == 0:
        with open(gz_path, "wb") as fh:
            fh.truncate(total)
    if rg > 1:
        dist.barrier(group=group)
    if gi in live:
        j = live.index(gi)
        data = gz if j == len(live) - 1 else gz[: gz.size - BGZF_EOF_LEN]
        fd = os.open(gz_path, os.O_WRONLY)
        try:
            os.pwrite(fd, data.tobytes(), offs[j])
        finally:
            os.close(fd)
    gzis = [None] * rg
    if rg > 1:
        dist.all_gather_object(gzis, gzi.tobytes(), group=group)
    else:
        gzis = [gzi.tobytes()]
    if gi == 0:
        Path(gzi_path).write_bytes(merge_gzi([gzis[r] for r in live], offs, [allm[r][2] for r in live],
                                             [allm[r][3] for r in live]))
    if rg > 1:
        dist.barrier(group=group)                                # every rank's members are in the file


# ------------------------------------------------------------------ one anchor -> its directory, written by all ranks
def anchor_fasta_sharded(sh: ShardedAnchorer, name: str, fasta, outdir, genome_names: list[str] | None = None,
                         strip_cr: bool = False, umap_bin_size: int = 100000) -> dict:
    """anchor.anchor_fasta for a genome group: every rank anchors against its shard, assembles, reduces and
    compresses its slice; the ranks write their BGZF members into bitmap.1.gz side by side (pwrite at the offsets
    the sizes imply), rank 0 of the group writes the index, the low-res bitmap and the text files. The directory
    is built under `<outdir>.tmp` and renamed when complete. Collective over the genome group."""
    import torch.distributed as dist
    from . import anchor as anchor_mod
    from . import layout
    eng = sh.engine
    outdir = Path(outdir)
    tmp = outdir.with_name(outdir.name + ".tmp")
    step, rb = sh.step, sh.row_bytes
    recs = anchor_mod.parse_fasta(fasta, strip_cr=strip_cr)
    for cname, seq in recs:
        nk = seq.size - sh.k + 1
        if nk < 1 or eng.bin_len(nk) == 0:
            raise ValueError(f"{fasta}: chromosome {cname!r} has {max(nk, 0)} k-mers; the reference "
                             "needs at least min_bin_count (cpp/anchor.cpp:116-120)")
    sh.umap_bin_size = umap_bin_size
    res = sh.anchor_genome([s for _, s in recs], bgzf=True)
    lead = sh.gi == 0
    if lead:
        if tmp.exists():
            import shutil
            shutil.rmtree(tmp)
        tmp.mkdir(parents=True)
    s0, s1 = res["slice"]
    assemble_bitmap(sh.gi, sh.rg, sh.group, tmp / "bitmap.1.gz", tmp / "bitmap.1.gzi", res["gz"], res["gzi"], s0 * rb,
                    (s1 - s0) * rb)
    out = {"positions": sum(res["nkmers"]), "chroms": len(recs), "col_sums": res["col_sums"]}
    if lead:
        # the low-res bitmap is small (1/step of the rows): the group's leader compresses it on its GPU
        import torch
        low = res["low"]
        with torch.cuda.stream(sh.stream):
            d_low = torch.from_numpy(np.ascontiguousarray(low)).to(sh.dev)
            cap_gz, cap_gzi = eng.bgzf_bound(low.size)
            gz = torch.empty(cap_gz, dtype=torch.uint8, device=sh.dev)
            gzi = torch.empty(cap_gzi // 8 + 1, dtype=torch.int64, device=sh.dev)
            tot = torch.zeros(2, dtype=torch.int64, device=sh.dev)
            eng.bgzf_compress_device(d_low.data_ptr(), low.size, rb, gz.data_ptr(), gzi.data_ptr(), tot.data_ptr(),
                                     sh.stream.cuda_stream)
            t = tot.cpu()
            (tmp / f"bitmap.{step}.gz").write_bytes(gz[: int(t[0])].cpu().numpy().tobytes())
            (tmp / f"bitmap.{step}.gzi").write_bytes(gzi.view(torch.uint8)[: int(t[1])].cpu().numpy().tobytes())
        anchor_mod.write_text_files(tmp, name, [c for c, _ in recs], res["nkmers"], res["binlen"], res["hist"],
                                    res["col_sums"], res["pc_counts"], sh.n_genomes, step, genome_names, umap_bin_size)
    if sh.rg > 1:
        dist.barrier(group=sh.group)                             # every rank's members are in the file
    if lead:
        if outdir.exists():
            import shutil
            shutil.rmtree(outdir)
        os.replace(tmp, outdir)
    return out
