"""`anchor_fasta`: one anchor genome -> its ``anchor/<name>/`` directory.

Host-side mirror of ``KMCdb::anchor_fasta`` (``cpp/anchor.cpp:37-109``) and of
``Genome.run_anchor`` (``panagram/index.py:1012-1097``): parse the FASTA, run every
chromosome through the engine (GPU), stream the rows into BGZF, and emit
``chrs.tsv``, ``bitsum.bins.tsv``, ``total_paircounts.csv`` and the two UMAP CSVs.
"""
from __future__ import annotations

import gzip
import os
import shutil
from pathlib import Path

import numpy as np

from . import layout
from .engine import Engine


def parse_fasta(path, strip_cr: bool = False) -> list[tuple[str, np.ndarray]]:
    """[(name, uint8 sequence)] through the native reader (pk_fasta_open: one memchr pass into page-locked memory;
    gzip / bgzip input inflated with zlib first). `parse_fasta_numpy` restates the same rules in numpy for the tests
    (tests/test_host.py compares the two)."""
    import ctypes as C
    import weakref
    from . import _lib
    path = str(path)
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.pk_fasta_open(path.encode(), int(bool(strip_cr)), C.byref(h))
    _lib.check(rc)
    n = L.pk_fasta_n_records(h)
    recs, base, total = [], None, 0
    for i in range(n):
        name, seq, ln = C.c_char_p(), C.c_void_p(), C.c_uint64()
        _lib.check(L.pk_fasta_record(h, i, C.byref(name), C.byref(seq), C.byref(ln)))
        if base is None:
            base = seq.value or 0
        recs.append((name.value.decode(errors="replace"), (seq.value or 0) - base, ln.value))
        total = max(total, (seq.value or 0) - base + ln.value)
    if not recs or not base:
        L.pk_fasta_close(h)
        return [(nm, np.zeros(0, dtype=np.uint8)) for nm, _, _ in recs]
    buf = (C.c_uint8 * max(total, 1)).from_address(base)
    weakref.finalize(buf, L.pk_fasta_close, h)      # the arrays below keep `buf` (hence the handle) alive
    whole = np.frombuffer(buf, dtype=np.uint8)
    return [(nm, whole[o:o + ln]) for nm, o, ln in recs]


def parse_fasta_numpy(path, strip_cr: bool = False) -> list[tuple[str, np.ndarray]]:
    """[(name, uint8 sequence)] with the C++ reference's rules (cpp/anchor.cpp:74-100):
    the record name is the header up to the first space, sequence lines are
    concatenated verbatim (a trailing \\r stays in the sequence unless strip_cr,
    which drops every control byte (< 32) the way `kmc -fm` reads a FASTA
    (kmc_core/splitter.cpp: "c < 32 // newliners") and Biopython strips line ends,
    index.py:922-930). gzip-aware like Genome.iter_fasta."""
    path = str(path)
    raw = Path(path).read_bytes()
    if path.endswith(".gz") or path.endswith(".bgz") or raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    buf = np.frombuffer(raw, dtype=np.uint8)
    nl = np.flatnonzero(buf == 10)
    starts = np.concatenate(([0], nl + 1))
    ends = np.concatenate((nl, [buf.size]))
    if starts.size and starts[-1] >= buf.size:          # file ends with \n: no empty last line
        starts, ends = starts[:-1], ends[:-1]
    if strip_cr:
        cr = (ends > starts) & (buf[np.maximum(ends - 1, 0)] == 13)
        ends = ends - cr
    is_hdr = (ends > starts) & (buf[np.minimum(starts, buf.size - 1)] == ord(">")) if buf.size else np.zeros(0, bool)
    hdr_idx = np.flatnonzero(is_hdr)
    recs = []
    keep = np.ones(buf.size, dtype=bool)
    keep[nl] = False
    if strip_cr:
        keep[buf < 32] = False
    for j, hi in enumerate(hdr_idx):
        name = raw[starts[hi] + 1:ends[hi]].split(b" ")[0].decode()
        if strip_cr:
            name = name.split()[0] if name.split() else ""
        lo = ends[hi] + 1
        up = starts[hdr_idx[j + 1]] if j + 1 < hdr_idx.size else buf.size
        lo = min(lo, up)
        seg = buf[lo:up]
        recs.append((name, np.ascontiguousarray(seg[keep[lo:up]])))
    return recs


def write_text_files(outdir, name: str, chrom_names: list[str], nks: list[int], binlens: list[int], hists: list,
                     col_sums, pc_counts: list, n_genomes: int, step: int, genome_names: list[str] | None,
                     umap_bin_size: int = 100000) -> None:
    """Everything of an anchor directory but the bitmaps: chrs.tsv, bitsum.bins.tsv (cpp/anchor.cpp:57-69,84-85,
    184-189), total_paircounts.csv (index.py:1068-1074) and chrom_umaps.csv / genome_umap.csv (Genome.write_umaps,
    index.py:1107-1131) from the pair-count bins `pc_counts` (per chromosome [bins, N] counts of low-res rows with
    each genome's bit set, reduced on the GPU: Engine.anchor_paircount_bins)."""
    outdir = Path(outdir)
    chrom_rows, genome_parts = [], []
    for cname, cnt in zip(chrom_names, pc_counts):
        starts, frac = layout.paircount_frac(cnt, umap_bin_size)
        chrom_rows += layout.umap_rows(cname, starts, frac, umap_bin_size)
        genome_parts.append((cname, starts, frac))
    (outdir / "chrom_umaps.csv").write_text(layout.umaps_csv(chrom_rows))
    if genome_parts:
        g_starts = np.concatenate([p[1] for p in genome_parts])
        g_frac = np.concatenate([p[2] for p in genome_parts])
        g_names = np.concatenate([[p[0]] * len(p[1]) for p in genome_parts])
        g_rows = layout.umap_rows("", g_starts, g_frac, umap_bin_size)
        g_rows = [(str(n),) + r[1:] for n, r in zip(g_names, g_rows)]
    else:
        g_rows = []
    (outdir / "genome_umap.csv").write_text(layout.umaps_csv(g_rows))
    (outdir / "chrs.tsv").write_text(layout.chrs_tsv(list(zip(chrom_names, nks))))
    (outdir / "bitsum.bins.tsv").write_text(layout.bins_tsv(n_genomes, list(zip(binlens, hists))))
    if genome_names is not None and name in genome_names:
        (outdir / "total_paircounts.csv").write_text(layout.paircounts_csv(genome_names, col_sums, name))


def anchor_fasta(engine: Engine, name: str, fasta, outdir, genome_names: list[str] | None = None,
                 bgzf_level: int = 6, threads: int | None = None, strip_cr: bool = False,
                 bgzf: str = "gpu", umap_bin_size: int = 100000) -> dict:
    """Anchor one genome and write its directory. Returns summary numbers.

    The engine must hold all N genomes (single-GPU layout; sharded.anchor_fasta_sharded is the multi-GPU
    form). Chromosomes shorter than k + min_bin_count - 1 have no defined bins in the reference
    (cpp/anchor.cpp:116-120 divides by zero): they raise ValueError here.
    bgzf="gpu": the .gz/.gzi files are compressed on the GPU (pk_anchor_genome_bgzf) and arrive as file
    images; bgzf="zlib": raw rows come back and are deflated by zlib on a host thread pool
    (`bgzf_level`, `threads`). Both decompress to the same bytes.
    The directory is built as `<outdir>.tmp` and renamed when complete (a killed run leaves no
    half-written anchor directory behind).
    """
    final = Path(outdir)
    outdir = final.with_name(final.name + ".tmp")
    if outdir.exists():
        shutil.rmtree(outdir)
    outdir.mkdir(parents=True)
    threads = threads or min(32, os.cpu_count() or 1)
    step = engine.lowres_step
    recs = parse_fasta(fasta, strip_cr=strip_cr)
    for cname, seq in recs:
        nk = seq.size - engine.k + 1
        if nk < 1 or engine.bin_len(nk) == 0:
            shutil.rmtree(outdir)
            raise ValueError(f"{fasta}: chromosome {cname!r} has {max(nk, 0)} k-mers; the reference "
                             "needs at least min_bin_count (cpp/anchor.cpp:116-120)")
    if bgzf == "gpu":
        res = engine.anchor_genome_bgzf([s for _, s in recs])             # the whole genome in one batch
        for key, fn in (("gz", "bitmap.1.gz"), ("gzi", "bitmap.1.gzi"), ("gz_low", f"bitmap.{step}.gz"),
                        ("gzi_low", f"bitmap.{step}.gzi")):
            with open(outdir / fn, "wb") as fh:
                fh.write(res[key].data)
    elif bgzf == "zlib":
        w1 = layout.BgzfWriter(outdir / "bitmap.1.gz", bgzf_level, threads)
        wl = layout.BgzfWriter(outdir / f"bitmap.{step}.gz", bgzf_level, threads)
        res = engine.anchor_genome([s for _, s in recs], pinned=True)
        for r in res["chroms"]:
            w1.write(r["bitmap1"])
            wl.write(r["low"])
        w1.close(outdir / "bitmap.1.gzi")
        wl.close(outdir / f"bitmap.{step}.gzi")
    else:
        shutil.rmtree(outdir)
        raise ValueError(f"bgzf={bgzf!r}: expected 'gpu' or 'zlib'")
    col = res["col_sums"]
    nks = [r["nkmers"] for r in res["chroms"]]
    pc_counts = engine.anchor_paircount_bins(nks, umap_bin_size)      # from the low-res rows still on the device
    write_text_files(outdir, name, [c for c, _ in recs], nks, [r["binlen"] for r in res["chroms"]],
                     [r["bin_hist"] for r in res["chroms"]], col, pc_counts, engine.n_local, step, genome_names, umap_bin_size)
    if final.exists():
        shutil.rmtree(final)
    os.replace(outdir, final)
    return {"positions": sum(nks), "chroms": len(nks), "col_sums": col}
