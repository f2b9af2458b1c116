#!/usr/bin/env python3
"""Regenerates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference compiled into oracle/_ref by
`make -C oracle ref`):   python tests/golden/make_golden.py

Fixtures (each a .tar.gz unpacked by tests/conftest.py):
  kat_k17    the known-answer test of KMC/tests/py_kmc_api/test_py_kmc_file.py:
             reads fixture (:28-32), k=17 (:23), kmc flags of _run_kmc (:113-132),
             query read of test_get_counters_for_read (:177). Holds the KMC2 DB
             written by the reference `kmc`, its KMC1 form (`kmc_tools transform
             sort`), and the counters returned by the reference's own
             py_kmc_api.KMCFile.GetCountersForRead on both.
  pan3_k21   3 genomes x 3 chromosomes, k=21, all anchors: N-run, lowercase,
             IUPAC codes, a CRLF line, a short last chromosome; the whole
             reference CPU path (kmc -> set_counts -> complex -> run_anchor).
  pan35_k31  35 genomes (nbytes=5, two bitvec DBs: the N > 32 byte-interleave
             path of cpp/anchor.cpp:139-162), k=31, 2 anchors.
Expected outputs are the reference's: decompressed bitmap.1 / bitmap.100 and the
text of chrs.tsv / bitsum.bins.tsv written by the unmodified cpp/anchor.cpp.
"""
from __future__ import annotations

import io
import json
import shutil
import subprocess
import sys
import tarfile
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
from oracle import refpipe  # noqa: E402
from panagram_b200 import synth  # noqa: E402

HERE = Path(__file__).resolve().parent
REF = refpipe.REF_DIR


def tar_dir(src: Path, name: str):
    out = HERE / f"{name}.tar.gz"
    with tarfile.open(out, "w:gz", compresslevel=9) as tf:
        for p in sorted(src.rglob("*")):
            if p.is_file():
                ti = tf.gettarinfo(str(p), arcname=f"{name}/{p.relative_to(src)}")
                ti.mtime = 0
                ti.uid = ti.gid = 0
                ti.uname = ti.gname = ""
                with open(p, "rb") as fh:
                    tf.addfile(ti, fh)
    print(f"wrote {out} ({out.stat().st_size} bytes)")


def make_kat(tmp: Path):
    import py_kmc_api as pka  # the reference's pybind11 module, built into oracle/_ref

    d = tmp / "kat_k17"
    d.mkdir()
    reads = ("GGCATTGCATGCAGTNNCAGTCATGCAGTCAGGCAGTCATGGCATGCAACGACGATCAGTCATGGTCGAG",
             "GGCATTGCATGCAGTNNCAGTCATGCAGTCAGGCAGTCATGGCATGCAACGACGATCAGTCATGGTCGAG",
             "GTCGATGCATCGATGCTGATGCTGCTGTGCTAGTAGCGTCTGAGGGCTA")
    query = "GGCATTGCATGCAGTNNCAGTCATGCAGTCAGGCAGTCATGGCATGCGTAAACGACGATCAGTCATGGTCGAG"
    with open(d / "input.fastq", "w") as fh:
        for r in reads:
            fh.write(f"@TEST\n{r}\n+TEST\n{'I' * len(r)}\n")
    subprocess.check_call([str(REF / "kmc"), "-ci1", "-k17", "-m2", "-p9", "input.fastq", "kmc_db", "."],
                          cwd=d, stdout=subprocess.DEVNULL)
    subprocess.check_call([str(REF / "kmc_tools"), "transform", "kmc_db", "sort", "kmc_db_sorted"],
                          cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    res = {}
    for name in ("kmc_db", "kmc_db_sorted"):
        f = pka.KMCFile()
        assert f.OpenForRA(str(d / name))
        v = pka.CountVec()
        assert f.GetCountersForRead(query, v)
        res[name] = list(v.value)
        short = pka.CountVec()
        assert not f.GetCountersForRead("ACGT", short)
        assert list(short.value) == []
        f.Close()
    assert res["kmc_db"] == res["kmc_db_sorted"]
    # listing, via the reference's kmc_dump
    subprocess.check_call([str(REF / "kmc_dump"), "kmc_db", "dump.txt"], cwd=d, stdout=subprocess.DEVNULL)
    json.dump({"k": 17, "reads": reads, "query": query, "counters": res["kmc_db"]},
              open(d / "kat.json", "w"), indent=1)
    tar_dir(d, "kat_k17")


def spice(chroms, g):
    """IUPAC codes, a stray lowercase n, so non-ACGT handling is pinned."""
    out = []
    for ci, (name, s) in enumerate(chroms):
        s = s.copy()
        if ci == 2 and s.shape[0] > 700:
            s[300] = ord("R"); s[301] = ord("y"); s[650] = ord("n")
        out.append((name + (" description text" if ci == 1 else ""), s))
    return out


def write_fasta_crlf(path, chroms, crlf_chrom=1):
    """synth.write_fasta, but one line of one chromosome ends in CRLF: cpp/anchor.cpp
    keeps the \\r in the sequence (:91-93), where it invalidates the windows over it."""
    buf = io.BytesIO()
    for ci, (name, s) in enumerate(chroms):
        buf.write(b">" + name.encode() + b"\n")
        for li, o in enumerate(range(0, s.shape[0], 60)):
            buf.write(s[o:o + 60].tobytes())
            buf.write(b"\r\n" if (ci == crlf_chrom and li == 3) else b"\n")
    Path(path).write_bytes(buf.getvalue())


def make_pan(tmp: Path, name: str, n: int, length: int, k: int, seed: int, anchors: list[int], n_chroms=3):
    d = tmp / name
    (d / "fasta").mkdir(parents=True)
    anc = synth.ancestor_codes(length, seed)
    samples = []
    for g in range(n):
        chroms = synth.genome_chroms(anc, g, seed, n_chroms=n_chroms, n_run=40, lower_run=150)
        # a short last chromosome (still >= 100 k-mers, the reference's lower limit)
        chroms[-1] = (chroms[-1][0], chroms[-1][1][: 260 + 7 * g])
        chroms = spice(chroms, g)
        p = d / "fasta" / f"g{g}.fa"
        write_fasta_crlf(p, chroms)
        samples.append((f"g{g}", str(p)))
    idx = d / "index"
    refpipe.build_index(idx, samples, k, anchors=[f"g{a}" for a in anchors], threads=2)
    exp = d / "expected"
    for a in anchors:
        r = refpipe.read_anchor_dir(idx / "anchor" / f"g{a}")
        (exp / f"g{a}").mkdir(parents=True)
        for key, val in r.items():
            mode = "w" if isinstance(val, str) else "wb"
            open(exp / f"g{a}" / key, mode).write(val)
    # keep: fasta, bitvec DBs, and for the small set the per-genome DBs as well
    keep = d / "kmc"
    keep.mkdir()
    pats = ["bitvec*"] + (["*.count.*", "*.onehot.*"] if n <= 3 else ["g0.count.*", "g34.onehot.*"])
    for pat in pats:
        for f in (idx / "kmc").glob(pat):
            shutil.copy(f, keep / f.name)
    shutil.rmtree(idx)
    json.dump({"n_genomes": n, "k": k, "anchors": [f"g{a}" for a in anchors], "seed": seed,
               "length": length, "names": [s[0] for s in samples]}, open(d / "meta.json", "w"), indent=1)
    tar_dir(d, name)


def main():
    assert refpipe.have_ref(), "build the reference first: make -C oracle ref"
    with tempfile.TemporaryDirectory() as t:
        tmp = Path(t)
        make_kat(tmp)
        make_pan(tmp, "pan3_k21", 3, 9000, 21, 4242, [0, 1, 2])
        make_pan(tmp, "pan35_k31", 35, 3000, 31, 777, [0, 34])


if __name__ == "__main__":
    main()
