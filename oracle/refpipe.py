"""TEST INFRASTRUCTURE ONLY — drives the UNMODIFIED reference binaries in
``oracle/_ref`` through the command sequence of the reference's Snakemake DAG
(snakemake itself is not installed here), so tests, fixture generation and
``bench.py``'s CPU-baseline leg can obtain the reference's own output.

Restates (commands and flags verbatim):
  * rule ``kmc_count``  — ``cpp/Snakefile:97-117`` / ``panagram/workflow/Snakefile:81-110``
  * ``get_onehot_tag``  — ``cpp/Snakefile:26-28``           (tag = 1 << (row % 32))
  * ``Index.init_opdefs`` — ``panagram/index.py:407-426``   (<=32 genomes per bitvec DB, ``-ocsum``)
  * rule ``kmc_bitvec`` — ``cpp/Snakefile:62-75``           (``kmc_tools complex opdef``)
  * rule ``anchors``    — ``cpp/Snakefile:35-55``           (``OMP_NUM_THREADS=c run_anchor N . name fasta ...``)

Nothing in the product package imports this module.
"""
from __future__ import annotations

import gzip
import os
import struct
import subprocess
import time
from pathlib import Path

import numpy as np

REF_DIR = Path(__file__).resolve().parent / "_ref"


def have_ref() -> bool:
    return all((REF_DIR / b).exists() for b in ("kmc", "kmc_tools", "run_anchor"))


def _run(cmd, cwd, log, env=None):
    with open(log, "ab") as fh:
        r = subprocess.run(cmd, cwd=cwd, stdout=fh, stderr=subprocess.STDOUT, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"reference command failed ({r.returncode}): {' '.join(map(str, cmd))}\n"
                           f"see {log}")


def kmc_count(root: Path, name: str, fasta: str, idx: int, k: int, threads: int = 1,
              memory: int = 8, fastq: bool = False, timings: dict | None = None):
    """rule kmc_count: kmc (-ci1 -fm | -ci2 -fq) then kmc_tools transform set_counts."""
    (root / "kmc").mkdir(parents=True, exist_ok=True)
    tmp = root / "tmp" / name
    tmp.mkdir(parents=True, exist_ok=True)
    (root / "logs").mkdir(parents=True, exist_ok=True)
    log = root / "logs" / f"kmc.{name}.txt"
    t0 = time.perf_counter()
    _run([str(REF_DIR / "kmc"), f"-k{k}", f"-t{threads}", f"-m{memory}",
          "-ci2" if fastq else "-ci1", "-cs1000", "-fq" if fastq else "-fm",
          str(fasta), f"kmc/{name}.count", f"tmp/{name}"], root, log)
    t1 = time.perf_counter()
    _run([str(REF_DIR / "kmc_tools"), f"-t{threads}", "transform", f"kmc/{name}.count",
          "set_counts", str(1 << (idx % 32)), f"kmc/{name}.onehot"], root, log)
    t2 = time.perf_counter()
    if timings is not None:
        timings["kmc_count_s"] = timings.get("kmc_count_s", 0.0) + (t1 - t0)
        timings["set_counts_s"] = timings.get("set_counts_s", 0.0) + (t2 - t1)


def write_opdefs(root: Path, names: list[str]) -> int:
    """Index.init_opdefs (index.py:407-426)."""
    groups = [names[i:i + 32] for i in range(0, len(names), 32)]
    for i, grp in enumerate(groups):
        with open(root / "kmc" / f"opdef{i}.txt", "w") as fh:
            fh.write("INPUT:\n")
            for n in grp:
                fh.write(f"{n} = kmc/{n}.onehot\n")
            fh.write(f"OUTPUT:\nkmc/bitvec{i} = {grp[0]}")
            for n in grp[1:]:
                fh.write(f" + {n}")
            fh.write("\n-ocsum\n")
    return len(groups)


def kmc_bitvec(root: Path, ndb: int, timings: dict | None = None):
    t0 = time.perf_counter()
    for i in range(ndb):
        _run([str(REF_DIR / "kmc_tools"), "complex", f"kmc/opdef{i}.txt"], root,
             root / "logs" / f"kmc.bitvec{i}.log.txt")
    if timings is not None:
        timings["kmc_bitvec_s"] = time.perf_counter() - t0


def run_anchor(root: Path, n_genomes: int, anchors: list[tuple[str, str]], threads: int = 1,
               timings: dict | None = None):
    """rule anchors (cpp/Snakefile:47-55). `anchors` = [(name, plain-text fasta path)]."""
    args = []
    for name, fasta in anchors:
        (root / "anchor" / name).mkdir(parents=True, exist_ok=True)
        args += [name, str(fasta)]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    t0 = time.perf_counter()
    _run([str(REF_DIR / "run_anchor"), str(n_genomes), "."] + args, root,
         root / "logs" / "anchor.log.txt", env=env)
    if timings is not None:
        timings["run_anchor_s"] = time.perf_counter() - t0


def build_index(root, samples: list[tuple[str, str]], k: int, anchors: list[str] | None = None,
                threads: int = 1, fastq: bool = False) -> dict:
    """Whole reference CPU path: samples = [(name, fasta)] in samples.tsv row order."""
    root = Path(root)
    root.mkdir(parents=True, exist_ok=True)
    timings: dict = {}
    names = [n for n, _ in samples]
    for i, (name, fasta) in enumerate(samples):
        kmc_count(root, name, fasta, i, k, threads=threads, fastq=fastq, timings=timings)
    ndb = write_opdefs(root, names)
    kmc_bitvec(root, ndb, timings)
    anchors = names if anchors is None else anchors
    fa = dict(samples)
    run_anchor(root, len(samples), [(a, fa[a]) for a in anchors], threads=threads, timings=timings)
    return timings


# ---- reading the reference's output (decompressed-bytes parity) -------------

def read_bgzf(path) -> bytes:
    """BGZF is a concatenation of gzip members; gzip.open reads them all."""
    with gzip.open(path, "rb") as fh:
        return fh.read()


def read_gzi(path) -> np.ndarray:
    raw = Path(path).read_bytes()
    (n,) = struct.unpack_from("<Q", raw, 0)
    return np.frombuffer(raw, dtype="<u8", count=2 * n, offset=8).reshape(n, 2)


def read_anchor_dir(d) -> dict:
    d = Path(d)
    out = {"chrs.tsv": (d / "chrs.tsv").read_text(),
           "bitsum.bins.tsv": (d / "bitsum.bins.tsv").read_text()}
    for f in sorted(d.glob("bitmap.*.gz")):
        out[f.name[:-3]] = read_bgzf(f)
    return out
