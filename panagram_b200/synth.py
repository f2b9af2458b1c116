"""Deterministic synthetic pan-genomes for tests and bench.py.

Ancestor: iid uniform ACGT of length L (seed = 20260000 + config index). Genome g
is the ancestor with independent per-base substitutions at rate
mu_g = 0.002 * (1 + g mod mu_period) (stream seed + 1 + g; mu_period = 16 unless a
caller says otherwise: bench.py's weak-scaling series uses the shard width, so that every GPU's shard has the
divergence ladder — hence the table size and the work — of configs[1] itself), split into C equal
chromosomes chr1..chrC. Every genome carries one N-run of 1000 at offset
L/(2C) of chr1 and a lowercase stretch of 10000 at the start of chr2, so the
non-ACGT and case-folding paths are exercised (kmer_api.h:264-275 semantics).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def ancestor_codes(length: int, seed: int) -> np.ndarray:
    """2-bit codes (uint8 in 0..3) of the ancestor."""
    return np.random.default_rng(seed).integers(0, 4, size=length, dtype=np.uint8)


def genome_codes(anc: np.ndarray, g: int, seed: int, mu_period: int = 16) -> np.ndarray:
    """Genome g's 2-bit codes: ancestor + substitutions at rate mu_g."""
    mu = 0.002 * (1 + g % mu_period)
    rng = np.random.default_rng(seed + 1 + g)
    n = anc.shape[0]
    nsub = rng.binomial(n, mu)
    pos = rng.integers(0, n, size=nsub)
    shift = rng.integers(1, 4, size=nsub, dtype=np.uint8)
    out = anc.copy()
    out[pos] = (anc[pos] + shift) & 3
    return out


def genome_chroms(anc: np.ndarray, g: int, seed: int, n_chroms: int = 5,
                  n_run: int = 1000, lower_run: int = 10000, mu_period: int = 16) -> list[tuple[str, np.ndarray]]:
    """[(name, ASCII uint8 array)] for genome g."""
    codes = genome_codes(anc, g, seed, mu_period)
    asc = _ACGT[codes]
    L = asc.shape[0]
    clen = L // n_chroms
    chroms = []
    for c in range(n_chroms):
        end = L if c == n_chroms - 1 else (c + 1) * clen
        s = asc[c * clen:end].copy()
        if c == 0 and n_run:
            o = min(L // (2 * n_chroms), max(0, s.shape[0] - n_run))
            s[o:o + n_run] = ord("N")
        if c == 1 and lower_run:
            s[:lower_run] |= 0x20
        chroms.append((f"chr{c + 1}", s))
    return chroms


def write_fasta(path, chroms, width: int = 60) -> None:
    with open(path, "wb") as fh:
        for name, s in chroms:
            fh.write(b">" + name.encode() + b"\n")
            n = s.shape[0]
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                fh.write(body.tobytes())
            if n > full:
                fh.write(s[full:].tobytes() + b"\n")


def make_pangenome(outdir, n_genomes: int, length: int, seed: int, n_chroms: int = 5,
                   n_run: int = 1000, lower_run: int = 10000) -> list[tuple[str, str]]:
    """Write g0.fa .. g{n-1}.fa under outdir; returns [(name, path)] in sample order."""
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    anc = ancestor_codes(length, seed)
    samples = []
    for g in range(n_genomes):
        p = outdir / f"g{g}.fa"
        write_fasta(p, genome_chroms(anc, g, seed, n_chroms, n_run, lower_run))
        samples.append((f"g{g}", str(p)))
    return samples
