"""`panagram index`, write side — the caller of the anchoring path.

Mirrors ``Index.__post_init__ / init_config / run`` (``panagram/index.py:196-295,172-191``) and
the anchor branch of the Snakemake ``all`` rule (``panagram/workflow/Snakefile:33-48,151-154``;
``cpp/Snakefile:35-55``) without snakemake: write ``samples.tsv`` + ``config.yaml``, obtain every
genome's k-mer set, then anchor each anchor genome into ``anchor/<name>/``.

k-mer sets come from, in order of preference per genome:
  1. ``kmc/bitvec{i}.kmc_pre`` — merged bitvec databases of an existing reference index
  2. ``kmc/<name>.count.kmc_pre`` / ``.onehot`` — per-genome KMC databases (rule kmc_count)
  3. the genome's FASTA itself, k-mers extracted and inserted on the GPU (equals ``kmc -ci1 -fm``)
Out of scope here: mash distances, UMAPs, GFF annotation.
"""
from __future__ import annotations

import dataclasses
import re
import time
from pathlib import Path

import numpy as np
import yaml

from . import anchor as anchor_mod
from .engine import Engine

NAME_REGEX = r"[A-Za-z0-9_-]+"          # index.py NAME_REGEX (:277)


@dataclasses.dataclass
class IndexConfig:
    """The knobs of ``panagram index`` that reach the anchoring path (index.py:85-138)."""
    k: int = 21
    cores: int = 1
    lowres_step: int = 100
    max_bin_kbp: int = 200
    min_bin_count: int = 100
    max_view_chrs: int = 50
    gff_gene_types: list = dataclasses.field(default_factory=lambda: ["gene"])
    gff_anno_types: list | None = None
    gff_name: str = "Name"
    anchor_genomes: list | None = None
    prepare: bool = False
    kmc: dict = dataclasses.field(default_factory=lambda: {"memory": 8, "threads": 1, "use_existing": False})


def read_samples(path) -> list[dict]:
    """samples TSV with at least `name` and `fasta` columns (index.py:269-283)."""
    lines = [l.rstrip("\n").split("\t") for l in Path(path).read_text().splitlines() if l.strip()]
    hdr = lines[0]
    if "name" not in hdr or "fasta" not in hdr:
        raise ValueError("Input samples must contain 'name' and 'fasta' column headers")
    rows = [dict(zip(hdr, l + [""] * (len(hdr) - len(l)))) for l in lines[1:]]
    bad = [r["name"] for r in rows if not re.fullmatch(NAME_REGEX, r["name"])]
    if bad:
        raise ValueError("Invalid genome names: '" + "', '".join(bad) + f"'\nMust match r'{NAME_REGEX}'.")
    return rows


class Index:
    def __init__(self, samples_tsv, prefix=None, config: IndexConfig | None = None, device: int = 0,
                 load_factor: float = 0.5):
        self.cfg = config or IndexConfig()
        self.input = Path(samples_tsv)
        self.prefix = Path(prefix) if prefix else self.input.parent
        self.device, self.load_factor = device, load_factor
        rows = read_samples(self.input)
        base = self.input.parent
        self.samples = []
        for i, r in enumerate(rows):
            fa = r.get("fasta", "")
            if fa and not Path(fa).is_absolute():
                fa = str((base / fa).resolve())
            self.samples.append({"name": r["name"], "fasta": fa, "gff": r.get("gff", ""), "id": i})
        names = [s["name"] for s in self.samples]
        if self.cfg.anchor_genomes is None:
            if "anchor" in rows[0]:
                self.cfg.anchor_genomes = [r["name"] for r in rows if r["anchor"].strip().lower() in ("true", "1")]
            else:
                self.cfg.anchor_genomes = [s["name"] for s in self.samples if s["fasta"]]
        unknown = [a for a in self.cfg.anchor_genomes if a not in names]
        if unknown:
            raise ValueError(f"anchor genomes not in samples: {unknown}")
        for s in self.samples:
            s["anchor"] = s["name"] in self.cfg.anchor_genomes

    # ---- files the reference's Index(mode="r") needs (Appendix A of SURVEY.md) ----
    def write_config(self):
        self.prefix.mkdir(parents=True, exist_ok=True)
        with open(self.prefix / "samples.tsv", "w") as fh:           # index.py:282-293
            fh.write("name\tfasta\tgff\tid\tanchor\n")
            for s in self.samples:
                fh.write(f"{s['name']}\t{s['fasta']}\t{s['gff']}\t{s['id']}\t{s['anchor']}\n")
        prm = dataclasses.asdict(self.cfg)
        prm["input"] = str(self.input)
        prm["mode"] = None
        with open(self.prefix / "config.yaml", "w") as fh:           # index.py:347-353
            yaml.dump(prm, fh)

    def _kmer_source(self, s) -> tuple[str, str]:
        kmc = self.prefix / "kmc"
        if (kmc / f"bitvec{s['id'] // 32}.kmc_pre").exists():
            return "bitvec", str(kmc / f"bitvec{s['id'] // 32}")
        for kind in ("count", "onehot"):
            if (kmc / f"{s['name']}.{kind}.kmc_pre").exists():
                return "kmc", str(kmc / f"{s['name']}.{kind}")
        if s["fasta"]:
            return "fasta", s["fasta"]
        raise ValueError(f"genome {s['name']}: no FASTA and no KMC database under {kmc}")

    def populate(self, eng: Engine, log=print) -> Engine:
        """Fill the engine's tables with the k-mer sets of the genomes of its shard, then finalize. The per-genome
        tables are build intermediates: each group of 8 genomes is sealed into its group table (one probe answers
        8 genomes) as soon as it is complete and its per-genome tables are freed (pk_engine_seal_group), so the
        resident footprint is the group tables + one group under construction."""
        done_bitvec = set()
        eng.tune(group_only=1)
        filled = set()
        for s in self.samples:
            if not (eng.genome_begin <= s["id"] < eng.genome_end):
                continue
            kind, src = self._kmer_source(s)
            t0 = time.perf_counter()
            if kind == "bitvec":
                if src not in done_bitvec:
                    eng.add_bitvec(32 * (s["id"] // 32), src)
                    done_bitvec.add(src)
            elif kind == "kmc":
                eng.add_kmc(s["id"], src)
            else:
                # what `kmc -ci1 -fm` counts: control bytes dropped, every all-ACGT window
                recs = anchor_mod.parse_fasta(src, strip_cr=True)
                eng.reserve(s["id"], sum(max(q.size - self.cfg.k + 1, 0) for _, q in recs))
                for _, q in recs:
                    eng.add_sequence(s["id"], q)
            log(f"k-mer set of {s['name']} from {kind} ({time.perf_counter() - t0:.2f}s)")
            filled.add(s["id"])
            gl = s["id"] - eng.genome_begin
            u0 = eng.genome_begin + gl // 8 * 8
            if kind != "bitvec" and all(g in filled for g in range(u0, min(u0 + 8, eng.genome_end))):
                eng.seal_group(gl // 8)
        t0 = time.perf_counter()
        eng.finalize()
        log(f"group tables complete ({time.perf_counter() - t0:.2f}s)")
        return eng

    def _engine_kw(self) -> dict:
        return dict(lowres_step=self.cfg.lowres_step, max_bin_kbp=self.cfg.max_bin_kbp,
                    min_bin_count=self.cfg.min_bin_count, load_factor=self.load_factor)

    def build_engine(self, genome_begin: int = 0, genome_end: int | None = None, log=print) -> Engine:
        t0 = time.perf_counter()
        eng = Engine(self.cfg.k, len(self.samples), genome_begin, genome_end, device=self.device, **self._engine_kw())
        log(f"engine on cuda:{self.device} ({time.perf_counter() - t0:.2f}s)")
        return self.populate(eng, log)

    def write_genome_dist(self, records, frac: float, log=print):
        """genome_dist.tsv (rule mash_triangle, workflow/Snakefile:139-149; read at `panagram view` start-up,
        view.py:64 -> figs.py:50-59) from the k-mer samples of all engines."""
        from . import layout
        t0 = time.perf_counter()
        names = [s["name"] for s in self.samples]
        inter = layout.pair_counts(records, len(names))
        (self.prefix / "genome_dist.tsv").write_text(layout.genome_dist_tsv(names, inter, self.cfg.k))
        log(f"genome_dist.tsv: Jaccard / Mash distances of {len(names)} genomes from {int(np.trace(inter))} sampled k-mer "
            f"memberships (sampling fraction {frac:.3g}) ({time.perf_counter() - t0:.2f}s)")

    DIST_SAMPLE_TARGET = 1_000_000       # distinct k-mers aimed for in the sample behind genome_dist.tsv (mash: 10000 per genome)

    def run(self, log=print, genome_ranks: int | None = None) -> dict:
        """Index.run (index.py:172-191): config, then — unless --prepare — the anchor rule for every
        anchor genome (workflow/Snakefile:33-48; cpp/Snakefile:35-55 runs them in one process,
        `#pragma omp parallel for` over anchors, cpp/anchor.cpp:217).

        Under torch.distributed (one process per GPU, `python -m panagram_b200 index --gpus R`) the run is
        genome-sharded: world = Rg x Rp, every group of Rg ranks holds the tables of all genomes between them
        (rank g the g-th shard) and the Rp groups take the anchors round-robin. Every rank returns the summaries
        of the anchors its group handled."""
        import os
        world = int(os.environ.get("WORLD_SIZE", "1"))
        dist = None
        if world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("WORLD_SIZE > 1 but torch.distributed is not initialised (use `index --gpus R`)")
        rank = dist.get_rank() if dist else 0
        if rank == 0:
            self.write_config()
        if self.cfg.prepare:
            return {}
        names = [s["name"] for s in self.samples]
        anchors = [s for s in self.samples if s["anchor"]]
        out = {}
        if not dist:
            eng = self.build_engine(log=log)
            total = sum((eng.group_stats(u) or {"n_keys": 0})["n_keys"] for u in range((eng.n_local + 7) // 8))
            frac = min(1.0, self.DIST_SAMPLE_TARGET / max(total, 1))
            t0 = time.perf_counter()
            keys, tags = eng.sample_kmers(frac)
            log(f"k-mer sample for genome_dist.tsv: {keys.size} k-mers ({time.perf_counter() - t0:.2f}s)")
            self.write_genome_dist([(keys, tags, 0)], frac, log)
            for s in anchors:
                t0 = time.perf_counter()
                out[s["name"]] = anchor_mod.anchor_fasta(eng, s["name"], s["fasta"], self.prefix / "anchor" / s["name"],
                                                         genome_names=names, threads=max(1, self.cfg.cores))
                log(f"Anchored {s['name']}: {out[s['name']]['positions']} positions ({time.perf_counter() - t0:.2f}s)")
            eng.close()
            return out
        from . import sharded
        sh = sharded.ShardedAnchorer(self.cfg.k, len(self.samples), rank, world, self.device, genome_ranks=genome_ranks,
                                     **self._engine_kw())
        qlog = log if sh.gi == 0 else (lambda m: None)
        self.populate(sh.engine, qlog)
        dist.barrier()                                   # rank 0 has written the config; every shard is built
        # genome_dist.tsv: the ranks of genome group 0 sample their shards with one common fraction, rank 0 merges
        import torch
        eng = sh.engine
        tot = torch.tensor([float(sum((eng.group_stats(u) or {"n_keys": 0})["n_keys"] for u in range((eng.n_local + 7) // 8)))
                            if sh.pi == 0 else 0.0], dtype=torch.float64, device=sh.dev)
        dist.all_reduce(tot)
        frac = min(1.0, self.DIST_SAMPLE_TARGET / max(float(tot.item()), 1.0))
        mine = None
        if sh.pi == 0:
            keys, tags = eng.sample_kmers(frac)
            mine = (keys, tags, sh.begin)
        allrec = [None] * world if rank == 0 else None
        dist.gather_object(mine, allrec, dst=0)
        if rank == 0:
            self.write_genome_dist([r for r in allrec if r is not None], frac, qlog)
        for j, s in enumerate(anchors):
            if j % sh.rp != sh.pi:
                continue
            t0 = time.perf_counter()
            out[s["name"]] = sharded.anchor_fasta_sharded(sh, s["name"], s["fasta"], self.prefix / "anchor" / s["name"],
                                                          genome_names=names)
            qlog(f"Anchored {s['name']} on ranks {sh.pi * sh.rg}..{sh.pi * sh.rg + sh.rg - 1}: "
                 f"{out[s['name']]['positions']} positions ({time.perf_counter() - t0:.2f}s; {sh.last})")
        sh.close_p2p()
        dist.barrier()
        sh.engine.close()
        return out


def make_bins_bits(anchor_dir, n_genomes: int) -> np.ndarray:
    """scripts/make_bins_bits.py:34-59,97 as a layout check: per bin, the number of low-res rows
    with popcount == 1 and popcount == N, i.e. columns `1` and `N` of a histogram over bitmap.100
    rows. Returns [n_rows_total_bins, 2] from the written bitmap.100 (decompressed on the host)."""
    from . import layout
    d = Path(anchor_dir)
    nbytes = (n_genomes + 7) // 8
    rows = np.frombuffer(layout.read_bgzf(d / "bitmap.100.gz"), dtype=np.uint8).reshape(-1, nbytes)
    pc = np.unpackbits(rows, axis=1).sum(axis=1)
    chrs = [l.split("\t") for l in (d / "chrs.tsv").read_text().splitlines()[1:]]
    out, off = [], 0
    for _, _, size, _ in chrs:
        n = (int(size) + 99) // 100
        seg = pc[off:off + n]
        for b in range(0, n, 2000):          # 200 kb bins of step-100 rows
            w = seg[b:b + 2000]
            out.append(((w == 1).sum(), (w == n_genomes).sum()))
        off += n
    return np.array(out, dtype=np.int64)
