"""Command line: `python -m panagram_b200 index|bitdump ...`

Keeps the surface of ``panagram index`` / ``panagram bitdump`` (``panagram/__main__.py:60-110``,
option names and defaults of ``panagram/index.py:85-138``); argparse stands in for
``simple_parsing``, which is not available in this environment.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np


def _index(args) -> int:
    import os
    from .index import Index, IndexConfig
    if args.gpus > 1 and "RANK" not in os.environ:
        # one process per GPU: re-launch this very command under torchrun (127.0.0.1 rendezvous)
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "-m", "panagram_b200"] + list(args.argv)
        return subprocess.call(cmd)
    if args.gpus > 1:
        import torch
        import torch.distributed as dist
        args.device = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(args.device)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{args.device}"))
    cfg = IndexConfig(k=args.k, cores=args.cores, lowres_step=args.lowres_step, max_bin_kbp=args.max_bin_kbp,
                      min_bin_count=args.min_bin_count, anchor_genomes=args.anchor_genomes, prepare=args.prepare,
                      kmc={"memory": args.kmc_memory, "threads": args.kmc_threads, "use_existing": args.kmc_use_existing})
    idx = Index(args.input, args.prefix, cfg, device=args.device, load_factor=args.load_factor)
    idx.run(log=lambda m: print(m, file=sys.stderr), genome_ranks=args.genome_ranks or None)
    if args.gpus > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def _bitdump(args) -> int:
    """`panagram bitdump <index> <genome> <chr:start-end> [step]` (__main__.py:77-88): rows of the
    pan-kmer bitmap as 0/1 text, read through the .gzi the way Genome._query_bytes does."""
    from . import layout
    d = Path(args.index_dir) / "anchor" / args.genome
    names = [l.split("\t")[0] for l in (Path(args.index_dir) / "samples.tsv").read_text().splitlines()[1:]]
    nbytes = (len(names) + 7) // 8
    chrom, _, span = args.coords.partition(":")
    start, _, end = span.partition("-")
    start, end = int(start), int(end)
    step = args.step
    chrs = [l.split("\t") for l in (d / "chrs.tsv").read_text().splitlines()[1:]]
    off = 0
    for name, _, size, _ in chrs:                      # offsets = cumsum(ceil(size/step)) (index.py:592-604)
        if name == chrom:
            break
        off += (int(size) + step - 1) // step
    else:
        raise SystemExit(f"unknown chromosome {chrom}")
    bstart = nbytes * (off + start // step)
    length = nbytes * ((end - start + step - 1) // step)
    raw = layout.query_bytes(d / f"bitmap.{step}.gz", d / f"bitmap.{step}.gzi", bstart, length)
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(-1, nbytes)
    bits = np.unpackbits(rows, axis=1, bitorder="little")[:, :len(names)]
    for r in bits:
        print("".join(map(str, r.tolist())))
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="panagram_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("index", help="Anchor KMC bitvectors to reference FASTA files to create pan-kmer bitmap")
    p.add_argument("input", metavar="config_file", help="samples TSV (name, fasta[, gff][, anchor])")
    p.add_argument("-o", "--prefix", default=None)
    p.add_argument("-k", "--k", type=int, default=21)
    p.add_argument("-c", "--cores", type=int, default=1)
    p.add_argument("--lowres_step", type=int, default=100)
    p.add_argument("--max_bin_kbp", type=int, default=200)
    p.add_argument("--min_bin_count", type=int, default=100)
    p.add_argument("--anchor_genomes", nargs="*", default=None)
    p.add_argument("-p", "--prepare", action="store_true")
    p.add_argument("--kmc.memory", dest="kmc_memory", type=int, default=8)
    p.add_argument("--kmc.threads", dest="kmc_threads", type=int, default=1)
    p.add_argument("--kmc.use_existing", dest="kmc_use_existing", action="store_true")
    p.add_argument("--device", type=int, default=0, help="CUDA ordinal")
    p.add_argument("--load_factor", type=float, default=0.5, help="k-mer table fill target")
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this node: one process per GPU, k-mer tables sharded by genome")
    p.add_argument("--genome_ranks", type=int, default=0,
                   help="ranks per genome group (default: all GPUs form one group); gpus / genome_ranks groups hold "
                        "replicas of the tables and take the anchors round-robin")
    p.set_defaults(fn=_index)
    b = sub.add_parser("bitdump", help="Query pan-kmer bitmap for debugging")
    b.add_argument("index_dir")
    b.add_argument("genome")
    b.add_argument("coords", help="chr:start-end")
    b.add_argument("step", type=int, nargs="?", default=1)
    b.set_defaults(fn=_bitdump)
    args = ap.parse_args(argv)
    args.argv = list(sys.argv[1:] if argv is None else argv)
    return args.fn(args)


if __name__ == "__main__":
    raise SystemExit(main())
