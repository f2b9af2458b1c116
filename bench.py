#!/usr/bin/env python3
"""bench.py — anchored k-mers/sec (positions x genomes) building the pan-kmer bitmap.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                  [--genome-ranks R] [--exchange slice|nccl] [--group-tables 0|1] [--index-e2e W]

One "step" = one pass of the hot path over one batch of input: every k-mer position of the workload's anchor
genome(s) probed against every genome's k-mer set and the N-bit rows written.

Workloads (BASELINE.json `configs`; synthetic genomes per SURVEY.md §8d, panagram_b200/synth.py):
  configs1   8 x 135 Mbp, k=21, 1 anchor — the configuration the metric is quoted on (default)
  configs2   32 x 150 Mbp, k=21, 1 anchor
  configs3   64 x 150 Mbp, k=31, 4 anchors — ONE workload at every N (strong scaling, `--gpus 1,2,4,8`)
  configs3s  one GPU's shard of configs[3]: 8 x 150 Mbp, k=31
  configs4s  one GPU's shard of configs[4]: 16 x 250 Mbp, k=21
  small / small3   plumbing checks of the weak / strong forms
Weak workloads at N GPUs: the genomes are sharded by genome (8 genomes' tables per GPU, 8N genomes in total), every
rank probes all anchor positions against its shard into its plane, and the position-split exchange assembles the
rows: rank r builds only its slice of the rows, reading that slice of all N planes in place over NVLink (one kernel,
pk_gather_slice_device) — the slice it goes on to reduce, compress and store. Strong workloads: the world is a grid
of Rp genome groups x Rg ranks (`--genome-ranks`; default: as few ranks per group as the tables allow); groups hold
replicas of the tables and take the anchors round-robin, ranks of a group shard the genomes and exchange as above.

`value`     device-resident: packed anchor(s) already in HBM, CUDA events around the probe stage
            (partition + probe + un-permute kernels [+ barrier + exchange kernel at Rg > 1; the exchange of an anchor
            runs on a side stream under the probe of the next one, the last one inside the timed region]).
`e2e`       the public call a user makes with HOST buffers (N=1: Engine.anchor_genome -> pk_anchor_genome;
            N>1: ShardedAnchorer.anchor_genome): ASCII in from pinned memory, bitmap rows / low-res rows /
            histograms / column sums out to host memory, all copies inside the timed region.
`e2e_files` the same with the bitmaps delivered as BGZF file images deflated on the GPU.
`roofline`  the dominant kernel (K3, the hash probe). `frac`: the DESIGN's algorithmic bytes (one 32 B sector per
            position and 8-genome group table + the streams) / its CUDA-event duration / measured HBM peak;
            `survey_accounting`: SURVEY §8d's figure (32 B per position AND genome), which a group table undercuts
            8-fold, so that fraction exceeds 1; `traffic` / `physical`: DRAM bytes of one launch measured LIVE by an
            ncu child of this very command (null when ncu cannot run); `stage`: all kernels of the probe stage.
`--impl reference`  the reference's own CPU implementation (oracle/_ref/run_anchor = the unmodified cpp/anchor.cpp +
            KMC API) on the SAME workload at full size: the KMC databases of all genomes are built once (cached under
            /tmp), every step runs `run_anchor` over the whole anchor genome, cut into one FASTA per host core because
            the reference parallelises over anchors only (cpp/anchor.cpp:217).
`--index-e2e W`     instead: the whole `panagram index` run, FASTA files on disk -> anchor directories,
            next to the reference's pipeline on the same files, outputs compared byte for byte.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "configs1": dict(n_per_gpu=8, length=135_000_000, k=21, seed=20260001, anchors=[0],
                     name="configs[1]: 8 synthetic 135 Mbp genomes, k=21, 1 anchor"),
    "configs2": dict(n_per_gpu=32, length=150_000_000, k=21, seed=20260002, anchors=[0],
                     name="configs[2]: 32 synthetic 150 Mbp genomes, k=21, 1 anchor"),
    "configs3": dict(n_total=64, length=150_000_000, k=31, seed=20260003, anchors=[0, 16, 32, 48],
                     name="configs[3]: 64 synthetic 150 Mbp genomes, k=31, 4 anchors"),
    "configs3s": dict(n_per_gpu=8, length=150_000_000, k=31, seed=20260003, anchors=[0],
                      name="configs[3], one GPU's shard: 8 of 64 synthetic 150 Mbp genomes, k=31 (64-bit slots), 1 of 4 anchors"),
    "configs4s": dict(n_per_gpu=16, length=250_000_000, k=21, seed=20260004, anchors=[0],
                      name="configs[4], one GPU's shard: 16 of 128 synthetic 250 Mbp genomes, k=21, 1 anchor"),
    "small": dict(n_per_gpu=8, length=8_000_000, k=21, seed=20260009, anchors=[0],
                  name="small: 8 synthetic 8 Mbp genomes, k=21, 1 anchor"),
    "small3": dict(n_total=16, length=6_000_000, k=31, seed=20260008, anchors=[0, 4, 8, 12],
                   name="small3: 16 synthetic 6 Mbp genomes, k=31, 4 anchors"),
}
TABLE_BYTES_PER_GENOME_BASE = 12.0      # group tables, k > 24 (fill 0.35): sizing estimate for the grid choice


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


NVLINK_PEAK_GBS = 770.0       # measured peer copy per direction per GPU on this pool (B200_PROFILING.md; nominal 900)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def mu_period_of(wl):
    """Weak workloads: genome g of shard r diverges from the ancestor like genome g of shard 0 (independent
    mutations), so every GPU's shard is a pan-genome with the statistics of the named configuration — the same
    number of distinct k-mers, the same table size, the same work: weak scaling in the strict sense. Strong
    workloads keep the generator's default ladder of 16 rates."""
    return min(16, wl["n_per_gpu"]) if "n_per_gpu" in wl else 16


def make_genome(anc, g, seed, mu_period=16):
    from panagram_b200 import synth
    return [s for _, s in synth.genome_chroms(anc, g, seed, mu_period=mu_period)]


REF_MAX_BASES = 1_100_000_000      # the reference arm builds KMC databases for at most this many bases in total


def reference_workload(wl, n):
    """The workload the reference arm runs: the same generator, the same genome count; the genome length is the
    workload's own while n x length stays within REF_MAX_BASES (configs[1]: 8 x 135 Mbp, in full), shortened beyond
    that (the KMC build alone would take tens of minutes and ~1 GB of /tmp per genome)."""
    length = min(wl["length"], REF_MAX_BASES // n)
    return dict(wl, length=length), length == wl["length"]


def n_genomes_of(wl, world):
    return wl["n_total"] if "n_total" in wl else wl["n_per_gpu"] * max(1, world)


def workload_config(wl, args, world):
    """The `config` object of the JSON line: what was run, identical in both arms."""
    n = n_genomes_of(wl, world)
    strong = "n_total" in wl
    return {"workload": wl["name"] + (f" x{world} genome shards ({n} genomes; genome g of every shard has the divergence rate of genome g of "
                                        f"shard 0, so every shard has configs-sized tables)" if world > 1 and not strong else ""),
            "k": wl["k"], "n_genomes": n, "anchors": len(wl["anchors"]), "genome_length": wl["length"],
            "positions_per_step": None}


# ------------------------------------------------------------------------ reference arm
def reference_setup(wl, n, cores, log=lambda m: None):
    """KMC databases of all n genomes of the workload at FULL size + the anchor genomes cut into one FASTA per host
    core, cached under /tmp (keyed by the generator's parameters) so that the reference arm and our arm's
    cpu_baseline leg, run back to back on one box, build them once. Returns (root, [(name, fasta)], positions,
    timings)."""
    from oracle import refpipe
    from panagram_b200 import synth
    k = wl["k"]
    root = Path(tempfile.gettempdir()) / f"pk_refcache_n{n}_m{mu_period_of(wl)}_L{wl['length']}_s{wl['seed']}_k{k}_a{'-'.join(map(str, wl['anchors']))}_c{min(cores, n)}"
    meta_p = root / "meta.json"
    if meta_p.exists():
        m = json.loads(meta_p.read_text())
        return root, [tuple(x) for x in m["pieces"]], m["positions"], m["timings"]
    if root.exists():
        shutil.rmtree(root)
    (root / "fa").mkdir(parents=True)
    timings = {}
    anc = synth.ancestor_codes(wl["length"], wl["seed"])
    names = [f"g{g}" for g in range(n)]
    pieces, positions = [], 0
    # run_anchor takes at most N anchors (cpp/anchor.cpp:208-211) and runs one thread per anchor
    per_anchor = max(1, min(cores, n) // len(wl["anchors"]))
    for g in range(n):
        chroms = synth.genome_chroms(anc, g, wl["seed"], mu_period=mu_period_of(wl))
        fa = root / "fa" / f"g{g}.fa"
        synth.write_fasta(fa, chroms)
        refpipe.kmc_count(root, names[g], str(fa), g, k, threads=cores, timings=timings)
        os.remove(root / "kmc" / f"{names[g]}.count.kmc_suf")        # only the .onehot databases feed `complex`
        if g in wl["anchors"]:
            # the anchor as `per_anchor` FASTA files holding chromosome pieces that overlap by k-1 bases: the same
            # k-mer positions, every one exactly once, spread evenly over the threads run_anchor can use
            total = sum(s.size - k + 1 for _, s in chroms)
            target = (total + per_anchor - 1) // per_anchor
            files, j, fill = [[] for _ in range(per_anchor)], 0, 0
            for cname, s in chroms:
                nk, o = s.size - k + 1, 0
                while o < nk:
                    if target - fill < 1000 and j < per_anchor - 1:
                        j, fill = j + 1, 0
                    m = min(target - fill, nk - o) if j < per_anchor - 1 else nk - o
                    if nk - o - m < 1000:
                        m = nk - o                     # a tiny tail joins this piece (a record needs >= 100 k-mers)
                    files[j].append((f"{cname}_{o}", s[o:o + m + k - 1]))
                    positions += m
                    fill += m
                    o += m
            for j, recs in enumerate(files):
                if recs:
                    p = root / "fa" / f"a{g}_{j}.fa"
                    synth.write_fasta(p, recs)
                    pieces.append((f"a{g}x{j}", str(p)))
        os.remove(fa)
        log(f"reference setup: genome {g + 1}/{n} counted")
    del anc
    ndb = refpipe.write_opdefs(root, names)
    refpipe.kmc_bitvec(root, ndb, timings)
    for g in range(n):
        for ext in ("kmc_pre", "kmc_suf"):
            os.remove(root / "kmc" / f"{names[g]}.onehot.{ext}")
    meta_p.write_text(json.dumps({"pieces": pieces, "positions": positions, "timings": timings}))
    return root, pieces, positions, timings


def run_reference(steps, warmup, wl, n, cores):
    """`run_anchor N . a0 a0.fa a1 a1.fa ...` (cpp/Snakefile:55) over the whole anchor genome(s) of the workload
    against the full-size bitvec databases; OMP_NUM_THREADS = host cores, one piece of the anchor per thread."""
    from oracle import refpipe
    if not refpipe.have_ref():
        return {"impl": "reference", "unavailable": "oracle/_ref binaries missing (build with make -C oracle ref)"}
    root, pieces, positions, timings = reference_setup(wl, n, cores)
    times = []
    for it in range(warmup + steps):
        for name, _ in pieces:
            shutil.rmtree(root / "anchor" / name, ignore_errors=True)
        t = {}
        refpipe.run_anchor(root, n, pieces, threads=cores, timings=t)
        if it >= warmup:
            times.append(t["run_anchor_s"])
    ms = 1e3 * sum(times) / len(times)
    sample = (f"the whole workload: {n} genomes x {wl['length'] / 1e6:g} Mbp (seed {wl['seed']}), k={wl['k']}, "
              f"{len(wl['anchors'])} anchor genome(s) = {positions} positions per step, cut into {len(pieces)} FASTA files "
              f"(pieces overlap by k-1 bases: every k-mer position once) because run_anchor parallelises over anchors only "
              f"(cpp/anchor.cpp:217); every step loads the bitvec databases, as every run_anchor invocation does; KMC database build "
              f"untimed and cached (kmc {timings.get('kmc_count_s', 0):.0f}s, set_counts {timings.get('set_counts_s', 0):.0f}s, "
              f"complex {timings.get('kmc_bitvec_s', 0):.0f}s)")
    return {"value": positions * n / (ms / 1e3), "ms_per_step": ms, "positions": positions, "sample": sample,
            "cores": min(cores, len(pieces)), "cores_available": cores, "kind": "reference", "steps_s": times}


# ------------------------------------------------------------------------ whole `panagram index` (time iii)
INDEX_WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "configs0": dict(n=2, length=5_000_000, k=21, seed=20260000, anchors=None,
                     name="configs[0]: 2 synthetic 5 Mbp genomes, k=21, both anchors"),
    "configs1_sample": dict(n=8, length=6_000_000, k=21, seed=20260001, anchors=["g0"],
                            name="configs[1] sample: 8 synthetic 6 Mbp genomes, k=21, 1 anchor"),
    "configs1": dict(n=8, length=135_000_000, k=21, seed=20260001, anchors=["g0"],
                     name="configs[1]: 8 synthetic 135 Mbp genomes, k=21, 1 anchor"),
}


def run_index_e2e(args, cores):
    """FASTA files on disk -> finished anchor/<name>/ directories, ours (`python -m panagram_b200 index`: parse,
    k-mer tables built on the GPU from the sequences, anchor, BGZF on the GPU, files written) and, as the
    cpu_baseline leg, the reference's pipeline on the same files (cpp/Snakefile restated by oracle/refpipe.py over
    the unmodified binaries: kmc + set_counts per genome, kmc_tools complex, run_anchor), outputs compared."""
    from panagram_b200 import layout, synth
    from panagram_b200.index import Index, IndexConfig
    wl = INDEX_WORKLOADS[args.index_e2e]
    tmp = Path(tempfile.mkdtemp(prefix="pk_idx_"))
    unit = "anchored k-mers/s"
    res = {"metric": "anchored k-mers/sec (positions x genomes), whole `panagram index` from FASTA files to anchor directories",
           "unit": unit, "n_gpus": 1, "higher_is_better": True, "data": "synthetic", "dtype": "u64",
           "config": {"workload": wl["name"], "k": wl["k"], "host_cores": cores}}
    try:
        samples = synth.make_pangenome(tmp / "fa", wl["n"], wl["length"], wl["seed"])
        res["config"]["fasta_bytes"] = sum(os.path.getsize(p) for _, p in samples)
        anchors = wl["anchors"] or [n for n, _ in samples]
        tsv = tmp / "samples.tsv"
        tsv.write_text("name\tfasta\n" + "".join(f"{n}\t{p}\n" for n, p in samples))
        log = []
        idx = Index(tsv, tmp / "ours", IndexConfig(k=wl["k"], cores=min(cores, 32), anchor_genomes=list(anchors)))
        t0 = time.perf_counter()
        out = idx.run(log=log.append)
        ours_s = time.perf_counter() - t0
        positions = sum(o["positions"] for o in out.values())
        res.update(value=positions * wl["n"] / ours_s, total_s=round(ours_s, 3), positions=positions, log=log)
        if not args.no_cpu_baseline:
            from oracle import refpipe
            if refpipe.have_ref():
                t0 = time.perf_counter()
                tm = refpipe.build_index(tmp / "ref", samples, wl["k"], anchors=list(anchors), threads=cores)
                ref_s = time.perf_counter() - t0
                res["cpu_baseline"] = {"value": positions * wl["n"] / ref_s, "unit": unit, "kind": "reference", "cores": cores,
                                       "total_s": round(ref_s, 3), "stages_s": {k: round(v, 3) for k, v in tm.items()},
                                       "sample": f"the whole workload; kmc/kmc_tools -t{cores}, run_anchor one thread per anchor "
                                                 f"(cpp/anchor.cpp:217)"}
                bad = []
                for a in anchors:
                    want = refpipe.read_anchor_dir(tmp / "ref" / "anchor" / a)
                    d = tmp / "ours" / "anchor" / a
                    got = {"chrs.tsv": (d / "chrs.tsv").read_text(), "bitsum.bins.tsv": (d / "bitsum.bins.tsv").read_text(),
                           "bitmap.1": layout.read_bgzf(d / "bitmap.1.gz"), "bitmap.100": layout.read_bgzf(d / "bitmap.100.gz")}
                    bad += [f"{a}/{key}" for key in got if got[key] != want[key]]
                    res.setdefault("compared_bytes", {})[a] = {key: len(got[key]) for key in got}
                res["outputs_identical_to_reference"] = not bad
                if bad:
                    res["mismatch"] = bad
        print(json.dumps(res))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------ live DRAM traffic of the probe kernel
def ncu_traffic(argv_child: list[str], kernel_regex: str, skip: int, timeout_s: int = 600):
    """dram__bytes_read/write of ONE launch of the probe kernel, measured now: an ncu child runs this command in
    --ncu-child mode (same workload, tables, knobs; one warm-up step + one profiled step). None + the reason when ncu
    cannot run here."""
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if Path("/usr/local/cuda/bin/ncu").exists() else None)
    if ncu is None:
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", f"regex:{kernel_regex}", "-s", str(skip), "-c", "1", "--csv",
           sys.executable, str(ROOT / "bench.py")] + argv_child + ["--ncu-child"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    except Exception as ex:       # noqa: BLE001
        return None, f"ncu child failed: {ex}"
    rows = [l for l in r.stdout.splitlines() if l.startswith('"')]
    if r.returncode != 0 or len(rows) < 2:
        return None, f"ncu child rc={r.returncode}: {(r.stderr or r.stdout)[-300:]}"
    out = {}
    rd = csv.DictReader(io.StringIO("\n".join(rows)))
    for row in rd:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
                "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        out[row["Metric Name"]] = v * mult
        out["kernel"] = row.get("Kernel Name", "")
    if "dram__bytes_read.sum" not in out:
        return None, "ncu output without dram__bytes"
    return out, "ncu child of this command, one launch, --clock-control none"


# ------------------------------------------------------------------------ our arm
def choose_grid(wl, world, genome_ranks, n_total, hbm_gb=170.0):
    """(Rg, Rp). Weak workloads: one genome group. Strong workloads: as many replica groups as there are anchors to
    hand out, provided a rank's share of the tables fits."""
    if genome_ranks:
        assert world % genome_ranks == 0, "--genome-ranks must divide --gpus"
        return genome_ranks, world // genome_ranks
    if "n_total" not in wl:
        return world, 1
    rp = min(world, len(wl["anchors"]))
    while world % rp:
        rp -= 1
    rg = world // rp
    while rg < world and (n_total / rg) * wl["length"] * TABLE_BYTES_PER_GENOME_BASE / 1e9 > 0.75 * hbm_gb:
        rg *= 2
    return rg, world // rg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs1", choices=sorted(WORKLOADS))
    ap.add_argument("--load-factor", type=float, default=0.5)
    ap.add_argument("--probe-mode", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="do not spawn the ncu child that measures roofline.traffic")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ncu-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--index-e2e", default="", choices=[""] + sorted(INDEX_WORKLOADS),
                    help="instead of the anchoring step: time the whole `panagram index` run from FASTA files on disk")
    ap.add_argument("--group-tables", type=int, default=1, help="0: per-genome tables only (one probe per genome and position)")
    ap.add_argument("--group-only", type=int, default=1, help="free the per-genome tables once their group table is built")
    ap.add_argument("--e2e-batches", type=int, default=0, help="override the engine's batches per pk_anchor_genome call (0 = default)")
    ap.add_argument("--tune", default="", help="engine knobs for experiments, e.g. e2e_front_small=1,k3_lean=0")
    ap.add_argument("--genome-ranks", type=int, default=0, help="ranks per genome group (0 = auto)")
    ap.add_argument("--exchange", default="slice", choices=["slice", "nccl"],
                    help="Rg>1: position-split peer-memory exchange (every rank assembles its slice of the rows) or "
                         "NCCL all-gather + interleave (every rank assembles every row)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    unit = "anchored k-mers/s"
    metric = "anchored k-mers/sec (positions x genomes) building pan-kmer bitmap"
    strong = "n_total" in wl
    k = wl["k"]

    if args.index_e2e:
        if rank == 0:
            run_index_e2e(args, cores)
        return
    if args.impl == "reference":
        if rank != 0:
            return
        n = n_genomes_of(wl, args.gpus)
        rwl, full = reference_workload(wl, n)
        r = run_reference(args.steps, args.warmup, rwl, n, cores)
        if "unavailable" in r:
            print(json.dumps(r)); return
        cfg = workload_config(rwl, args, args.gpus)
        if not full:
            r["sample"] = f"genomes shortened from {wl['length'] / 1e6:g} to {rwl['length'] / 1e6:g} Mbp (bounded KMC build); " + r["sample"]
        cfg["positions_per_step"] = r["positions"]
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference",
                                 "sample": r["sample"], "cores_available": r["cores_available"]},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line)); return

    import torch
    import torch.distributed as dist
    from panagram_b200 import sharded, synth
    from panagram_b200.engine import pinned_empty

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_total = n_genomes_of(wl, world)
    rg, rp = choose_grid(wl, world, args.genome_ranks, n_total)
    hbm_peak, peak_src = peaks()

    # ---- setup (untimed): synthetic genomes -> tables of this rank's shard
    t0 = time.perf_counter()
    sh = sharded.ShardedAnchorer(k, n_total, rank, world, local_rank, genome_ranks=rg, load_factor=args.load_factor,
                                 probe_mode=args.probe_mode)
    eng = sh.engine
    g_begin, g_end = sh.begin, sh.end
    npg = g_end - g_begin
    if not args.group_tables:
        eng.tune(group_tables=0)
    elif args.group_only:
        eng.tune(group_only=1)
    anc = synth.ancestor_codes(wl["length"], wl["seed"])
    my_anchors = [a for j, a in enumerate(wl["anchors"]) if j % rp == sh.pi]
    anchor_chroms = {}
    for g in range(g_begin, g_end):
        chroms = make_genome(anc, g, wl["seed"], mu_period_of(wl))
        if g in my_anchors:
            anchor_chroms[g] = chroms
        eng.reserve(g, sum(c.size for c in chroms))
        for c in chroms:
            eng.add_sequence(g, c)
        if args.group_tables and args.group_only and (g - g_begin) % 8 == 7:
            eng.seal_group((g - g_begin) // 8)       # per-genome tables of the finished group are freed here
    for a in my_anchors:
        if a not in anchor_chroms:
            anchor_chroms[a] = make_genome(anc, a, wl["seed"], mu_period_of(wl))
    del anc
    eng.finalize()
    if args.e2e_batches:
        eng.tune(e2e_batches=args.e2e_batches)
    if args.tune:
        eng.tune(**{kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.tune.split(",") if kv})
    tstats = [eng.table_stats(g) for g in range(g_begin, g_end)]
    gstats = [eng.group_stats(u) for u in range((npg + 7) // 8)]
    setup_s = time.perf_counter() - t0

    # ---- device-resident inputs: every anchor of this rank's group packed in HBM
    rb_local, rb_full = eng.row_bytes, (n_total + 7) // 8
    stream = sh.stream
    st = stream.cuda_stream
    packed = []
    positions_mine = 0
    positions_all = len(wl["anchors"]) * (wl["length"] - 5 * (k - 1))       # every genome: 5 chromosomes, len - k + 1 positions each
    with torch.cuda.stream(stream):
        for a in my_anchors:
            chroms = anchor_chroms[a]
            lens = [c.size for c in chroms]
            # one concatenated sequence ('N' between chromosomes): cat row of chromosome c's k-mer p = off[c] + p
            cat = np.full(sum(lens) + len(lens) - 1, ord("N"), dtype=np.uint8)
            off, o = [], 0
            for c in chroms:
                cat[o:o + c.size] = c
                off.append(o)
                o += c.size + 1
            ltot = cat.size
            npos = ltot - k + 1
            nks = [l - k + 1 for l in lens]
            d_ascii = torch.from_numpy(cat).to(dev)
            nw = eng.packed_words(ltot)
            d_words = torch.empty(nw, dtype=torch.int64, device=dev)
            d_mask = torch.empty(nw, dtype=torch.int32, device=dev)
            eng.pack_device(d_ascii.data_ptr(), ltot, d_words.data_ptr(), d_mask.data_ptr(), st)
            sb = sharded.slice_bounds(sum(nks), rb_full, rg)
            s0, s1 = sb[sh.gi], sb[sh.gi + 1]
            packed.append({"a": a, "words": d_words, "mask": d_mask, "npos": npos, "nks": nks, "off": off, "ltot": ltot,
                           "segs": sharded.stream_segments(off, nks, s0, s1), "slice": (s0, s1), "chroms": chroms})
            positions_mine += sum(nks)
            del d_ascii
        max_npos = max(p["npos"] for p in packed) if packed else 1
        if rg > 1:
            sh._ensure_planes(max_npos)
            max_slice = max(p["slice"][1] - p["slice"][0] for p in packed)
            d_slice = torch.empty((max(max_slice, 1), rb_full), dtype=torch.uint8, device=dev)
            if args.exchange == "nccl":
                d_local = torch.empty((max_npos, rb_local), dtype=torch.uint8, device=dev)
                d_planes = torch.empty((rg, max_npos, rb_local), dtype=torch.uint8, device=dev)
                d_rows = torch.empty((max_npos, rg * rb_local), dtype=torch.uint8, device=dev)
        else:
            d_local = torch.empty((max_npos, rb_local), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    tev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def device_step(timing=None):
        for p in packed:
            if rg == 1:
                eng.probe_device(p["words"].data_ptr(), p["mask"].data_ptr(), 0, p["npos"], d_local.data_ptr(), rb_local, 0, st)
            elif args.exchange == "nccl":
                assert p["npos"] == max_npos
                sh.probe_allgather(p["words"].data_ptr(), p["mask"].data_ptr(), p["npos"], d_local, d_planes, d_rows)
            else:
                # the gather of this anchor runs on a side stream under the probe of the next one (the planes alternate)
                sh.probe_exchange(p["words"].data_ptr(), p["mask"].data_ptr(), p["npos"], p["segs"], d_slice, timing=timing,
                                  overlap=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.ncu_child:
        with torch.cuda.stream(stream):
            for _ in range(2):
                device_step()
        torch.cuda.synchronize()
        return

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            device_step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        ev[0].record(stream)
        for i in range(args.steps):
            device_step(timing=tev if i == args.steps - 1 else None)
            if i == args.steps - 1:
                sh.finish_exchange()             # the last gather belongs to the timed region
            ev[i + 1].record(stream)
        barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    ks = eng.stats()            # kernels of the last probe launch, CUDA events on the launching stream
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "genomes": [g_begin, g_end], "group_table_bytes": sum((g or {"bytes": 0})["bytes"] for g in gstats),
                "kernels_ms": {x: round(ks[x], 4) for x in ("k_partition_ms", "k_fine_ms", "k_probe_ms", "k_spill_ms", "k_unpermute_ms")},
                "step_ms_median": statistics.median(step_ms)}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    exch = None
    if rg > 1 and args.exchange == "slice" and packed:
        p = packed[-1]
        gather_ms = tev[1].elapsed_time(tev[2])
        remote = (p["slice"][1] - p["slice"][0]) * rb_local * (rg - 1)          # bytes this rank read from its peers
        exch = {"what": "last anchor of the last step on rank 0: probe into the plane | barrier (1-element all-reduce) | "
                        "gather_slice_kernel reading the rank's slice of all planes over NVLink",
                "probe_plus_barrier_ms": tev[0].elapsed_time(tev[1]), "gather_ms": gather_ms,
                "nvlink_bytes_read": int(remote), "nvlink_gbs": remote / (gather_ms / 1e3) / 1e9 if gather_ms > 0 else None,
                "nvlink_peak_gbs": NVLINK_PEAK_GBS,
                "nvlink_frac": remote / (gather_ms / 1e3) / 1e9 / NVLINK_PEAK_GBS if gather_ms > 0 else None,
                "rows_written_bytes": int((p["slice"][1] - p["slice"][0]) * rb_full)}
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = positions_all * n_total / (ms_per_step / 1e3)

    # ---- e2e: the public call with pinned host buffers, H2D + D2H inside the timed region
    e2e = e2e_files = None
    if not args.no_e2e:
        h_sets = []
        for p in packed:
            hs = []
            for c in p["chroms"]:
                h = pinned_empty(c.size)
                h[:] = c
                hs.append(h)
            h_sets.append(hs)
        if world == 1 and not strong:
            h_chroms = h_sets[0]
            res = eng.anchor_genome(h_chroms, pinned=True)          # allocates the pinned output buffers once
            for _ in range(max(1, args.warmup - 1)):
                res = eng.anchor_genome(h_chroms, out=res)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.steps):
                t1 = time.perf_counter()
                res = eng.anchor_genome(h_chroms, out=res)
                ts.append((time.perf_counter() - t1) * 1e3)
            e2e_ms = statistics.median(ts)          # per-call wall times; the median keeps one host hiccup out of a 5-step mean
            es = eng.stats()
            h2d = sum(c.size for c in h_chroms)
            d2h = sum(r["bitmap1"].nbytes + r["low"].nbytes + r["bin_hist"].nbytes for r in res["chroms"]) + 8 * npg
            e2e = {"value": positions_all * n_total / (e2e_ms / 1e3), "unit": unit, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step_mean": sum(ts) / len(ts),
                   "call": "Engine.anchor_genome -> pk_anchor_genome",
                   # CUDA events of the last call on the engine's streams
                   "device_timeline_ms": {"first_batch_h2d_pack_partition": es["h2d_ms"], "first_batch_probe": es["probe_ms"],
                                          "unpermute_reduce_and_second_batch": es["reduce_ms"], "d2h_tail": es["d2h_ms"],
                                          "total": es["total_ms"]}}
            e2e_launches = es["kernel_launches"]
            rz = eng.anchor_genome_bgzf(h_chroms)
            rz = eng.anchor_genome_bgzf(h_chroms, out=rz)
            ts = []
            for _ in range(args.steps):
                t1 = time.perf_counter()
                rz = eng.anchor_genome_bgzf(h_chroms, out=rz)
                ts.append((time.perf_counter() - t1) * 1e3)
            e2e_files = {"what": "pk_anchor_genome_bgzf: ASCII in (pinned host) -> bitmap.1.gz/.gzi + bitmap.100.gz/.gzi file "
                                 "images (BGZF deflated on the GPU) + histograms + column sums out",
                         "ms_per_step": statistics.median(ts), "value": positions_all * n_total / (statistics.median(ts) / 1e3), "unit": unit,
                         "h2d_bytes_per_step": int(h2d),
                         "d2h_bytes_per_step": int(rz["gz"].size + rz["gzi"].size + rz["gz_low"].size + rz["gzi_low"].size),
                         "raw_bitmap_bytes": int(positions_all * rb_local), "gz_bytes": int(rz["gz"].size)}
        else:
            # every rank: H2D of its group's anchors, pack, probe its shard, exchange, reduce its slice, rows (or BGZF
            # members) of its slice to host memory
            def e2e_pass(bgzf):
                nb = 0
                for hs in h_sets:
                    r = sh.anchor_genome(hs, bgzf=bgzf, rows_to_host=not bgzf)
                    nb += (r["gz"].size + r["gzi"].size) if bgzf else r["rows_host"].nbytes
                    nb += r["low"].nbytes + sum(h.nbytes for h in r["hist"] if h is not None) + r["col_sums"].nbytes
                return nb
            out = {}
            for bgzf in (False, True):
                e2e_pass(bgzf); barrier()
                t1 = time.perf_counter()
                d2h = 0
                for _ in range(args.steps):
                    d2h = e2e_pass(bgzf)
                barrier()
                ms = (time.perf_counter() - t1) * 1e3 / args.steps
                # host -> device: every rank copies its 1/Rg share of its group's anchors (the rest arrives over NVLink)
                t = torch.tensor([ms, float(d2h), float(sum(c.size for hs in h_sets for c in hs)) / rg], device=dev, dtype=torch.float64)
                tmax = t.clone()
                if world > 1:
                    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                out[bgzf] = {"value": positions_all * n_total / (float(tmax[0]) / 1e3), "unit": unit, "ms_per_step": float(tmax[0]),
                             "h2d_bytes_per_step": int(t[2]), "d2h_bytes_per_step": int(t[1]),
                             "call": "ShardedAnchorer.anchor_genome (pk_anchor_genome_plane + exchange + reduce" +
                                     (" + BGZF on the GPU)" if bgzf else ", rows of the slice to host)"),
                             "bytes": "summed over all ranks", "last_call_ms": dict(sh.last)}
            e2e, e2e_files = out[False], out[True]
            e2e_launches = ks["kernel_launches"] + 4
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        # roofline of the dominant kernel (K3, the hash probe), per launch = one anchor against this rank's tables
        pos1 = sum(packed[-1]["nks"]) if packed else 0
        n_group_tabs = (npg + 7) // 8
        win = int(ks.get("k_probe_window", 0))
        alg_design = pos1 * ((n_group_tabs if win in (2, 3, 4, 5) else npg) * 32 + 0.375 + rb_local * 1.01)
        alg_survey = pos1 * (npg * 32 + 0.375 + rb_local * 1.01)
        k3_ms = ks["k_probe_ms"] if ks["k_probe_ms"] > 0 else None
        stage1_ms = sum(ks[x] for x in ("k_partition_ms", "k_fine_ms", "k_probe_ms", "k_spill_ms", "k_unpermute_ms"))
        kernel_name = {2: "probe_win_kernel<group tables>", 1: "probe_win_kernel", 3: "items_group_kernel",
                       4: "probe_g32c_kernel (lean window kernel, one 32-bit-slot group table)",
                       5: "probe_g32l2_kernel (coarse regions through L2, no K2)"}.get(win, "probe_part_kernel") \
            if k3_ms else "probe_kernel (direct)"
        traffic, traffic_src, ncu_ms = None, "not measured (--no-ncu)", None
        if world == 1 and k3_ms and not args.no_ncu:
            child = ["--workload", args.workload, "--load-factor", str(args.load_factor), "--probe-mode", args.probe_mode,
                     "--group-tables", str(args.group_tables), "--group-only", str(args.group_only), "--no-cpu-baseline", "--no-ncu", "--no-e2e"] + \
                    (["--tune", args.tune] if args.tune else [])
            # the child builds the same tables beside ours: only when this GPU has the room. It runs two steps; the
            # capture is the probe kernel of the second (the spill drain reuses probe_part / items_group: skipped)
            free_b, total_b = torch.cuda.mem_get_info()
            if free_b > (total_b - free_b) + (16 << 30):
                rx, skip = {2: ("probe_win_kernel", 1), 1: ("probe_win_kernel", 1), 3: ("items_group_kernel", 2), 4: ("probe_g32c_kernel", 1), 5: ("probe_g32l2_kernel", 1)}.get(win, ("probe_part_kernel", 2))
                m, traffic_src = ncu_traffic(child, rx, skip)
            else:
                m, traffic_src = None, "not measured: no room for the ncu child's tables beside ours"
            if m:
                traffic = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
                ncu_ms = m.get("gpu__time_duration.sum", 0) / 1e6
        t_k3 = (k3_ms or ms_per_step) / 1e3
        roof = {"bound": "hbm", "kernel": kernel_name, "unit": "GB/s", "peak": hbm_peak, "peak_source": peak_src,
                "kernel_ms": k3_ms, "launch": f"one anchor ({pos1} positions) against {npg} genomes' tables on one GPU",
                "algorithmic": "design: one 32 B sector per position and 8-genome group table + 0.375 B/position of packed "
                               "sequence + 1.01 x row bytes written" if win in (2, 3, 4, 5) else
                               "SURVEY §8d: one 32 B sector per (position, genome) + 0.375 B/position + 1.01 x row bytes",
                "algorithmic_bytes_per_launch": alg_design, "achieved": alg_design / t_k3 / 1e9,
                "frac": alg_design / t_k3 / 1e9 / hbm_peak,
                "survey_accounting": {"what": "SURVEY §8d charges 32 B per (position, genome); a group table answers 8 genomes per "
                                              "sector, so this fraction exceeds 1 by design and is not a roofline fraction",
                                      "bytes_per_launch": alg_survey, "achieved": alg_survey / t_k3 / 1e9,
                                      "frac": alg_survey / t_k3 / 1e9 / hbm_peak},
                "traffic": traffic, "traffic_source": traffic_src,
                "physical": ({"dram_bytes_per_launch": traffic, "achieved": traffic / t_k3 / 1e9, "frac": traffic / t_k3 / 1e9 / hbm_peak,
                              "ncu_kernel_ms": ncu_ms, "waste_vs_algorithmic": traffic / alg_design} if traffic else None),
                "stage": {"what": "all kernels of one probe launch (partition_seq + partition_fine + probe + spill + unpermute), CUDA events",
                          "ms": stage1_ms, "achieved": alg_design / (stage1_ms / 1e3) / 1e9 if stage1_ms > 0 else None,
                          "frac": alg_design / (stage1_ms / 1e3) / 1e9 / hbm_peak if stage1_ms > 0 else None,
                          "frac_survey_accounting": alg_survey / (stage1_ms / 1e3) / 1e9 / hbm_peak if stage1_ms > 0 else None,
                          "kernels_ms": {"partition_seq": ks["k_partition_ms"], "partition_fine": ks["k_fine_ms"],
                                         "probe": ks["k_probe_ms"], "spill": ks["k_spill_ms"], "unpermute": ks["k_unpermute_ms"]}}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rwl, full = reference_workload(wl, n_total)
            r = run_reference(1, 0, rwl, n_total, cores)
            if "unavailable" not in r and not full:
                r["sample"] = f"genomes shortened from {wl['length'] / 1e6:g} to {rwl['length'] / 1e6:g} Mbp (bounded KMC build); " + r["sample"]
            if "unavailable" not in r:
                cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                       "ms": r["ms_per_step"]}
            else:
                cpu = {"value": None, "unit": unit, "cores": 0, "kind": "reference", "sample": r["unavailable"]}
        per_probe = 3 + (1 if ks["k_fine_ms"] > 0 and win != 5 else 0) + (1 if ks["k_unpermute_ms"] > 0 else 0)      # K1, [K2,] K3, spill drain, [K4]
        launches = (per_probe + (1 if rg > 1 else 0)) * len(packed) * args.steps
        cfg = workload_config(wl, args, world)
        cfg["positions_per_step"] = positions_all
        run = {}
        run.update({"genomes_per_gpu": npg, "load_factor": args.load_factor, "probe_mode": args.probe_mode,
                    "grid": {"genome_ranks": rg, "groups": rp, "anchors_on_rank0": len(packed)},
                    "l2": "no flush needed: every probe launch streams its tables (%.1f GB of group tables on rank 0) and %.1f GB of "
                          "partition scratch, far more than the 126 MB L2"
                          % (sum((g or {"bytes": 0})["bytes"] for g in gstats) / 1e9, (packed[-1]["npos"] if packed else 0) * 24 / 1e9),
                    "parallelism": ("1 GPU" if world == 1 else
                                    f"{rp} genome group(s) x {rg} rank(s); " +
                                    ("no exchange (replica groups take the anchors round-robin)" if rg == 1 else
                                     ("position-split peer-memory exchange (gather_slice_kernel over NVLink on a side stream under the next probe, one barrier per anchor)"
                                      if args.exchange == "slice" else "NCCL all-gather + interleave"))),
                    "setup_s": round(setup_s, 1)})
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": cfg, "run": run, "e2e": e2e, "e2e_files": e2e_files, "gpu_launches": int(launches), "roofline": roof,
                "exchange": exch, "per_rank": per_rank, "cpu_baseline": cpu, "clocks": clocks, "step_ms": step_ms,
                "value_excludes": "pack (ASCII -> 2-bit), reduce and low-res kernels (~0.4 ms per 135 M positions): SURVEY §8d (i) "
                                  "times the probe stage; they are inside `e2e`",
                "tables": {"keys": [t["n_keys"] for t in tstats], "per_genome_table_bytes": sum(t["bytes"] for t in tstats),
                           "overflow_frac": sum(t["n_overflow"] for t in tstats) / max(1, sum(t["n_keys"] for t in tstats)),
                           "group_tables": gstats}}
        print(json.dumps(line))
    sh.close_p2p()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
