#!/usr/bin/env python3
"""bench.py — anchored k-mers/sec (positions x genomes) building the pan-kmer bitmap.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                  [--group-tables 0|1] [--e2e-batches B] [--exchange p2p|p2p-serial|nccl] [--index-e2e W]

One "step" = one pass of the hot path over one anchor genome: every k-mer position of the
anchor probed against every genome's k-mer set and the N-bit rows written.

Workloads (BASELINE.json `configs`; synthetic genomes per SURVEY.md §8d, panagram_b200/synth.py):
  configs1   8 x 135 Mbp, k=21, 1 anchor — the configuration the metric is quoted on (default)
  configs2   32 x 150 Mbp, k=21, 1 anchor
  configs3s  one GPU's shard of configs[3]: 8 x 150 Mbp, k=31
  small      8 x 8 Mbp, k=21 (plumbing check)
At N GPUs the genomes are sharded by genome (weak scaling: 8 genomes' tables per GPU, 8N
genomes in total, every rank probes all anchor positions against its shard, one fused peer-memory
gather+interleave kernel — or an NCCL all-gather + interleave kernel — assembles the N-bit rows).

`value`     device-resident: packed anchor already in HBM, CUDA events around the probe stage
            (partition + probe kernels [+ exchange at N>1]).
`e2e`       the public call a user makes (Engine.anchor_genome -> pk_anchor_genome) with pinned
            HOST buffers: ASCII in, bitmap rows / low-res rows / histograms / column sums out.
`e2e_files` the same with the two bitmaps delivered as BGZF file images deflated on the GPU.
`roofline`  the probe kernel: SURVEY §8d's algorithmic bytes / its CUDA-event duration (`frac`), and the DRAM
            bytes ncu counted for the same launch / that duration (`physical`).
`--impl reference`  the reference's own CPU implementation (oracle/_ref/run_anchor, built from
            the unmodified cpp/anchor.cpp + KMC API) on a bounded sample of the same workload.
`--index-e2e W`     instead: the whole `panagram index` run, FASTA files on disk -> anchor directories,
            next to the reference's pipeline on the same files, outputs compared.
"""
from __future__ import annotations

import argparse
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
    "configs1": dict(n_per_gpu=8, length=135_000_000, k=21, seed=20260001,
                     name="configs[1]: 8 synthetic 135 Mbp genomes, k=21, 1 anchor"),
    "configs2": dict(n_per_gpu=32, length=150_000_000, k=21, seed=20260002,
                     name="configs[2]: 32 synthetic 150 Mbp genomes, k=21, 1 anchor"),
    "configs3s": dict(n_per_gpu=8, length=150_000_000, k=31, seed=20260003,
                      name="configs[3], one GPU's shard: 8 of 64 synthetic 150 Mbp genomes, k=31 (64-bit slots), 1 of 4 anchors"),
    "small": dict(n_per_gpu=8, length=8_000_000, k=21, seed=20260009,
                  name="small: 8 synthetic 8 Mbp genomes, k=21, 1 anchor"),
}
CPU_SAMPLE_LEN = 6_000_000      # per-genome length of the bounded CPU sample


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def make_genome(anc, g, seed):
    from panagram_b200 import synth
    return [s for _, s in synth.genome_chroms(anc, g, seed)]


# ------------------------------------------------------------------------ reference arm
def run_reference(args, wl, cores):
    """The reference's CPU path on a bounded sample of the workload: KMC databases are built once
    (untimed, like our table build); each step times `run_anchor N . g0 g0.fa` (cpp/Snakefile:55)."""
    from oracle import refpipe
    from panagram_b200 import synth
    n = wl["n_per_gpu"] * max(1, getattr(args, "gpus", 1))     # the same genome count as our arm at this N (weak scaling)
    if not refpipe.have_ref():
        return {"impl": "reference", "unavailable": "oracle/_ref binaries missing (build with make -C oracle ref)"}
    length = min(wl["length"], CPU_SAMPLE_LEN)
    tmp = Path(tempfile.mkdtemp(prefix="pk_ref_"))
    try:
        samples = synth.make_pangenome(tmp / "fa", n, length, wl["seed"])
        timings = {}
        names = [s[0] for s in samples]
        for i, (name, fa) in enumerate(samples):
            refpipe.kmc_count(tmp / "idx", name, fa, i, wl["k"], threads=cores, timings=timings)
        ndb = refpipe.write_opdefs(tmp / "idx", names)
        refpipe.kmc_bitvec(tmp / "idx", ndb, timings)
        positions = sum(len(s) - wl["k"] + 1 for _, s in
                        __import__("oracle.oracle", fromlist=["x"]).parse_fasta(samples[0][1]))
        times = []
        for it in range(args.warmup + args.steps):
            t = {}
            refpipe.run_anchor(tmp / "idx", n, [samples[0]], threads=cores, timings=t)
            if it >= args.warmup:
                times.append(t["run_anchor_s"])
        ms = 1e3 * sum(times) / len(times)
        value = positions * n / (ms / 1e3)
        sample = (f"{n} genomes x {length / 1e6:g} Mbp of the same generator (seed {wl['seed']}), k={wl['k']}, "
                  f"1 anchor = {positions} positions; KMC DB build untimed "
                  f"(kmc {timings.get('kmc_count_s', 0):.1f}s, set_counts {timings.get('set_counts_s', 0):.1f}s, "
                  f"complex {timings.get('kmc_bitvec_s', 0):.1f}s)")
        return {"value": value, "ms_per_step": ms, "positions": positions, "sample": sample,
                "cores": 1, "cores_available": cores, "kind": "reference"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------ whole `panagram index` (time iii)
INDEX_WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "configs0": dict(n=2, length=5_000_000, k=21, seed=20260000, anchors=None,
                     name="configs[0]: 2 synthetic 5 Mbp genomes, k=21, both anchors"),
    # the bounded sample the CPU arm uses for configs[1]
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
                res["outputs_identical_to_reference"] = not bad
                if bad:
                    res["mismatch"] = bad
        print(json.dumps(res))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------ our arm
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
    ap.add_argument("--index-e2e", default="", choices=[""] + sorted(INDEX_WORKLOADS),
                    help="instead of the anchoring step: time the whole `panagram index` run from FASTA files on disk")
    ap.add_argument("--group-tables", type=int, default=1, help="0: per-genome tables only (one probe per genome and position)")
    ap.add_argument("--e2e-batches", type=int, default=0, help="override the engine's batches per pk_anchor_genome call (0 = default)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "p2p-serial", "nccl"],
                    help="N>1: assemble rows with the fused peer-memory gather+interleave kernel (on a side stream under the "
                         "next probe, or serially on the probe's stream) or NCCL all-gather + interleave")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    unit = "anchored k-mers/s"
    metric = "anchored k-mers/sec (positions x genomes) building pan-kmer bitmap"

    if args.index_e2e:
        if rank == 0:
            run_index_e2e(args, cores)
        return
    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference(args, wl, cores)
        if "unavailable" in r:
            print(json.dumps(r)); return
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic",
                "config": {"workload": wl["name"] + (f" x{args.gpus} genome shards ({wl['n_per_gpu'] * args.gpus} genomes)" if args.gpus > 1 else ""),
                           "k": wl["k"], "n_genomes": wl["n_per_gpu"] * max(1, args.gpus), "sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference",
                                 "sample": r["sample"],
                                 "note": f"run_anchor parallelises over anchors only (cpp/anchor.cpp:217): 1 anchor "
                                         f"uses 1 of {cores} host cores"},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line)); return

    import torch
    import torch.distributed as dist
    from panagram_b200 import synth
    from panagram_b200.engine import Engine, pinned_empty

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    npg, k = wl["n_per_gpu"], wl["k"]
    n_total = npg * world
    g_begin, g_end = rank * npg, (rank + 1) * npg
    hbm_peak, peak_src = peaks()

    # ---- setup (untimed): synthetic genomes -> per-genome tables on this rank's shard
    t0 = time.perf_counter()
    anc = synth.ancestor_codes(wl["length"], wl["seed"])
    eng = Engine(k, n_total, g_begin, g_end, device=local_rank, load_factor=args.load_factor,
                 probe_mode=args.probe_mode)
    anchor_chroms = None
    for g in range(g_begin, g_end):
        chroms = make_genome(anc, g, wl["seed"])
        if g == 0:
            anchor_chroms = chroms
        eng.reserve(g, sum(c.size for c in chroms))
        for c in chroms:
            eng.add_sequence(g, c)
    if anchor_chroms is None:
        anchor_chroms = make_genome(anc, 0, wl["seed"])
    del anc
    if not args.group_tables:
        eng.tune(group_tables=0)
    eng.finalize()
    if args.e2e_batches:
        eng.tune(e2e_batches=args.e2e_batches)
    tstats = [eng.table_stats(g) for g in range(g_begin, g_end)]
    setup_s = time.perf_counter() - t0

    # anchor as one concatenated sequence ('N' between chromosomes), the layout pk_anchor_genome uses
    lens = [c.size for c in anchor_chroms]
    positions = sum(l - k + 1 for l in lens)
    cat = np.full(sum(lens) + len(lens) - 1, ord("N"), dtype=np.uint8)
    o = 0
    for c in anchor_chroms:
        cat[o:o + c.size] = c
        o += c.size + 1
    ltot = cat.size
    npos = ltot - k + 1
    rb_local = eng.row_bytes
    rb_full = (n_total + 7) // 8
    d_ascii = torch.from_numpy(cat).to(dev)
    nw = eng.packed_words(ltot)
    d_words = torch.empty(nw, dtype=torch.int64, device=dev)
    d_mask = torch.empty(nw, dtype=torch.int32, device=dev)
    # an explicit (non-default) torch stream: the library launches on it and the torch events below
    # are recorded on it (stream handle 0 would mean "the engine's own stream" to the library)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    assert st != 0
    eng.pack_device(d_ascii.data_ptr(), ltot, d_words.data_ptr(), d_mask.data_ptr(), st)
    d_local = torch.empty((npos, rb_local), dtype=torch.uint8, device=dev)
    if world > 1:
        d_planes = torch.empty((world, npos, rb_local), dtype=torch.uint8, device=dev)
        d_rows = torch.empty((npos, rb_full), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    p2p = None
    if world > 1 and args.exchange in ("p2p", "p2p-serial"):
        # fused exchange over peer memory: each rank's planes (two, ping-pong) are IPC-mapped into every peer;
        # stream-ordered barriers (1-element NCCL all-reduce) order the ranks, the data itself never goes
        # through NCCL. The gather+interleave of step i runs on a side stream under the probe of step i+1
        # (consecutive steps = consecutive anchor genomes).
        planes, peers = [], []
        for b in range(2):
            pl = eng.device_alloc(npos * rb_local)
            handles = [None] * world
            dist.all_gather_object(handles, eng.ipc_export(pl))
            planes.append(pl)
            peers.append([pl if r == rank else eng.ipc_open(handles[r]) for r in range(world)])
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        side = torch.cuda.Stream(device=dev)
        d_rows2 = [d_rows, torch.empty_like(d_rows)]
        p2p = {"i": 0, "probed": [torch.cuda.Event() for _ in range(2)], "gathered": [torch.cuda.Event() for _ in range(2)]}

    def device_step():
        if p2p:
            i = p2p["i"]; b = i & 1
            if i >= 2:
                tstream.wait_event(p2p["gathered"][b])      # own gather of step i-2 done (rows/plane b free)
            dist.all_reduce(flag)                            # ... on every rank: plane b may be overwritten
            eng.probe_device(d_words.data_ptr(), d_mask.data_ptr(), 0, npos, planes[b], rb_local, 0, st)
            dist.all_reduce(flag)                            # every rank's plane b is complete
            if args.exchange == "p2p-serial":
                eng.gather_interleave_device(peers[b], npos, rb_local, d_rows2[b].data_ptr(), rb_full, st)
                p2p["gathered"][b].record(tstream)
            else:
                p2p["probed"][b].record(tstream)
                side.wait_event(p2p["probed"][b])
                eng.gather_interleave_device(peers[b], npos, rb_local, d_rows2[b].data_ptr(), rb_full, side.cuda_stream)
                p2p["gathered"][b].record(side)
            p2p["i"] = i + 1
            return
        eng.probe_device(d_words.data_ptr(), d_mask.data_ptr(), 0, npos, d_local.data_ptr(), rb_local, 0, st)
        if world > 1:
            dist.all_gather_into_tensor(d_planes.view(-1), d_local.view(-1))
            eng.interleave_device(d_planes.data_ptr(), world, npos, rb_local, d_rows.data_ptr(), rb_full, st)

    def drain():
        if p2p:
            for ev_ in p2p["gathered"]:
                tstream.wait_event(ev_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kstats = []
    barrier()
    ev[0].record()
    for i in range(args.steps):
        device_step()
        if i == args.steps - 1:
            drain()                  # the last step's exchange is inside the timed region
        ev[i + 1].record()
    barrier()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    ks = eng.stats()            # kernels of the last launch, CUDA events on the launching stream
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = positions * n_total / (ms_per_step / 1e3)

    # ---- e2e: the public call with pinned host buffers, H2D + D2H inside the timed region
    h_chroms = []
    for c in anchor_chroms:
        h = pinned_empty(c.size)
        h[:] = c
        h_chroms.append(h)
    e2e_ms = None
    e2e_ms_mean = None
    e2e_files = None
    e2e_timeline = None
    if world == 1:
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
        e2e_ms_mean = sum(ts) / len(ts)
        e2e_stats = eng.stats()
        # CUDA events of the last call on the engine's streams: H2D + pack + K1 (overlapped) | K2 + K3 | K4 + reduce per
        # chromosome | tail of the D2H copies
        e2e_timeline = {"first_batch_h2d_pack_partition": e2e_stats["h2d_ms"], "first_batch_probe": e2e_stats["probe_ms"],
                        "unpermute_reduce_and_second_batch": e2e_stats["reduce_ms"], "d2h_tail": e2e_stats["d2h_ms"],
                        "total": e2e_stats["total_ms"]}
        h2d = sum(lens)
        d2h = sum(r["bitmap1"].nbytes + r["low"].nbytes + r["bin_hist"].nbytes for r in res["chroms"]) + 8 * npg
        e2e_launches = e2e_stats["kernel_launches"]
        # the same call delivering bitmap.1.gz/.gzi + bitmap.100.gz/.gzi as file images compressed on the GPU
        rz = eng.anchor_genome_bgzf(h_chroms)
        rz = eng.anchor_genome_bgzf(h_chroms, out=rz)
        ts = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            rz = eng.anchor_genome_bgzf(h_chroms, out=rz)
            ts.append((time.perf_counter() - t1) * 1e3)
        e2e_files = {"what": "pk_anchor_genome_bgzf: ASCII in (pinned host) -> bitmap.1.gz/.gzi + bitmap.100.gz/.gzi file "
                             "images (BGZF deflated on the GPU) + histograms + column sums out",
                     "ms_per_step": statistics.median(ts), "value": positions * n_total / (statistics.median(ts) / 1e3), "unit": unit,
                     "h2d_bytes_per_step": int(h2d),
                     "d2h_bytes_per_step": int(rz["gz"].size + rz["gzi"].size + rz["gz_low"].size + rz["gzi_low"].size),
                     "raw_bitmap_bytes": int(positions * rb_local), "gz_bytes": int(rz["gz"].size)}
    else:
        # rank r: H2D of the anchor, pack, probe its shard, all-gather, interleave; rank 0 reads the rows back
        h_cat = pinned_empty(ltot); h_cat[:] = cat
        t_cat = torch.from_numpy(h_cat)
        h_out = torch.empty((npos, rb_full), dtype=torch.uint8).pin_memory() if rank == 0 else None

        def e2e_step():
            d_ascii.copy_(t_cat, non_blocking=True)
            eng.pack_device(d_ascii.data_ptr(), ltot, d_words.data_ptr(), d_mask.data_ptr(), st)
            device_step()
            drain()
            if rank == 0:
                h_out.copy_(d_rows2[(p2p["i"] - 1) & 1] if p2p else d_rows, non_blocking=True)
        e2e_step(); barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t1) * 1e3 / args.steps
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        h2d, d2h = ltot, npos * rb_full
        e2e_launches = ks["kernel_launches"] + 2
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        # roofline of the dominant kernel (probe_part): algorithmic bytes per launch / its duration
        nbytes_local = rb_local
        alg_bytes = positions * (npg * 32 + 0.375 + nbytes_local * 1.01)
        k3_ms = ks["k_probe_ms"] if ks["k_probe_ms"] > 0 else ms_per_step
        stage_ms = statistics.median(step_ms)
        traffic = None
        try:
            key = args.workload if int(ks.get("k_probe_window", 0)) == 2 else args.workload + "_per_genome_tables"
            tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get(key)
            if tj and world == 1 and ks["k_probe_ms"] > 0:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": ({2: "probe_win_kernel<group tables>", 1: "probe_win_kernel"}.get(int(ks.get("k_probe_window", 0)), "probe_part_kernel"))
                if ks["k_probe_ms"] > 0 else "probe_kernel",
                "achieved": alg_bytes / (k3_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (k3_ms / 1e3) / 1e9 / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k3_ms,
                # the same kernel against the bytes it really moves (ncu dram__bytes of one launch / its live duration):
                # `frac` above charges 32 B per (position, genome) as SURVEY §8d defines; a group table answers 8
                # genomes per sector, so that figure exceeds 1 while this one says how close the kernel is to the pins
                "physical": ({"dram_bytes_per_launch": traffic, "achieved": traffic / (k3_ms / 1e3) / 1e9, "unit": "GB/s",
                              "frac": traffic / (k3_ms / 1e3) / 1e9 / hbm_peak} if traffic else None),
                "stage": {"what": "all kernels of the probe stage (partition_seq + partition_fine + probe_part + spill + unpermute)",
                          "ms": stage_ms, "achieved": alg_bytes / (stage_ms / 1e3) / 1e9,
                          "frac": alg_bytes / (stage_ms / 1e3) / 1e9 / hbm_peak,
                          "kernels_ms": {"partition_seq": ks["k_partition_ms"], "partition_fine": ks["k_fine_ms"],
                                         "probe_part": ks["k_probe_ms"], "spill": ks["k_spill_ms"],
                                         "unpermute": ks["k_unpermute_ms"]}}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            a2 = argparse.Namespace(steps=1, warmup=0, gpus=1)
            r = run_reference(a2, wl, cores)
            if "unavailable" not in r:
                cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference", "sample": r["sample"],
                       "ms": r["ms_per_step"]}
            else:
                cpu = {"value": None, "unit": unit, "cores": 0, "kind": "reference", "sample": r["unavailable"]}
        launches = (ks["kernel_launches"] + (1 if world > 1 else 0)) * args.steps
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": wl["name"] + (f" x{world} genome shards ({n_total} genomes)" if world > 1 else ""),
                           "k": k, "n_genomes": n_total, "genomes_per_gpu": npg, "positions": positions,
                           "load_factor": args.load_factor, "probe_mode": args.probe_mode,
                           "l2": "no flush needed: every step streams its tables (group tables %.1f GB/GPU, derived from %.1f GB of "
                                 "per-genome tables) and %.1f GB of partition scratch, far more than the 126 MB L2"
                                 % (sum((eng.group_stats(u) or {"bytes": 0})["bytes"] for u in range((npg + 7) // 8)) / 1e9,
                                    sum(t["bytes"] for t in tstats) / 1e9, positions * 24 / 1e9),
                           "parallelism": (f"genome-sharded x{world}, exchange=" +
                                           (("fused peer-memory gather+interleave kernel" + (" (serial)" if args.exchange == "p2p-serial" else " (under the next probe)")) if p2p else "NCCL all-gather + interleave"))
                           if world > 1 else "1 GPU",
                           "setup_s": round(setup_s, 1)},
                "e2e": {"value": positions * n_total / (e2e_ms / 1e3), "unit": unit, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step_mean": e2e_ms_mean, "device_timeline_ms": e2e_timeline},
                "e2e_files": e2e_files, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "tables": {"keys": [t["n_keys"] for t in tstats], "overflow_frac": sum(t["n_overflow"] for t in tstats) /
                           max(1, sum(t["n_keys"] for t in tstats)),
                           "group_tables": [eng.group_stats(u) for u in range((npg + 7) // 8)]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
