#!/usr/bin/env python3
"""Sweep the tuning knobs of the partitioned probe on one workload with the tables built ONCE.

  python bench/k3_sweep.py [--workload configs1] [--steps 3] "k3_window=0" "k3w_variant=0,k3w_group=3" ...

For every setting: CUDA-event times of K1..K4 (pk_engine_stats) of the last of `steps` launches, the stage
time over all steps, and a checksum of the rows (all settings must agree with the first)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="configs1")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--load-factor", type=float, default=0.5)
    ap.add_argument("--out", default="")
    ap.add_argument("--chunk-positions", type=int, default=0, help="max positions per probe launch (pk_config.chunk_positions)")
    ap.add_argument("settings", nargs="*")
    args = ap.parse_args()
    import torch
    import bench as B
    from panagram_b200 import synth
    from panagram_b200.engine import Engine

    wl = B.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    k, n = wl["k"], wl["n_per_gpu"]
    t0 = time.perf_counter()
    anc = synth.ancestor_codes(wl["length"], wl["seed"])
    eng = Engine(k, n, load_factor=args.load_factor, chunk_positions=args.chunk_positions)
    anchor = None
    for g in range(n):
        chroms = B.make_genome(anc, g, wl["seed"])
        if g == 0:
            anchor = chroms
        eng.reserve(g, sum(c.size for c in chroms))
        for c in chroms:
            eng.add_sequence(g, c)
    eng.finalize()
    print(f"setup {time.perf_counter() - t0:.1f}s", flush=True)
    lens = [c.size for c in anchor]
    positions = sum(l - k + 1 for l in lens)
    cat = np.full(sum(lens) + len(lens) - 1, ord("N"), dtype=np.uint8)
    o = 0
    for c in anchor:
        cat[o:o + c.size] = c
        o += c.size + 1
    ltot = cat.size
    npos = ltot - k + 1
    rb = eng.row_bytes
    d_ascii = torch.from_numpy(cat).to(dev)
    nw = eng.packed_words(ltot)
    d_words = torch.empty(nw, dtype=torch.int64, device=dev)
    d_mask = torch.empty(nw, dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    eng.pack_device(d_ascii.data_ptr(), ltot, d_words.data_ptr(), d_mask.data_ptr(), st)
    d_rows = torch.empty((npos, rb), dtype=torch.uint8, device=dev)
    hbm = B.peaks()[0]
    alg = positions * (n * 32 + 0.375 + rb * 1.01)
    ref_sum = None
    results = []
    for spec in args.settings or ["k3_window=1"]:
        knobs = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in spec.split(",") if kv}
        try:
            refinalize = "group_tables" in knobs
            eng.tune(**knobs)
            if refinalize:
                eng.finalize()
                print("   group tables:", eng.group_stats(0), flush=True)
            d_rows.zero_()
            for _ in range(2):
                eng.probe_device(d_words.data_ptr(), d_mask.data_ptr(), 0, npos, d_rows.data_ptr(), rb, 0, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                eng.probe_device(d_words.data_ptr(), d_mask.data_ptr(), 0, npos, d_rows.data_ptr(), rb, 0, st)
            e1.record()
            torch.cuda.synchronize()
            stage = e0.elapsed_time(e1) / args.steps
            ks = eng.stats()
            w = d_rows.view(-1)[: (d_rows.numel() // 8) * 8].view(torch.int64)
            csum = int(w.sum().item()) ^ int((w[::7].to(torch.float64).sum().item()) % (1 << 52))
            if ref_sum is None:
                ref_sum = csum
            ok = csum == ref_sum
            r = dict(setting=spec, stage_ms=stage, k1=ks["k_partition_ms"], k2=ks["k_fine_ms"], k3=ks["k_probe_ms"],
                     spill=ks["k_spill_ms"], k4=ks["k_unpermute_ms"], k3_frac=alg / (ks["k_probe_ms"] / 1e3) / 1e9 / hbm,
                     stage_frac=alg / (stage / 1e3) / 1e9 / hbm, rows_equal_first=ok)
            results.append(r)
            print("%-40s stage %6.2f ms | K1 %.2f K2 %.2f K3 %.2f K4 %.2f | frac K3 %.3f stage %.3f | %s" % (
                spec, stage, r["k1"], r["k2"], r["k3"], r["k4"], r["k3_frac"], r["stage_frac"],
                "rows ok" if ok else "ROWS DIFFER"), flush=True)
        except Exception as ex:       # keep sweeping: one bad variant must not lose the others
            print(spec, "FAILED", repr(ex), flush=True)
            results.append(dict(setting=spec, error=repr(ex)))
    if args.out:
        Path(args.out).write_text(json.dumps(dict(workload=wl["name"], positions=positions, hbm_gbs=hbm, results=results), indent=1))


if __name__ == "__main__":
    main()
