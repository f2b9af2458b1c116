#!/usr/bin/env python3
"""Times the two kernels that follow the probe stage on device-resident rows: pk_reduce_device (popcount histograms +
column sums) and pk_bgzf_compress_device (BGZF members + .gzi), CUDA events, for the row widths of the named
workloads. Rows are synthetic runs (a new random row every 1..13 positions, like a pan-k-mer bitmap).

  python bench/micro_post.py [--rows 135000000] [--out gpurun_out/micro_post.json]"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=135_000_000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from panagram_b200.engine import Engine
    eng = Engine(21, 1)
    eng.add_keys(0, np.array([1], dtype=np.uint64))
    eng.finalize()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev); g.manual_seed(5)
    res = []
    for rb, n_cols, n in ((1, 8, args.rows), (2, 16, args.rows // 2), (4, 32, args.rows // 4), (8, 64, args.rows // 8), (8, 64, args.rows),
                          (16, 128, args.rows // 16)):
        nrun = n // 7 + 1
        base = torch.randint(0, 256, (nrun, rb), dtype=torch.uint8, device=dev, generator=g)
        # a bitmap-like mix: most bytes 0xff / 0x00 with some random ones
        keep = torch.rand((nrun, rb), device=dev, generator=g) < 0.3
        base = torch.where(keep, base, torch.full_like(base, 0xff))
        reps = torch.randint(1, 14, (nrun,), device=dev, generator=g)
        rows = torch.repeat_interleave(base, reps, dim=0)[:n].contiguous()
        del base, keep, reps
        n = rows.shape[0]
        binlen = 200_000
        nb = (n + binlen - 1) // binlen
        d_hist = torch.zeros(nb * (n_cols + 1), dtype=torch.int64, device=dev)
        d_col = torch.zeros(n_cols, dtype=torch.int64, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for it in range(3):
            d_hist.zero_(); d_col.zero_()
            ev[0].record()
            eng.reduce_device(rows.data_ptr(), rb, n_cols, 0, n, binlen, d_hist.data_ptr(), d_col.data_ptr(), 0, st)
            ev[1].record()
        torch.cuda.synchronize()
        red_ms = ev[0].elapsed_time(ev[1])
        assert int(d_hist.sum().item()) == n
        nbytes = n * rb
        cap_gz, cap_gzi = eng.bgzf_bound(nbytes)
        d_gz = torch.empty(cap_gz, dtype=torch.uint8, device=dev)
        d_gzi = torch.empty(cap_gzi // 8 + 1, dtype=torch.int64, device=dev)
        d_tot = torch.zeros(2, dtype=torch.int64, device=dev)
        for it in range(3):
            ev[2].record()
            eng.bgzf_compress_device(rows.data_ptr(), nbytes, rb, d_gz.data_ptr(), d_gzi.data_ptr(), d_tot.data_ptr(), st)
            ev[3].record()
        torch.cuda.synchronize()
        gz_ms = ev[2].elapsed_time(ev[3])
        r = dict(row_bytes=rb, n_cols=n_cols, rows=n, mbytes=nbytes / 1e6, reduce_ms=red_ms, reduce_gbs=nbytes / red_ms / 1e6,
                 bgzf_ms=gz_ms, bgzf_gbs=nbytes / gz_ms / 1e6, gz_mbytes=int(d_tot[0].item()) / 1e6)
        res.append(r)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
        del rows, d_gz, d_gzi, d_hist
        torch.cuda.empty_cache()
    if args.out:
        Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
