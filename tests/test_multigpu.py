"""GPU, >= 2 devices (skipped on a single-GPU box; the builder's own 2-GPU run is committed under profiles/):
the genome-sharded path over real peer memory and NCCL, one process per GPU.

  * the position-split exchange (IPC-mapped planes read over NVLink) against the NCCL all-gather form and a single
    engine holding every genome;
  * `anchor_fasta_sharded`: the directory two ranks write together equals the directory one engine writes, file
    for file, byte for byte;
  * `python -m panagram_b200 index --gpus 2` on the N=35 golden fixture equals the reference's outputs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(k=21, n=16, length=1_400_000, seed=5):
    from panagram_b200 import synth
    anc = synth.ancestor_codes(length, seed)
    return [synth.genome_chroms(anc, g, seed, n_chroms=2, n_run=300, lower_run=1000) for g in range(n)]


def _write_fasta(path, chroms):
    with open(path, "wb") as fh:
        for name, s in chroms:
            fh.write(b">" + name.encode() + b" synthetic\n")
            b = s.tobytes()
            for o in range(0, len(b), 60):
                fh.write(b[o:o + 60] + b"\n")


def _worker(rank, world, port, tmp, q):
    import torch
    import torch.distributed as dist
    from panagram_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        k, n = 21, 16
        genomes = _case(k, n)
        sa = sharded.ShardedAnchorer(k, n, rank, world, device=rank)
        for g, chroms in enumerate(genomes):
            if sa.owns(g):
                sa.engine.reserve(g, sum(s.size for _, s in chroms))
                for _, s in chroms:
                    sa.engine.add_sequence(g, s)
        sa.engine.finalize()
        # ---- device level: one chromosome, NCCL all-gather form vs the position-split peer-memory exchange
        seq = genomes[3][0][1]
        npos = seq.size - k + 1
        rb = sa.row_bytes
        with torch.cuda.stream(sa.stream):
            st = sa.stream.cuda_stream
            d_ascii = torch.from_numpy(seq).to(dev)
            nw = sa.engine.packed_words(seq.size)
            d_words = torch.empty(nw, dtype=torch.int64, device=dev)
            d_mask = torch.empty(nw, dtype=torch.int32, device=dev)
            sa.engine.pack_device(d_ascii.data_ptr(), seq.size, d_words.data_ptr(), d_mask.data_ptr(), st)
            d_local = torch.zeros((npos, sa.w), dtype=torch.uint8, device=dev)
            d_planes = torch.empty((world, npos, sa.w), dtype=torch.uint8, device=dev)
            d_rows = torch.empty((npos, world * sa.w), dtype=torch.uint8, device=dev)
            rows = sa.probe_allgather(d_words.data_ptr(), d_mask.data_ptr(), npos, d_local, d_planes, d_rows)
            sa.stream.synchronize()
            nccl_rows = rows.cpu().numpy()
            sa._ensure_planes(npos)
            sb = sharded.slice_bounds(npos, rb, world)
            s0, s1 = sb[rank], sb[rank + 1]
            d_slice = torch.zeros((s1 - s0, rb), dtype=torch.uint8, device=dev)
            for _ in range(3):          # several steps: the planes alternate, one barrier per step
                sa.probe_exchange(d_words.data_ptr(), d_mask.data_ptr(), npos, [(s0, s1 - s0, 0)], d_slice)
            sa.stream.synchronize()
            ok_slice = bool((d_slice.cpu().numpy() == nccl_rows[s0:s1]).all())
            # the same with every gather on the side stream, under the next step's probe
            d_slice.zero_()
            for _ in range(4):
                sa.probe_exchange(d_words.data_ptr(), d_mask.data_ptr(), npos, [(s0, s1 - s0, 0)], d_slice, overlap=True)
            sa.finish_exchange()
            sa.stream.synchronize()
            ok_slice = ok_slice and bool((d_slice.cpu().numpy() == nccl_rows[s0:s1]).all())
        # ---- product level: a whole anchor directory written by both ranks
        fa = os.path.join(tmp, "g3.fa")
        if rank == 0:
            _write_fasta(fa, genomes[3])
        dist.barrier()
        out = sharded.anchor_fasta_sharded(sa, "g3", fa, os.path.join(tmp, "sharded", "g3"), genome_names=[f"g{i}" for i in range(n)])
        sa.close_p2p()
        q.put((rank, nccl_rows if rank == 0 else None, ok_slice, out["positions"], dict(sa.last)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_equals_single_engine(tmp_path):
    import torch.multiprocessing as mp
    from panagram_b200 import anchor
    from panagram_b200.engine import Engine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, _, ok, _, _ in got), "position-split exchange differs from the NCCL all-gather form"
    nccl_rows = [g[1] for g in got if g[0] == 0][0]
    k, n = 21, 16
    genomes = _case(k, n)
    eng = Engine(k, n)
    for g, chroms in enumerate(genomes):
        eng.reserve(g, sum(s.size for _, s in chroms))
        for _, s in chroms:
            eng.add_sequence(g, s)
    eng.finalize()
    want = eng.anchor_chrom(genomes[3][0][1])["bitmap1"]
    assert (nccl_rows == want).all()
    one = anchor.anchor_fasta(eng, "g3", tmp_path / "g3.fa", tmp_path / "one" / "g3", genome_names=[f"g{i}" for i in range(n)])
    assert one["positions"] == got[0][3] == got[1][3]
    files = sorted(p.name for p in (tmp_path / "one" / "g3").iterdir())
    assert sorted(p.name for p in (tmp_path / "sharded" / "g3").iterdir()) == files
    for f in files:
        assert (tmp_path / "sharded" / "g3" / f).read_bytes() == (tmp_path / "one" / "g3" / f).read_bytes(), f
    print("2-rank sharded timings (ms):", [g[4] for g in got])


@pytest.mark.skipif(_ndev() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("genome_ranks", [2, 1])
def test_index_cli_two_gpus_equals_reference(pan35, tmp_path, genome_ranks):
    """`python -m panagram_b200 index --gpus 2` on the N=35, k=31 fixture (5 row bytes: rank 0 owns 3, rank 1 owns 2)
    — genome-sharded (one group of 2 ranks) and as two replica groups taking the anchors round-robin — equals the
    unmodified reference's outputs (tests/golden)."""
    from panagram_b200 import layout
    tsv = tmp_path / "samples.tsv"
    tsv.write_text("name\tfasta\n" + "".join(f"{n}\t{pan35['fasta'][n]}\n" for n in pan35["names"]))
    out = tmp_path / "idx"
    cmd = [sys.executable, "-m", "panagram_b200", "index", str(tsv), "-o", str(out), "-k", str(pan35["k"]), "--gpus", "2",
           "--genome_ranks", str(genome_ranks), "--anchor_genomes"] + list(pan35["anchors"])
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for a in pan35["anchors"]:
        d = out / "anchor" / a
        exp = pan35["expected"][a]
        assert layout.read_bgzf(d / "bitmap.1.gz") == exp["bitmap.1"]
        assert layout.read_bgzf(d / "bitmap.100.gz") == exp["bitmap.100"]
        assert (d / "chrs.tsv").read_text() == exp["chrs.tsv"]
        assert (d / "bitsum.bins.tsv").read_text() == exp["bitsum.bins.tsv"]
        assert (d / "total_paircounts.csv").exists() and not (out / "anchor" / (a + ".tmp")).exists()
