"""GPU, >= 2 devices (skipped on a single-GPU box): genome-sharded anchoring over NCCL equals the
single-engine result. One process per GPU; rank r owns 8 of 16 genomes."""
import os
import socket

import numpy as np
import pytest

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
    return [[s for _, s in synth.genome_chroms(anc, g, seed, n_chroms=2, n_run=300, lower_run=1000)] for g in range(n)]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from panagram_b200.sharded import ShardedAnchorer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        k, n = 21, 16
        genomes = _case(k, n)
        sa = ShardedAnchorer(k, n, rank, world, device=rank)
        for g, chroms in enumerate(genomes):
            if sa.owns(g):
                sa.engine.reserve(g, sum(c.size for c in chroms))
                for c in chroms:
                    sa.engine.add_sequence(g, c)
        sa.engine.finalize()
        seq = genomes[3][0]
        npos = seq.size - k + 1
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        st = stream.cuda_stream
        d_ascii = torch.from_numpy(seq).to(dev)
        nw = sa.engine.packed_words(seq.size)
        d_words = torch.empty(nw, dtype=torch.int64, device=dev)
        d_mask = torch.empty(nw, dtype=torch.int32, device=dev)
        sa.engine.pack_device(d_ascii.data_ptr(), seq.size, d_words.data_ptr(), d_mask.data_ptr(), st)
        d_local = torch.zeros((npos, sa.w), dtype=torch.uint8, device=dev)
        d_planes = torch.empty((world, npos, sa.w), dtype=torch.uint8, device=dev)
        d_rows = torch.empty((npos, world * sa.w), dtype=torch.uint8, device=dev)
        rows = sa.probe_rows(d_words.data_ptr(), d_mask.data_ptr(), npos, st, d_local, d_planes, d_rows)
        torch.cuda.synchronize()
        nccl_rows = rows.cpu().numpy()
        # the same exchange through peer memory (one fused kernel, no NCCL in the data path)
        sa.setup_p2p(npos)
        d_rows2 = torch.zeros((npos, world * sa.w), dtype=torch.uint8, device=dev)
        for _ in range(2):          # twice: the entry barrier protects planes that are still being read
            sa.probe_rows_p2p(d_words.data_ptr(), d_mask.data_ptr(), npos, st, d_rows2)
        torch.cuda.synchronize()
        p2p_rows = d_rows2.cpu().numpy()
        sa.close_p2p()
        q.put((rank, nccl_rows, bool((p2p_rows == nccl_rows).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_equals_single_engine():
    import torch.multiprocessing as mp
    from panagram_b200.engine import Engine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    assert all(ok for _, _, ok in got), "peer-memory exchange differs from the NCCL exchange"
    res = {r: rows for r, rows, _ in got}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    k, n = 21, 16
    genomes = _case(k, n)
    eng = Engine(k, n)
    for g, chroms in enumerate(genomes):
        eng.reserve(g, sum(c.size for c in chroms))
        for c in chroms:
            eng.add_sequence(g, c)
    eng.finalize()
    want = eng.anchor_chrom(genomes[3][0])["bitmap1"]
    assert (res[0] == want).all() and (res[1] == want).all()
