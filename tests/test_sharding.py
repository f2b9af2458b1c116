"""CPU (gloo, world_size 2): the host logic of the genome-sharded path — shard bounds, the
all-gather plumbing and the plane -> row layout the interleave kernel implements."""
import os
import socket

import numpy as np
import pytest

from panagram_b200 import sharded


@pytest.mark.parametrize("n,world", [(8, 1), (16, 2), (64, 8), (128, 8), (35, 2), (35, 4), (9, 2), (3, 2), (70, 3)])
def test_shard_bounds_cover_on_byte_boundaries(n, world):
    b = sharded.shard_bounds(n, world)
    assert len(b) == world and b[0][0] == 0 and max(e for _, e in b) == n
    w = sharded.plane_width(n, world)
    for r, (s, e) in enumerate(b):
        assert s <= e and (e == s or (s % 8 == 0 and (e % 8 == 0 or e == n)))
        assert (e - s + 7) // 8 <= w
        if r:
            assert s == b[r - 1][1] or (s == n and e == n)
        if e > s:
            assert s == 8 * r * w           # rank r's bytes start at byte r*w of a row


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_genomes, npos, seed, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        nb = (n_genomes + 7) // 8
        full = rng.integers(0, 256, size=(npos, nb), dtype=np.uint8)
        if n_genomes % 8:
            full[:, -1] &= (1 << (n_genomes % 8)) - 1
        s, e = sharded.shard_bounds(n_genomes, world)[rank]
        w = sharded.plane_width(n_genomes, world)
        local = np.zeros((npos, w), dtype=np.uint8)            # what probe_device leaves on rank r
        local[:, : (e - s + 7) // 8] = full[:, s // 8: (e + 7) // 8]
        planes = sharded.gather_planes(torch.from_numpy(local), world).numpy()
        rows = sharded.interleave_planes_host(planes, n_genomes)
        q.put((rank, bool((rows == full).all()), rows.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_genomes", [16, 35, 128])
def test_gather_and_interleave_world2(n_genomes):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_genomes, 1000, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok for _, ok, _ in res)
    assert all(shape == (1000, (n_genomes + 7) // 8) for _, _, shape in res)
