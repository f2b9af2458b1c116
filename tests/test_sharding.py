"""CPU (gloo, world_size 2): the host logic of the genome-sharded path — shard bounds, the position slices and
their segments, the BGZF part merge, the all-gather plumbing, the collective file assembly, and the per-chunk body
of the exchange kernel compiled for the host."""
import gzip
import os
import socket
import struct
import subprocess

import numpy as np
import pytest

from panagram_b200 import layout, sharded

from conftest import ROOT


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


@pytest.mark.parametrize("total,rb,world", [(134_999_900, 1, 8), (134_999_900, 8, 8), (1000, 2, 2), (200_000, 3, 4),
                                            (0, 1, 2), (65280, 1, 4), (65281, 1, 2), (10_000_000, 16, 8), (5_000_000, 5, 3)])
def test_slice_bounds_fall_on_bgzf_member_boundaries(total, rb, world):
    b = sharded.slice_bounds(total, rb, world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == total and all(x <= y for x, y in zip(b, b[1:]))
    for x in b[1:-1]:
        assert x == total or (x * rb) % sharded.BGZF_PAYLOAD == 0
    if total * rb >= world * 4 * sharded.BGZF_PAYLOAD * rb:      # balanced to within one aligned unit
        sizes = np.diff(b)
        assert sizes.max() - sizes.min() <= sharded.BGZF_PAYLOAD * rb


def test_stream_segments_and_pieces_against_brute_force():
    rng = np.random.default_rng(3)
    for _ in range(200):
        c = int(rng.integers(1, 7))
        nks = [int(x) for x in rng.integers(0, 400, size=c)]
        gaps = rng.integers(1, 64, size=c)
        cat_off, o = [], 0
        for nk, g in zip(nks, gaps):
            cat_off.append(o)
            o += nk + int(g)
        # stream row -> (cat row, chromosome, position)
        table = [(cat_off[ci] + p, ci, p) for ci in range(c) for p in range(nks[ci])]
        total = len(table)
        s0 = int(rng.integers(0, total + 1)); s1 = int(rng.integers(s0, total + 1))
        got = {}
        for src, n, dst in sharded.stream_segments(cat_off, nks, s0, s1):
            for i in range(n):
                assert dst + i not in got
                got[dst + i] = src + i
        assert got == {i - s0: table[i][0] for i in range(s0, s1)}
        gotp = {}
        for ci, p0, n, r0 in sharded.slice_pieces(nks, s0, s1):
            for i in range(n):
                gotp[r0 + i] = (ci, p0 + i)
        assert gotp == {i - s0: table[i][1:] for i in range(s0, s1)}


def _bgzf_part(tmp, data: bytes, tag: str):
    w = layout.BgzfWriter(tmp / f"{tag}.gz", threads=2)
    w.write(data)
    w.close(tmp / f"{tag}.gzi")
    return np.frombuffer((tmp / f"{tag}.gz").read_bytes(), dtype=np.uint8), np.frombuffer((tmp / f"{tag}.gzi").read_bytes(), dtype=np.uint8)


@pytest.mark.parametrize("total_rows,rb,world", [(300_000, 1, 2), (300_000, 1, 4), (70_000, 8, 8), (10, 1, 4), (0, 1, 2), (200_000, 3, 3)])
def test_bgzf_parts_concatenate_into_the_single_writer_file(tmp_path, total_rows, rb, world):
    rng = np.random.default_rng(total_rows + world)
    data = (rng.integers(0, 3, size=total_rows * rb, dtype=np.uint8) * 85).tobytes()
    whole_gz, whole_gzi = _bgzf_part(tmp_path, data, "whole")
    b = sharded.slice_bounds(total_rows, rb, world)
    parts = [_bgzf_part(tmp_path, data[b[r] * rb:b[r + 1] * rb], f"p{r}") for r in range(world)]
    live = [r for r in range(world) if b[r + 1] > b[r]] or [0]
    offs, total = sharded.merge_bgzf_parts([(parts[r][0].size, parts[r][1].size) for r in live], [b[r] * rb for r in live])
    out = bytearray(total)
    for j, r in enumerate(live):
        d = parts[r][0] if j == len(live) - 1 else parts[r][0][:-sharded.BGZF_EOF_LEN]
        out[offs[j]:offs[j] + d.size] = d.tobytes()
    assert bytes(out) == whole_gz.tobytes()              # member for member the file one writer produces
    assert gzip.decompress(bytes(out)) == data
    gzi = sharded.merge_gzi([parts[r][1].tobytes() for r in live], offs, [b[r] * rb for r in live],
                            [(b[r + 1] - b[r]) * rb for r in live])
    assert gzi == whole_gzi.tobytes()
    (n,) = struct.unpack_from("<Q", gzi)
    assert n == max((len(data) + 0xFF00 - 1) // 0xFF00 - 1, 0)


def test_gather_kernel_body_on_the_host(tmp_path):
    """pk_gather.cuh's per-chunk body (what gather_slice_kernel runs per thread), compiled for the host and compared
    with a restatement of its contract over ranks 1..16, widths 1..16, ragged segments, padded strides, narrow
    last shards and the plane's last chunk."""
    exe = tmp_path / "gather_host_check"
    subprocess.check_call(["g++", "-O2", "-Wall", "-Wno-unknown-pragmas", "-Werror", "-o", str(exe),
                           str(ROOT / "tests" / "native" / "gather_host_check.cpp")])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_genomes, npos, seed, tmp, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        nb = (n_genomes + 7) // 8
        full = rng.integers(0, 4, size=(npos, nb), dtype=np.uint8) * 85
        if n_genomes % 8:
            full[:, -1] &= (1 << (n_genomes % 8)) - 1
        s, e = sharded.shard_bounds(n_genomes, world)[rank]
        w = sharded.plane_width(n_genomes, world)
        local = np.zeros((npos, w), dtype=np.uint8)            # what the probe leaves in rank r's plane
        local[:, : (e - s + 7) // 8] = full[:, s // 8: (e + 7) // 8]
        planes = sharded.gather_planes(torch.from_numpy(local), world).numpy()
        rows = sharded.interleave_planes_host(planes, n_genomes)
        ok = bool((rows == full).all())
        # position-split product path: this rank compresses ITS slice, all ranks write one file
        b = sharded.slice_bounds(npos, nb, world)
        mine = rows[b[rank]:b[rank + 1]].tobytes()
        from pathlib import Path
        tmp = Path(tmp)
        gz, gzi = _bgzf_part(tmp, mine, f"rank{rank}")
        sharded.assemble_bitmap(rank, world, None, tmp / "bitmap.1.gz", tmp / "bitmap.1.gzi", gz, gzi, b[rank] * nb,
                                (b[rank + 1] - b[rank]) * nb)
        if rank == 0:
            whole_gz, whole_gzi = _bgzf_part(tmp, full.tobytes(), "whole")
            ok = ok and (tmp / "bitmap.1.gz").read_bytes() == whole_gz.tobytes()
            ok = ok and (tmp / "bitmap.1.gzi").read_bytes() == whole_gzi.tobytes()
            ok = ok and layout.query_bytes(tmp / "bitmap.1.gz", tmp / "bitmap.1.gzi", 70_000, 5000) == full.tobytes()[70_000:75_000]
        q.put((rank, ok, rows.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_genomes", [16, 35, 128])
def test_gather_interleave_and_file_assembly_world2(n_genomes, tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    npos = 200_000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_genomes, npos, 11, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok for _, ok, _ in res)
    assert all(shape == (npos, (n_genomes + 7) // 8) for _, _, shape in res)


def test_grid_shape():
    assert sharded.grid_shape(8, None) == (8, 1)
    assert sharded.grid_shape(8, 2) == (2, 4)
    assert sharded.grid_shape(1, None) == (1, 1)
    with pytest.raises(ValueError):
        sharded.grid_shape(8, 3)
