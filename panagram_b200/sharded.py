"""Genome-sharded anchoring across the GPUs of one node (one process per GPU).

The bitmap shards by genome = by column: rank r owns the tables of a contiguous, byte-aligned
genome range, probes EVERY anchor position against them, and one exchange step assembles the
N-bit rows. torch.distributed is the plumbing (NCCL on GPUs; the same code runs
on gloo/CPU tensors for the host-logic tests); all compute is in libpkanchor.so.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_genomes: int, world: int) -> list[tuple[int, int]]:
    """Contiguous genome ranges on 8-genome (byte) boundaries, as equal as possible.
    Every rank gets the same byte width except possibly trailing ranks, which may be
    narrower or empty when N is small."""
    nbytes = (n_genomes + 7) // 8
    per = (nbytes + world - 1) // world            # bytes per rank
    out = []
    for r in range(world):
        b0 = min(r * per, nbytes)
        b1 = min((r + 1) * per, nbytes)
        out.append((min(8 * b0, n_genomes), min(8 * b1, n_genomes)))
    return out


def plane_width(n_genomes: int, world: int) -> int:
    """Bytes per row every rank contributes to the all-gather (narrow last shards are padded)."""
    return ((n_genomes + 7) // 8 + world - 1) // world


def gather_planes(local, world: int, group=None):
    """all-gather of per-rank row planes: local [n, w] uint8 -> [world, n, w]."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def interleave_planes_host(planes: np.ndarray, n_genomes: int) -> np.ndarray:
    """Host restatement of interleave_kernel for tests: [R, n, w] -> [n, ceil(N/8)]."""
    r, n, w = planes.shape
    rows = planes.transpose(1, 0, 2).reshape(n, r * w)
    return np.ascontiguousarray(rows[:, : (n_genomes + 7) // 8])


class ShardedAnchorer:
    """Rank-local engine + the exchange step. Device pointers come from torch tensors."""

    def __init__(self, k: int, n_genomes: int, rank: int, world: int, device: int, **engine_kw):
        from .engine import Engine
        self.rank, self.world, self.n_genomes = rank, world, n_genomes
        self.bounds = shard_bounds(n_genomes, world)
        self.begin, self.end = self.bounds[rank]
        if self.end <= self.begin:
            raise ValueError(f"rank {rank} owns no genomes: use at most {(n_genomes + 7) // 8} ranks")
        self.w = plane_width(n_genomes, world)
        self.row_bytes = (n_genomes + 7) // 8
        self.engine = Engine(k, n_genomes, self.begin, self.end, device=device, **engine_kw)

    def owns(self, genome: int) -> bool:
        return self.begin <= genome < self.end

    # ---- exchange over peer memory: one fused kernel instead of all-gather + interleave ----
    def setup_p2p(self, npos: int):
        """Allocate this rank's plane [npos, w], exchange IPC handles, map the peers' planes."""
        import torch
        import torch.distributed as dist
        eng = self.engine
        self.p2p_npos = npos
        self.plane = eng.device_alloc(npos * self.w)
        handles = [None] * self.world
        dist.all_gather_object(handles, eng.ipc_export(self.plane))
        self.peer_planes = [self.plane if r == self.rank else eng.ipc_open(handles[r]) for r in range(self.world)]
        self._flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{eng.cfg.device}")

    def close_p2p(self):
        import torch.distributed as dist
        dist.barrier()
        for r, p in enumerate(self.peer_planes):
            if r != self.rank:
                self.engine.ipc_close(p)
        dist.barrier()
        self.engine.device_free(self.plane)

    def probe_rows_p2p(self, d_words: int, d_mask: int, npos: int, stream: int, d_rows):
        """local probe into this rank's plane -> barrier -> gather_interleave reading every peer's plane over
        NVLink. The all-reduces are stream-ordered barriers: peers' planes are complete before they are read,
        and nobody overwrites a plane that a peer may still be reading (barrier at entry)."""
        import torch.distributed as dist
        assert npos == self.p2p_npos
        eng = self.engine
        dist.all_reduce(self._flag)
        eng.probe_device(d_words, d_mask, 0, npos, self.plane, self.w, 0, stream)
        dist.all_reduce(self._flag)
        eng.gather_interleave_device(self.peer_planes, npos, self.w, d_rows.data_ptr(), self.world * self.w, stream)
        return d_rows

    def probe_rows(self, d_words: int, d_mask: int, npos: int, stream: int, d_local, d_planes, d_rows):
        """local probe -> all-gather -> interleave. d_local [npos, w], d_planes [world, npos, w],
        d_rows [npos, world*w] are torch uint8 tensors on this rank's device."""
        import torch.distributed as dist
        eng = self.engine
        eng.probe_device(d_words, d_mask, 0, npos, d_local.data_ptr(), self.w, 0, stream)
        if self.world == 1:
            return d_local
        dist.all_gather_into_tensor(d_planes.view(-1), d_local.view(-1))
        eng.interleave_device(d_planes.data_ptr(), self.world, npos, self.w, d_rows.data_ptr(),
                              self.world * self.w, stream)
        return d_rows
