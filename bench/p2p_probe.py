#!/usr/bin/env python3
"""Peer-memory bandwidth as the exchange kernel sees it (run under torchrun, >= 2 ranks):
each rank maps every peer's buffer over CUDA IPC and times gather_slice_kernel as a plain copy (one plane, w = 16:
one 16-byte peer load and one 16-byte store per thread) from (a) its own buffer, (b) one peer, (c) all peers at once
(n_ranks planes, w = 1: the shape of the real exchange), plus an NCCL all-gather of the same bytes for comparison.
Prints one JSON line per rank 0. Evidence for profiles/: NVLink GB/s of the gather, topology (nvidia-smi topo -m)."""
import json
import os
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    from panagram_b200.engine import Engine
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device(f"cuda:{lr}")
    dist.init_process_group("nccl", device_id=dev)
    eng = Engine(21, 8, device=lr)
    eng.add_keys(0, np.array([1], dtype=np.uint64))
    eng.finalize()
    nbytes = 256 << 20
    own = eng.device_alloc(nbytes)
    handles = [None] * world
    dist.all_gather_object(handles, eng.ipc_export(own))
    peers = [own if r == rank else eng.ipc_open(handles[r]) for r in range(world)]
    out = torch.empty(nbytes * 2, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    res = {}

    def timed(fn, reps=10):
        with torch.cuda.stream(stream):
            fn(); fn()
            dist.barrier(); stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            stream.synchronize()
            dist.barrier()
        return e0.elapsed_time(e1) / reps

    st = stream.cuda_stream
    rows16 = nbytes // 16
    ms = timed(lambda: eng.gather_slice_device([own], rows16, 16, [(0, rows16, 0)], out.data_ptr(), 16, 16, st))
    res["copy_local_GBs"] = nbytes / ms / 1e6
    peer = peers[(rank + 1) % world]
    ms = timed(lambda: eng.gather_slice_device([peer], rows16, 16, [(0, rows16, 0)], out.data_ptr(), 16, 16, st))
    res["copy_one_peer_GBs"] = nbytes / ms / 1e6
    # the exchange shape: w = 1, this rank's 1/world slice of all planes (world * slice bytes written)
    n = 135_000_000
    sl = n // world // 16 * 16
    s0 = rank * sl
    ms = timed(lambda: eng.gather_slice_device(peers, nbytes, 1, [(s0, sl, 0)], out.data_ptr(), world, world, st))
    res["exchange_w1_ms"] = ms
    res["exchange_w1_remote_GBs"] = sl * (world - 1) / ms / 1e6
    res["exchange_w1_written_GBs"] = sl * world / ms / 1e6
    # NCCL all-gather of one 135 MB plane per rank
    loc = torch.empty(n, dtype=torch.uint8, device=dev)
    allp = torch.empty(n * world, dtype=torch.uint8, device=dev)
    ms = timed(lambda: dist.all_gather_into_tensor(allp, loc))
    res["nccl_allgather_135MB_ms"] = ms
    res["nccl_allgather_recv_GBs"] = n * (world - 1) / ms / 1e6
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timed(lambda: dist.all_reduce(flag), reps=50)
    res["barrier_allreduce_us"] = ms * 1e3
    if rank == 0:
        res["world"] = world
        try:
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        except Exception as ex:   # noqa: BLE001
            res["topo"] = str(ex)
        print(json.dumps(res))
    dist.barrier()
    for r, p in enumerate(peers):
        if r != rank:
            eng.ipc_close(p)
    dist.barrier()
    eng.device_free(own)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
