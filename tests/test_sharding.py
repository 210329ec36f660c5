"""N>1 path on CPU: world_size-2 gloo run of the frame sharding + result gathering used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from face_detection_tflite_b200.sharding import gather_counts, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(total, rank, world)
    frames = np.arange(total)
    local = (frames[s:e] * 7 + 3) % 5                      # stand-in for per-frame face counts
    allc = gather_counts(local, rank, world)
    t = max_over_ranks(10.0 + rank, world)
    if rank == 0:
        q.put((allc.numpy().tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total = 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, t = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert counts == [int((i * 7 + 3) % 5) for i in range(total)]
    assert t == 11.0
