"""N>1 host logic on CPU: two gloo ranks shard the streams by rank (contiguous blocks, no data-path
collective), reduce the step time with MAX and gather the per-rank segment lists -- the only
communication bench.py and a multi-GPU caller perform."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(n_streams: int, rank: int, world: int):
    """contiguous block of streams per rank"""
    per = (n_streams + world - 1) // world
    return range(min(n_streams, rank * per), min(n_streams, (rank + 1) * per))


def _worker(rank, world, port, n_streams, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard(n_streams, rank, world))
    # fake per-stream results: stream s has (s % 3) segments [s, s+1), ...
    counts = torch.tensor([s % 3 for s in mine], dtype=torch.int32)
    segs = torch.tensor([[s + k, s + k + 1] for s in mine for k in range(s % 3)], dtype=torch.int32).reshape(-1, 2)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, counts.tolist(), segs.tolist()))
    if rank == 0:
        ret["ms"] = float(ms)
        ret["gathered"] = gathered
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    n_streams = 11
    mp.spawn(_worker, args=(2, port, n_streams, ret), nprocs=2, join=True)
    assert ret["ms"] == 11.0                                   # max over ranks
    streams = [s for part in ret["gathered"] for s in part[0]]
    assert streams == list(range(n_streams))                   # disjoint, complete, ordered
    total = sum(len(part[2]) for part in ret["gathered"])
    assert total == sum(s % 3 for s in range(n_streams))
    assert list(shard(4, 5, 8)) == [] and list(shard(4, 3, 8)) == [3] and list(shard(9, 0, 8)) == [0, 1]
