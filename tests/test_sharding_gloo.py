"""N>1 host logic on CPU: two (and three) gloo ranks run the PRODUCT module vadx.distributed -- contiguous stream blocks
by rank (no data-path collective), the step time reduced with MAX, and the final gather of seg_count / segments in global
stream order -- the only communication bench.py and a multi-GPU caller perform."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vadx  # noqa: F401
from vadx import distributed as D

MAX_SEG = 4


def _fake_results(streams):
    """stream s has (s % 3) segments [s + k, s + k + 1)"""
    cnt = torch.tensor([s % 3 for s in streams], dtype=torch.int32)
    seg = torch.full((len(streams), MAX_SEG, 2), -1, dtype=torch.int32)
    for i, s in enumerate(streams):
        for k in range(s % 3):
            seg[i, k, 0], seg[i, k, 1] = s + k, s + k + 1
    return cnt, seg


def _worker_weighted(rank, world, port, sizes, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D.init("gloo")
    mine = list(D.block_of(sizes, rank))
    cnt, seg = _fake_results(mine)
    all_cnt, all_seg = D.gather_segments(cnt, seg, sizes=sizes)
    ret[rank] = (all_cnt.tolist(), all_seg.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_weighted_blocks_and_unequal_gather():
    """Bandwidth-aware sharding: blocks proportional to per-rank rates, gathered back in global stream order."""
    assert D.weighted_blocks(10, [1, 1]) == [5, 5]
    assert D.weighted_blocks(4096, [23.3, 23.3, 23.3, 23.3, 35.5, 35.5, 35.5, 35.5]) == [406, 406, 406, 406, 618, 618, 618, 618]
    assert sum(D.weighted_blocks(11, [3, 1, 0.5])) == 11 and D.weighted_blocks(11, [3, 1, 0.5])[0] == 7
    with pytest.raises(ValueError):
        D.weighted_blocks(5, [0, 0])
    sizes = D.weighted_blocks(11, [3, 1, 2])
    assert sizes == [5, 2, 4] and [list(D.block_of(sizes, r)) for r in range(3)] == [[0, 1, 2, 3, 4], [5, 6], [7, 8, 9, 10]]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_weighted, args=(3, port, sizes, ret), nprocs=3, join=True)
    want_cnt, want_seg = _fake_results(list(range(11)))
    for rank in range(3):
        cnt, seg = ret[rank]
        assert cnt == want_cnt.tolist()
        got = np.asarray(seg, np.int32).reshape(11, MAX_SEG, 2)
        for s_ in range(11):
            assert np.array_equal(got[s_, :cnt[s_]], want_seg[s_, :cnt[s_]].numpy())


def _worker(rank, world, port, n_streams, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, dev = D.init("gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    mine = list(D.shard_streams(n_streams, rank, world))
    cnt, seg = _fake_results(mine)
    ms = D.max_over_ranks(10.0 + rank)
    all_cnt, all_seg = D.gather_segments(cnt, seg, n_streams)
    ret[rank] = (ms, all_cnt.tolist(), all_seg.tolist())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_streams", [(2, 11), (2, 8), (3, 4)])
def test_stream_sharding_and_segment_gather(world, n_streams):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_streams, ret), nprocs=world, join=True)
    want_cnt, want_seg = _fake_results(list(range(n_streams)))
    for rank in range(world):
        ms, cnt, seg = ret[rank]
        assert ms == 10.0 + world - 1                                         # max over ranks
        assert cnt == want_cnt.tolist()                                       # every rank holds the global result, in order
        got = np.asarray(seg, np.int32).reshape(n_streams, MAX_SEG, 2)
        for s in range(n_streams):
            assert np.array_equal(got[s, :cnt[s]], want_seg[s, :cnt[s]].numpy())


def test_shard_streams_blocks():
    assert list(D.shard_streams(4, 5, 8)) == [] and list(D.shard_streams(4, 3, 8)) == [3]
    assert list(D.shard_streams(9, 0, 8)) == [0, 1]
    for n, world in ((11, 2), (8192, 8), (5, 8), (0, 4)):
        blocks = [list(D.shard_streams(n, r, world)) for r in range(world)]
        assert [s for b in blocks for s in b] == list(range(n))               # disjoint, complete, ordered
        assert max(len(b) for b in blocks) == D.block_size(n, world)
    with pytest.raises(ValueError):
        D.shard_streams(4, 2, 2)


def test_single_process_gather_is_identity():
    cnt, seg = _fake_results([0, 1, 2])
    c2, s2 = D.gather_segments(cnt, seg)
    assert c2 is cnt and s2 is seg
    assert D.max_over_ranks(3.5) == 3.5
    assert D._cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
