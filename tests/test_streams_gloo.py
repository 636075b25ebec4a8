"""N > 1 host logic on CPU: world_size-2 gloo processes shard streams, agree on timing and collect outputs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import et_streams


def test_partition_is_a_balanced_disjoint_cover():
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            parts = [et_streams.partition_streams(total, world, r) for r in range(world)]
            flat = [s for p in parts for s in p]
            assert flat == list(range(total))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        et_streams.partition_streams(4, 2, 2)
    assert et_streams.stream_seed(100, 3) == et_streams.stream_seed(100, 3) != et_streams.stream_seed(100, 4)


def _worker(rank, world, port, total_streams, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = et_streams.partition_streams(total_streams, world, rank)
        # every stream's "output" is a deterministic function of its GLOBAL id (as the seeds are)
        local = torch.stack([torch.full((4, 3), float(et_streams.stream_seed(0, s))) for s in mine]) if mine \
            else torch.zeros((0, 4, 3))
        gathered = et_streams.gather_stream_outputs(local, mine, total_streams, dist)
        slowest = et_streams.max_over_ranks(10.0 + rank, dist)
        results[rank] = (mine, gathered[:, 0, 0].tolist(), slowest)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_streams", [5, 8])
def test_two_rank_gloo_shards_streams_and_collects_outputs(total_streams):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, port, total_streams, results), nprocs=2, join=True)
    assert sorted(results[0][0] + results[1][0]) == list(range(total_streams))
    want = [float(et_streams.stream_seed(0, s)) for s in range(total_streams)]
    for rank in (0, 1):
        assert results[rank][1] == want          # ordered by global stream id on every rank
        assert results[rank][2] == 11.0          # max over ranks
