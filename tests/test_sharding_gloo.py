"""N > 1 host logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
    import torch
    import torch.distributed as dist

    from concrete_fft_b200.sharding import job_throughput, max_over_ranks, shard_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, rows = shard_rows(batch, world, rank)
    owned = torch.zeros(batch, dtype=torch.int64)
    owned[first:first + rows] = 1
    dist.all_reduce(owned)  # test-only check that the shards tile the batch exactly once
    seconds = 0.010 * (rank + 1)  # rank 1 is the slow one
    t = max_over_ranks(seconds, dist)
    thr = job_throughput(rows, seconds, dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, first, rows, int(owned.min()), int(owned.max()), t, thr))


@pytest.mark.parametrize("batch", [65536, 4097])
def test_shards_tile_the_batch_and_timing_is_max_over_ranks(batch):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == 0 and out[0][2] + out[1][2] == batch and out[1][1] == out[0][2]
    for r in out:
        assert r[3] == 1 and r[4] == 1          # every row owned exactly once
        assert abs(r[5] - 0.020) < 1e-12         # max over ranks
        assert abs(r[6] - batch / 0.020) < 1e-6  # whole-job units / slowest rank


def test_shard_rows_edge_cases():
    from concrete_fft_b200.sharding import shard_rows

    assert shard_rows(10, 1, 0) == (0, 10)
    assert [shard_rows(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert [shard_rows(2, 4, r)[1] for r in range(4)] == [1, 1, 0, 0]
    with pytest.raises(ValueError):
        shard_rows(8, 2, 2)
