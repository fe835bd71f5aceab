"""CPU, world_size 2 over gloo: host-side logic of the sharded export (index sharding, padding,
the single gather, reordering).  The per-item work is a CPU stand-in; the CUDA path is covered by
the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from panoptic_forecasting_b200.export import ShardedExporter, padded_local_count, shard_indices


def fake_forecast(idx, h=6, w=10):
    out = torch.empty((len(idx), h, w), dtype=torch.uint8)
    for j, i in enumerate(idx):
        out[j] = (torch.arange(h * w, dtype=torch.int64).reshape(h, w) * 7 + i * 13) % 251
    return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = ShardedExporter(fake_forecast, n_items, 6, 10, "cpu", batch_size=2).run()
        got = ex.gather()
        if rank == 0:
            ref = fake_forecast(list(range(n_items)))
            q.put(bool(torch.equal(got, ref)))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 4, 1])
def test_sharded_export_world2_equals_single_rank(n_items):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shard_arithmetic():
    for n in (0, 1, 5, 8, 500):
        for world in (1, 2, 4, 8):
            all_items = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert all_items == list(range(n))
            assert max(len(shard_indices(n, r, world)) for r in range(world)) == (padded_local_count(n, world) if n else 0)
    single = ShardedExporter(fake_forecast, 5, 6, 10, "cpu", rank=0, world=1).run().gather()
    assert torch.equal(single, fake_forecast(list(range(5))))
