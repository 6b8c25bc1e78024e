"""world_size-2 gloo test of the N>1 host logic: contiguous sharding + the optional final feature gather."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_music_style_transfer_b200 import sharding
    a, b = sharding.shard_range(n_items, rank, world)
    local = torch.arange(a, b, dtype=torch.float32).view(-1, 1).repeat(1, 3)  # "features" of my clips
    full = sharding.gather_features(local, n_items)
    ok = full.shape == (n_items, 3) and torch.equal(full[:, 0], torch.arange(n_items, dtype=torch.float32))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _run(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_gather_equal_shards():
    _run(8)


def test_gather_ragged_shards():
    _run(7)
