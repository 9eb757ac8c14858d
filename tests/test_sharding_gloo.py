"""N>1 host logic on CPU: world_size-2 gloo run of the utterance sharding + order-stable gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reazonspeech_b200.sharding import run_sharded, shard_indices


def test_shard_indices_balanced_and_complete():
    lengths = [480000, 80000, 160000, 320000, 480000, 16000, 240000]
    for world in (1, 2, 4, 8):
        shards = shard_indices(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)
    assert shard_indices(lengths, 2) == shard_indices(lengths, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = [torch.full((n,), float(i)) for i, n in enumerate((50, 10, 30, 20, 40))]
    fake_transcribe = lambda batch: [(int(x[0].item()), len(x)) for x in batch]      # stands in for the engine
    out = run_sharded(items, [len(x) for x in items], fake_transcribe, rank, world)
    q.put((rank, out))
    dist.destroy_process_group()


def test_world2_gather_is_order_stable():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [(0, 50), (1, 10), (2, 30), (3, 20), (4, 40)]
    assert got[0] == want and got[1] == want
