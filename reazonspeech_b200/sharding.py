"""Utterance sharding across ranks: utterances are independent, every rank
holds a full weight replica, there is no collective on the data path.  The reference scales out the
same way -- one spawned process per rank, device cuda:{rank % num_gpus}
(pkg/evaluation/src/base.py:194-212, examples/rs-nemo/eval.py:19-27) -- and merges through files;
here the optional merge is one all_gather of small Python objects (token lists)."""
from __future__ import annotations

from typing import Callable, List, Sequence


def shard_indices(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Deal utterances to ranks: longest first, each to the currently least-loaded rank
    (greedy bin packing on total samples), so ranks finish together and per-rank batches
    hold similar lengths.  Deterministic; ties broken by index."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(lengths[i])
    return shards


def run_sharded(items: Sequence, lengths: Sequence[int], fn: Callable[[List], List], rank: int, world: int,
                gather: bool = True):
    """Apply ``fn`` (a batched transcribe) to this rank's shard; with ``gather`` every rank returns the
    full result list in the original order (torch.distributed all_gather_object), else only its shard
    as {index: result}."""
    mine = shard_indices(lengths, world)[rank]
    local = dict(zip(mine, fn([items[i] for i in mine]))) if mine else {}
    if not gather or world == 1:
        return [local[i] for i in range(len(items))] if world == 1 else local
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(len(items))]
