"""Host logic of the one-process multi-GPU model (reazonspeech_b200/nemo/asr/multi_gpu.py) with stub replicas: utterances
are dealt by length, every replica sees only its shard, the merged result is in input order whatever order the replicas
finish in, and a failing replica raises on the caller's thread."""
import threading
import time

import numpy as np
import pytest

from reazonspeech_b200.nemo.asr.multi_gpu import MultiGpuRnntModel


class _Cfg:
    blank = 99


class _Engine:
    def __init__(self, name):
        self.device = name


class _Replica:
    def __init__(self, name, delay=0.0, fail_at=None, max_batch=2):
        self.engine, self.cfg, self.tokenizer, self.max_batch = _Engine(name), _Cfg(), None, max_batch
        self.delay, self.fail_at, self.seen, self.thread = delay, fail_at, [], None

    def iter_token_batches(self, waveforms, pad=0):
        self.thread = threading.get_ident()
        order = sorted(range(len(waveforms)), key=lambda i: len(waveforms[i]))
        for lo in range(0, len(order), self.max_batch):
            idx = order[lo:lo + self.max_batch]
            time.sleep(self.delay)
            if self.fail_at is not None and lo >= self.fail_at:
                raise RuntimeError(f"device {self.engine.device} fell over")
            self.seen.extend(len(waveforms[i]) for i in idx)
            yield idx, [([len(waveforms[i]) + pad], [0]) for i in idx]       # "token" = padded length: identifies the utterance


def _waves(lengths):
    return [np.zeros(n, np.float32) for n in lengths]


def test_results_are_in_input_order_and_every_replica_gets_a_balanced_shard():
    lengths = [50, 7, 300, 12, 12, 260, 90, 4, 33, 120, 41]
    reps = [_Replica("cuda:0", delay=0.02), _Replica("cuda:1", delay=0.0), _Replica("cuda:2", delay=0.01)]
    m = MultiGpuRnntModel(reps)
    out = m.transcribe_tokens(_waves(lengths), pad=5)
    assert [t[0] for t, f in out] == [n + 5 for n in lengths]
    assert sorted(sum((r.seen for r in reps), [])) == sorted(lengths)
    loads = [sum(r.seen) for r in reps]
    assert max(loads) - min(loads) <= max(lengths)                         # greedy bin packing by total samples
    assert len({r.thread for r in reps}) == 3 and threading.get_ident() not in {r.thread for r in reps}
    assert m.devices == ["cuda:0", "cuda:1", "cuda:2"]


def test_fewer_utterances_than_devices_and_empty_input():
    reps = [_Replica(f"cuda:{i}") for i in range(4)]
    m = MultiGpuRnntModel(reps)
    assert m.transcribe_tokens([]) == []
    out = m.transcribe_tokens(_waves([9, 3]))
    assert [t[0] for t, f in out] == [9, 3]
    assert sum(1 for r in reps if r.seen) == 2


def test_a_failing_replica_raises_on_the_callers_thread():
    reps = [_Replica("cuda:0"), _Replica("cuda:1", fail_at=2)]
    m = MultiGpuRnntModel(reps)
    with pytest.raises(RuntimeError, match="cuda:1 fell over"):
        m.transcribe_tokens(_waves([5, 6, 7, 8, 9, 10, 11, 12]))
