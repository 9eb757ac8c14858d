"""CER utilities and the evaluator's host logic (no GPU: the model is a stub with the transcribe_batch seam)."""
import json

import pytest

from reazonspeech_b200.evaluation import BaseEvaluator, calculate_cer, normalize
from reazonspeech_b200.evaluation.utils import edit_distance, japanese_number


def test_normalize_strips_punctuation_and_folds_fullwidth():
    assert normalize("こんにちは、世界。") == "こんにちは世界"
    assert normalize("「ＡＢＣ」ｘｙｚ！？") == "ABCxyz"            # utils.py:15-18
    assert normalize("１２３") == "百二十三"                      # full-width digits fold, then read as a number


@pytest.mark.parametrize("text,want", [("0", "零"), ("7", "七"), ("10", "十"), ("11", "十一"), ("20", "二十"), ("105", "百五"),
                                       ("1000", "千"), ("2024", "二千二十四"), ("10000", "一万"), ("12345", "一万二千三百四十五"),
                                       ("100000000", "一億"), ("3.14", "三点一四")])
def test_japanese_number(text, want):
    assert japanese_number(text) == want


def test_edit_distance_known_values():
    assert edit_distance("kitten", "sitting") == 3
    assert edit_distance("", "abc") == 3 and edit_distance("abc", "abc") == 0
    assert edit_distance("こんにちは", "こんばんは") == 2


def test_calculate_cer():
    r = calculate_cer("今日は、いい天気です。", "今日はいい天気でした")
    assert r["length"] == 9 and r["distance"] == 2 and abs(r["cer"] - 2 / 9) < 1e-12


class _Stub(BaseEvaluator):
    calls = []

    def _evaluate(self, example, **kw):
        return {"prediction": example["audio"]["path"].upper()}

    def _evaluate_batch(self, batch, **kw):
        _Stub.calls.append(len(batch["audio"]))
        return {"predictions": [a["path"].upper() for a in batch["audio"]]}


def test_evaluate_single_and_batched_agree(tmp_path, capsys):
    rows = [{"audio": {"path": p}, "transcription": t} for p, t in (("abc", "ABC"), ("de", "DX"), ("f", "F"), ("gh", "GH"), ("ijk", "IJK"))]
    ev = _Stub(text_column="transcription")
    one = ev.evaluate(dataset=rows)
    out = tmp_path / "r.jsonl"
    _Stub.calls.clear()
    many = ev.evaluate(dataset=rows, batch_size=2, output_file=out)
    assert _Stub.calls == [2, 2, 1]                              # really batched (the reference's map is not, base.py:205-212)
    assert [r["prediction"] for r in one] == [r["prediction"] for r in many] == ["ABC", "DE", "F", "GH", "IJK"]
    assert sum(r["distance"] for r in many) == 1 and sum(r["length"] for r in many) == 11
    assert capsys.readouterr().out.count("CER: 9.09%") == 2       # base.py:223-225 report line
    lines = [json.loads(l) for l in out.read_text().splitlines()]
    assert len(lines) == 5 and lines[1]["prediction"] == "DE" and lines[1]["distance"] == 1
    assert abs(ev.calculate_cer(many, text_column="transcription") - 1 / 11) < 1e-12


def test_evaluate_without_dataset_raises():
    with pytest.raises(ValueError):
        _Stub().evaluate()


def test_multi_gpu_worker_shards_and_merges_in_order():
    """The per-GPU worker of evaluate(num_gpus > 1), run in-process with a plain queue: each rank transcribes only its
    shard (dealt by length) on device index rank % num_gpus, the merge restores input order."""
    import queue
    from reazonspeech_b200.evaluation.base import _gpu_worker
    from reazonspeech_b200.sharding import shard_indices

    seen = {}

    class Probe(_Stub):
        def _evaluate(self, example, rank=None, num_gpus=None, **kw):
            seen.setdefault(rank, []).append(example["audio"]["path"])
            return {"prediction": f"{example['audio']['path']}@{rank % num_gpus}"}

        def _length_of(self, example):
            return len(example["audio"]["path"])

    rows = [{"audio": {"path": p}} for p in ("aaaa", "b", "ccc", "dd", "eeeeee")]
    ev = Probe()
    shards = shard_indices([ev._length_of(r) for r in rows], 2)
    q = queue.Queue()
    for rank in range(2):
        _gpu_worker(rank, ev, rows, shards, None, 2, q)
    merged = {}
    while not q.empty():
        merged.update(q.get()[1])
    out = [merged[i] for i in range(len(rows))]
    assert [o.split("@")[0] for o in out] == [r["audio"]["path"] for r in rows]
    assert sorted(len(v) for v in seen.values()) == sorted(len(s) for s in shards) and set(seen) == {0, 1}


def test_multi_gpu_spawned_workers_merge_and_keep_the_callers_model():
    """evaluate(num_gpus=2) with real spawned processes (stub model): order-stable merge, and the model object the caller
    passed in is still there afterwards (the workers load their own replicas)."""
    from eval_stubs import SpawnStub
    rows = [{"audio": {"path": p}, "text": p.upper()} for p in ("aaaa", "b", "ccc", "dd", "eeeeee")]
    marker = object()
    ev = SpawnStub(model=marker)
    out = ev.evaluate(dataset=rows, num_gpus=2)
    assert [r["prediction"].split("@")[0] for r in out] == [r["text"] for r in rows]
    assert {r["prediction"].split("@")[1] for r in out} == {"0", "1"}
    assert ev.model is marker


@pytest.mark.parametrize("how", ["raise", "die"])
def test_multi_gpu_worker_failure_raises_instead_of_hanging(how):
    """A worker that raises (e.g. load_model without a checkpoint) or dies silently makes evaluate() raise -- the
    reference's datasets.map(num_proc=...) propagates worker errors too (pkg/evaluation/src/base.py:198-204)."""
    from eval_stubs import SpawnStub
    rows = [{"audio": {"path": p}, "text": p} for p in ("aaaa", "b", "ccc", "dd")]
    ev = SpawnStub()
    if how == "raise":
        ev.fail_rank = 1
    else:
        ev.die_rank = 1
    with pytest.raises(RuntimeError, match="no checkpoint on this rank" if how == "raise" else "exited with code 3"):
        ev.evaluate(dataset=rows, num_gpus=2)
