"""Picklable evaluator stubs for the spawn-based multi-GPU evaluator tests (no GPU, no model)."""
import os

from reazonspeech_b200.evaluation import BaseEvaluator


class SpawnStub(BaseEvaluator):
    fail_rank = None          # rank that raises
    die_rank = None           # rank that exits without a word

    def _length_of(self, example):
        return len(example["audio"]["path"])

    def _evaluate(self, example, rank=None, num_gpus=None, **kw):
        if rank is not None and rank == self.fail_rank:
            raise FileNotFoundError("no checkpoint on this rank")
        if rank is not None and rank == self.die_rank:
            os._exit(3)
        return {"prediction": f"{example['audio']['path'].upper()}@{rank % num_gpus}"}

    def _evaluate_batch(self, batch, **kw):
        raise NotImplementedError
