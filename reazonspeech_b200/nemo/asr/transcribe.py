"""``load_model`` / ``transcribe`` with the reference's signatures (pkg/nemo-asr/src/transcribe.py:9-60)
on top of the B200 engine, plus the batched ``transcribe_batch`` the reference lacks
(its evaluators leave ``_evaluate_batch`` unimplemented, pkg/evaluation/examples/rs-nemo/eval.py:31-32).

The object returned by ``load_model`` is a duck-typed stand-in for NeMo's EncDecRNNTBPEModel at
exactly the three points the reference touches it:
``model.transcribe(list_of_tensors, batch_size=..., return_hypotheses=True, verbose=...)``,
``hyp.y_sequence`` / ``hyp.timestamp`` and ``model.tokenizer.ids_to_text``; the reference's own
transcribe()/decode_hypothesis() therefore run unmodified on it."""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from ...config import ModelConfig
from ...engine import Engine
from ...tokenizer import PieceTableTokenizer, SentencePieceTokenizer, synthetic_pieces
from ...weights import load_nemo_archive, random_state_dict
from .audio import norm_audio, pad_audio
from .decode import PAD_SECONDS, build_result, decode_hypothesis
from .interface import AudioData, TranscribeConfig, TranscribeResult

HF_REPO = "reazon-research/reazonspeech-nemo-v2"
ENV_CHECKPOINT = "REAZONSPEECH_NEMO_CHECKPOINT"
ENV_SYNTHETIC = "REAZONSPEECH_B200_SYNTHETIC"


@dataclass
class Hypothesis:
    """The two NeMo Hypothesis fields the reference reads (decode.py:40,44), ALSD-shaped:
    y_sequence = [blank, tok_0, ...]; timestamp[i] = frame_i + i + 1 so that decode.py:48's
    ``step - idx - 1`` recovers the emitting encoder frame."""
    y_sequence: torch.Tensor
    timestamp: List[int]
    score: float = 0.0

    @staticmethod
    def from_greedy(tokens: Sequence[int], frames: Sequence[int], blank: int) -> "Hypothesis":
        y = torch.tensor([blank, *[int(t) for t in tokens]], dtype=torch.long)
        return Hypothesis(y, [int(f) + i + 1 for i, f in enumerate(frames)])


class B200RnntModel:
    """Engine + tokenizer behind NeMo's model surface."""

    def __init__(self, engine: Engine, tokenizer, max_batch: int = 64):
        self.engine = engine
        self.cfg = engine.cfg
        self.tokenizer = tokenizer
        self.max_batch = max_batch

    # -- token-level batched path
    def transcribe_tokens(self, waveforms: Sequence[np.ndarray]):
        """Padded 16 kHz mono float32 waveforms -> [(tokens, frames)] in input order.

        Utterances are sorted by length and cut into batches of at most ``max_batch`` so padding
        waste stays small; results are scattered back to the caller's order."""
        order = sorted(range(len(waveforms)), key=lambda i: len(waveforms[i]))
        results = [None] * len(waveforms)
        for lo in range(0, len(order), self.max_batch):
            idx = order[lo:lo + self.max_batch]
            L = max(len(waveforms[i]) for i in idx)
            host = torch.zeros(len(idx), L, dtype=torch.float32).pin_memory()
            for r, i in enumerate(idx):
                host[r, : len(waveforms[i])] = torch.from_numpy(np.ascontiguousarray(waveforms[i], dtype=np.float32))
            lens = torch.tensor([len(waveforms[i]) for i in idx], dtype=torch.int32)
            tokens, frames, ntok = self.engine.transcribe_host(host, lens)
            for r, i in enumerate(idx):
                n = int(ntok[r])
                results[i] = (tokens[r, :n].tolist(), frames[r, :n].tolist())
        return results

    # -- NeMo's call shape (transcribe.py:48-53)
    def transcribe(self, audio, batch_size: int = 1, return_hypotheses: bool = True, verbose: bool = True, **_):
        waves = [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in audio]
        out = [Hypothesis.from_greedy(t, f, self.cfg.blank) for t, f in self.transcribe_tokens(waves)]
        if return_hypotheses:
            return out
        return [self.tokenizer.ids_to_text(h.y_sequence.tolist()[1:]) for h in out]


def _find_checkpoint() -> Optional[str]:
    env = os.environ.get(ENV_CHECKPOINT)
    if env:
        return env
    hub = os.path.expanduser(os.environ.get("HF_HOME", "~/.cache/huggingface"))
    hits = glob.glob(os.path.join(hub, "hub", "models--" + HF_REPO.replace("/", "--"), "snapshots", "*", "*.nemo"))
    return sorted(hits)[-1] if hits else None


def load_model(device=None, *, checkpoint: Optional[str] = None, synthetic: Optional[bool] = None,
               config: Optional[ModelConfig] = None, seed: int = 0, max_batch: int = 64):
    """Load the ReazonSpeech FastConformer-RNNT onto a B200.

    ``device``: None / "cuda" / "cuda:N" as in the reference (transcribe.py:9-22, eval.py:26).
    "cpu" raises: this engine has no CPU path.  Weights come from ``checkpoint`` (a .nemo file),
    $REAZONSPEECH_NEMO_CHECKPOINT or the local Hugging Face cache of reazonspeech-nemo-v2.
    With ``synthetic=True`` (or $REAZONSPEECH_B200_SYNTHETIC=1) seeded random weights of the same
    architecture are used instead -- the only option offline."""
    if device is None:
        device = "cuda"
    if str(device).startswith("cpu"):
        raise RuntimeError("reazonspeech_b200: device='cpu' is not supported (hand-written sm_100a kernels only)")
    if synthetic is None:
        synthetic = os.environ.get(ENV_SYNTHETIC, "") not in ("", "0")
    path = checkpoint or (None if synthetic else _find_checkpoint())
    if path is not None:
        cfg, sd, tok = load_nemo_archive(path)
        tokenizer = SentencePieceTokenizer(tok) if tok else PieceTableTokenizer(synthetic_pieces(cfg.vocab_size))
    elif synthetic:
        cfg = config or ModelConfig()
        sd = random_state_dict(cfg, seed)
        tokenizer = PieceTableTokenizer(synthetic_pieces(cfg.vocab_size))
    else:
        raise FileNotFoundError(
            f"no .nemo checkpoint for {HF_REPO}: pass checkpoint=..., set ${ENV_CHECKPOINT}, populate the Hugging Face "
            f"cache, or request seeded synthetic weights with synthetic=True / ${ENV_SYNTHETIC}=1")
    return B200RnntModel(Engine(cfg, sd, str(device)), tokenizer, max_batch=max_batch)


def _prepare(audio: AudioData) -> np.ndarray:
    return pad_audio(norm_audio(audio), PAD_SECONDS).waveform.astype(np.float32, copy=False)


def transcribe(model, audio: AudioData, config: Optional[TranscribeConfig] = None) -> TranscribeResult:
    """One utterance, same contract as the reference (transcribe.py:30-60)."""
    if config is None:
        config = TranscribeConfig()
    wave = torch.from_numpy(_prepare(audio))
    hyp = model.transcribe([wave], batch_size=1, return_hypotheses=True, verbose=config.verbose)[0]
    result = decode_hypothesis(model, hyp)
    if config.raw_hypothesis:
        result.hypothesis = hyp
    return result


def transcribe_batch(model, audios: Sequence[AudioData], config: Optional[TranscribeConfig] = None) -> List[TranscribeResult]:
    """Many utterances through one or a few engine launches; results in input order."""
    if config is None:
        config = TranscribeConfig()
    waves = [torch.from_numpy(_prepare(a)) for a in audios]
    hyps = model.transcribe(waves, batch_size=len(waves), return_hypotheses=True, verbose=config.verbose)
    out = []
    for hyp in hyps:
        r = decode_hypothesis(model, hyp)
        if config.raw_hypothesis:
            r.hypothesis = hyp
        out.append(r)
    return out
