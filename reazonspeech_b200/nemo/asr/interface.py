"""Argument / result types of the ``reazonspeech.nemo.asr`` API.

Field names, order and defaults follow the reference dataclasses
(pkg/nemo-asr/src/interface.py:4-36) so results are interchangeable; the extra helpers are ours."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np


@dataclass
class AudioData:
    """A waveform (float array, mono [n] or channels-first [c, n]) and its sample rate."""
    waveform: np.ndarray
    samplerate: int

    @property
    def seconds(self) -> float:
        return self.waveform.shape[-1] / float(self.samplerate)


@dataclass
class Subword:
    """One emitted token with the time of the encoder frame that emitted it."""
    seconds: float
    token_id: int
    token: str


@dataclass
class Segment:
    start_seconds: float
    end_seconds: float
    text: str


@dataclass
class TranscribeResult:
    text: str
    subwords: List[Subword]
    segments: List[Segment]
    hypothesis: Any = None


@dataclass
class TranscribeConfig:
    verbose: bool = True
    raw_hypothesis: bool = False
