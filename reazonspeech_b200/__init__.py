"""B200-native batched ASR engine behind the reazonspeech.nemo.asr API."""
from .config import ModelConfig  # noqa: F401

__version__ = "0.1.0"
