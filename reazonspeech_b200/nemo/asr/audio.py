"""AudioData constructors and host-side normalisation, API-compatible with
pkg/nemo-asr/src/audio.py:8-83 (same names, arguments and return types).

librosa / soundfile are not dependencies here: WAV files are read with scipy (or soundfile when
it happens to be installed), resampling uses a polyphase filter.  The resampler is therefore not
bit-identical to librosa's default (soxr_hq); 16 kHz mono input -- every BASELINE configuration --
passes through untouched, exactly as in the reference (audio.py:64-67)."""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from .interface import AudioData

SAMPLERATE = 16000


def audio_from_numpy(array, samplerate):
    return AudioData(np.asarray(array), int(samplerate))


def audio_from_tensor(tensor, samplerate):
    # the reference calls tensor.numpy() (audio.py:30), which fails for CUDA tensors; accept both
    return audio_from_numpy(tensor.detach().cpu().numpy(), samplerate)


def audio_from_path(path):
    """Decode an audio file to float32 at its native rate (librosa.load(path, sr=None) semantics:
    mono float32 in [-1, 1])."""
    try:
        import soundfile
        data, rate = soundfile.read(path, dtype="float32", always_2d=True)
        wave = data.T
    except ImportError:
        from scipy.io import wavfile
        rate, data = wavfile.read(path)
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        elif data.dtype.kind == "u":
            data = (data.astype(np.float32) - 128.0) / 128.0
        wave = np.atleast_2d(data.astype(np.float32).T)
    mono = wave.mean(axis=0) if wave.shape[0] > 1 else wave[0]
    return audio_from_numpy(np.ascontiguousarray(mono, dtype=np.float32), rate)


def to_mono(waveform: np.ndarray) -> np.ndarray:
    return waveform.mean(axis=0) if waveform.ndim > 1 else waveform


def resample(waveform: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    from scipy.signal import resample_poly
    ratio = Fraction(int(target_sr), int(orig_sr))
    return resample_poly(waveform, ratio.numerator, ratio.denominator, axis=-1).astype(np.float32)


def norm_audio(audio: AudioData) -> AudioData:
    """16 kHz mono float waveform (audio.py:54-68): resample first, then average channels."""
    wave = audio.waveform
    if audio.samplerate != SAMPLERATE:
        wave = resample(wave, audio.samplerate, SAMPLERATE)
    return AudioData(to_mono(wave), SAMPLERATE)


def pad_audio(audio: AudioData, seconds: float) -> AudioData:
    """Symmetric zero padding of int(seconds * rate) samples (audio.py:70-83); never mutates the input."""
    n = int(seconds * audio.samplerate)
    return AudioData(np.pad(audio.waveform, pad_width=n, mode="constant"), audio.samplerate)
