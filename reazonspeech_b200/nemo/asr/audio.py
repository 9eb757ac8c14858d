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


def _decode(path):
    """-> (channels-first array [c, n], rate).  soundfile when it is installed and can open the file, scipy for RIFF/WAV,
    librosa (audioread / ffmpeg behind it: mp3, mp4, webm ...) as the last resort, like the reference (audio.py:41)."""
    errors = []
    try:
        import soundfile
        data, rate = soundfile.read(path, dtype="float32", always_2d=True)
        return data.T, rate, None
    except ImportError:
        pass
    except (OSError, RuntimeError) as exc:             # a broken libsndfile, or a container it does not know
        errors.append(f"soundfile: {exc}")
    try:
        from scipy.io import wavfile
        rate, data = wavfile.read(path)
        raw16 = data if data.dtype == np.int16 else None
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        elif data.dtype.kind == "u":
            data = (data.astype(np.float32) - 128.0) / 128.0
        return np.atleast_2d(data.astype(np.float32).T), rate, raw16
    except (ValueError, OSError) as exc:
        errors.append(f"scipy.io.wavfile: {exc}")
    try:
        import librosa
        wave, rate = librosa.load(path, sr=None, mono=False)
        return np.atleast_2d(wave), rate, None
    except ImportError:
        errors.append("librosa is not installed (needed for compressed containers such as mp3 / mp4 / webm)")
    raise RuntimeError(f"cannot decode {path!r}: " + "; ".join(errors))


def audio_from_path(path, pcm16: bool = False):
    """Decode an audio file at its native rate (librosa.load(path, sr=None) semantics: mono float32 in [-1, 1]).

    ``pcm16=True`` (an extension) keeps a mono 16-bit PCM file as int16: the engine ingests such batches as int16 and scales
    them by 2^-15 on the device -- the same values, half the bytes (see transcribe.HostStaging.stage)."""
    wave, rate, raw16 = _decode(path)
    if pcm16 and raw16 is not None and raw16.ndim == 1:
        return audio_from_numpy(np.ascontiguousarray(raw16), rate)
    mono = wave.mean(axis=0) if wave.shape[0] > 1 else wave[0]
    return audio_from_numpy(np.ascontiguousarray(mono, dtype=np.float32), rate)


def _as_float(waveform: np.ndarray) -> np.ndarray:
    """int16 PCM -> float32 in [-1, 1) as a file decoder would (sample / 32768); floats pass through."""
    return waveform.astype(np.float32) * np.float32(1.0 / 32768.0) if waveform.dtype == np.int16 else waveform


def to_mono(waveform: np.ndarray) -> np.ndarray:
    return _as_float(waveform).mean(axis=0) if waveform.ndim > 1 else waveform


def resample(waveform: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    from scipy.signal import resample_poly
    waveform = _as_float(waveform)
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
