"""Host side of ``transcribe_batch`` (reazonspeech_b200/nemo/asr/transcribe.py) with the engine replaced by a CPU stand-in
that hashes exactly the bytes it is handed: batching by length, the in-place 0.5 s padding (the reference's pad_audio,
pkg/nemo-asr/src/audio.py:70-83), reuse of the staging buffers across calls, the double-buffered pipeline and the
scatter back to input order must be indistinguishable from padding every clip with numpy and transcribing it alone."""
import threading
import time

import numpy as np
import pytest
import torch

import importlib

T = importlib.import_module("reazonspeech_b200.nemo.asr.transcribe")     # the package exports a FUNCTION of the same name
from reazonspeech_b200.nemo.asr.interface import AudioData, TranscribeConfig


class _Cfg:
    blank = 50
    max_symbols = 10


def _fingerprint(row: np.ndarray, n: int):
    """Tokens / frames that depend on every sample of the valid part AND on the tail being zero."""
    valid = row[:n].astype(np.float64)
    a = int(np.abs(valid).sum() * 1e4) % 50
    b = int((valid * np.arange(1, n + 1)).sum() * 1e2) % 50
    c = n % 50
    d = int(np.count_nonzero(row[n:]))            # must be 0: stale data from an earlier, longer batch would show here
    k = n % 4                                      # a varying token count
    return [a, b, c, d][: k + 1], [0, n // 1280, n // 640, n // 320][: k + 1]


class FakeEngine:
    def __init__(self, delay=0.0):
        self.cfg = _Cfg()
        self.calls = []              # (B, L, thread name)
        self.in_flight = 0
        self.max_in_flight = 0
        self.delay = delay
        self.ws = (0, 0)

    def ensure_workspace(self, B, L):
        self.ws = (max(self.ws[0], B), max(self.ws[1], L))

    def u_max(self, L):
        return L // 1280 * self.cfg.max_symbols + 4

    def transcribe_host(self, wav, lens, U, out):
        self.in_flight += 1
        self.max_in_flight = max(self.max_in_flight, self.in_flight)
        assert wav.is_contiguous() and wav.dtype == torch.float32 and lens.dtype == torch.int32
        B, L = wav.shape
        assert B <= self.ws[0] and L <= self.ws[1], "workspace was not sized on the caller's thread"
        self.calls.append((B, L, threading.current_thread().name))
        tokens, frames, ntok = out
        assert tokens.shape == (B, U) and frames.shape == (B, U) and ntok.shape == (B,)
        time.sleep(self.delay)
        for r in range(B):
            t, f = _fingerprint(wav[r].numpy(), int(lens[r]))
            ntok[r] = len(t)
            tokens[r, : len(t)] = torch.tensor(t, dtype=torch.int32)
            frames[r, : len(f)] = torch.tensor(f, dtype=torch.int32)
        self.in_flight -= 1
        return tokens, frames, ntok


class _Tok:
    def ids_to_text(self, ids):
        return "".join(chr(0x3042 + i) for i in ids)


def _clips(n, seed=0):
    g = np.random.default_rng(seed)
    return [g.standard_normal(int(g.integers(2000, 40000))).astype(np.float32) * 0.1 for _ in range(n)]


def _alone(wave, pad):
    padded = np.pad(wave.astype(np.float32), pad)
    return _fingerprint(padded, len(padded))


@pytest.mark.parametrize("n,max_batch", [(1, 64), (5, 64), (7, 3), (64, 16), (33, 32)])
def test_batched_tokens_equal_per_clip_padding(n, max_batch):
    eng = FakeEngine()
    model = T.B200RnntModel(eng, _Tok(), max_batch=max_batch)
    clips = _clips(n, seed=n)
    for pad in (0, 8000):
        got = model.transcribe_tokens(clips, pad=pad)
        assert got == [_alone(w, pad) for w in clips]
    assert eng.max_in_flight == 1                                   # an engine is not re-entrant
    n_batches = -(-n // max_batch)
    assert len(eng.calls) == 2 * n_batches
    assert all(b <= max_batch for b, _, _ in eng.calls)


def test_staging_buffers_are_reused_and_tails_are_cleared():
    eng = FakeEngine()
    model = T.B200RnntModel(eng, _Tok(), max_batch=4)
    long = _clips(4, seed=1)
    short = [w[: len(w) // 3] for w in _clips(4, seed=2)]
    model.transcribe_tokens(long, pad=8000)
    buf = model._staging[0]._wav
    got = model.transcribe_tokens(short, pad=8000)                  # shorter rows inside the same (dirty) buffer
    assert model._staging[0]._wav is buf, "the staging buffer was reallocated for a smaller batch"
    assert got == [_alone(w, 8000) for w in short]                  # fingerprint includes 'tail is zero'


def test_pipeline_overlaps_staging_with_the_engine_call():
    eng = FakeEngine(delay=0.02)
    model = T.B200RnntModel(eng, _Tok(), max_batch=2)
    clips = _clips(8, seed=3)
    seen = []
    for idx, items in model.iter_token_batches(clips, pad=0):
        seen.append((list(idx), len(eng.calls)))
    # when batch k is handed out, batch k+1 has already been submitted (except after the last one)
    assert [calls for _, calls in seen] == [2, 3, 4, 4]
    assert sorted(i for idx, _ in seen for i in idx) == list(range(8))
    assert all(name != threading.main_thread().name for _, _, name in eng.calls)
    # a single batch has nothing to overlap with and stays on the caller's thread
    eng.calls.clear()
    assert model.transcribe_tokens(clips[:2]) == [_alone(w, 0) for w in clips[:2]]
    assert [name for _, _, name in eng.calls] == [threading.main_thread().name]


def test_transcribe_batch_matches_transcribe_per_clip():
    eng = FakeEngine()
    model = T.B200RnntModel(eng, _Tok(), max_batch=3)
    audios = [AudioData(w, 16000) for w in _clips(7, seed=4)]
    audios[2] = AudioData(np.stack([audios[2].waveform, audios[2].waveform]), 16000)      # stereo -> mono in norm_audio
    batch = T.transcribe_batch(model, audios, TranscribeConfig(verbose=False, raw_hypothesis=True))
    single = [T.transcribe(model, a, TranscribeConfig(verbose=False, raw_hypothesis=True)) for a in audios]
    for b, s in zip(batch, single):
        assert b.text == s.text and b.subwords == s.subwords and b.segments == s.segments
        assert b.hypothesis.y_sequence.tolist() == s.hypothesis.y_sequence.tolist()
        assert b.hypothesis.y_sequence[0] == _Cfg.blank and list(b.hypothesis.timestamp) == list(s.hypothesis.timestamp)
    assert T.transcribe_batch(model, []) == []


def test_models_without_the_batch_iterator_go_through_their_transcribe_method():
    class NemoLike:                                                  # the three touch points of SURVEY.md section 8b
        tokenizer = _Tok()

        def transcribe(self, tensors, batch_size, return_hypotheses, verbose):
            assert return_hypotheses and batch_size == len(tensors)
            return [T.Hypothesis.from_greedy([len(t) % 50], [1], _Cfg.blank) for t in tensors]

    audios = [AudioData(w, 16000) for w in _clips(3, seed=5)]
    res = T.transcribe_batch(NemoLike(), audios)
    assert [r.text for r in res] == [_Tok().ids_to_text([(len(a.waveform) + 16000) % 50]) for a in audios]


def test_random_batches_property():
    """Property check over random clip counts, lengths, batch limits and paddings (hypothesis, bounded examples)."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=40, deadline=None)
    @hyp.given(st.lists(st.integers(min_value=1, max_value=3000), min_size=1, max_size=12), st.integers(min_value=1, max_value=5),
               st.sampled_from([0, 1, 160, 8000]), st.integers(min_value=0, max_value=2 ** 31 - 1))
    def check(lengths, max_batch, pad, seed):
        g = np.random.default_rng(seed)
        clips = [g.standard_normal(n).astype(np.float32) for n in lengths]
        model = T.B200RnntModel(FakeEngine(), _Tok(), max_batch=max_batch)
        assert model.transcribe_tokens(clips, pad=pad) == [_alone(w, pad) for w in clips]
        assert model.transcribe_tokens(clips[::-1], pad=pad) == [_alone(w, pad) for w in clips[::-1]]     # reused staging

    check()


def test_pcm16_batches_are_staged_as_int16_and_mixed_batches_as_float():
    """HostStaging.stage: an all-int16 batch stays 16-bit PCM (scaled by 2^-15 on the device, rs_transcribe_batch_pcm16), a
    mixed batch becomes float32 with the int16 members converted the way a WAV decoder would (sample / 32768); padding and
    lengths are the same in both."""
    import numpy as np
    import torch
    from reazonspeech_b200.nemo.asr.transcribe import HostStaging
    g = np.random.default_rng(0)
    a = g.integers(-32768, 32767, 1000, dtype=np.int16)
    b = g.integers(-32768, 32767, 333, dtype=np.int16)
    st = HostStaging(pin=False)
    wav, lens = st.stage([a, b], pad=8)
    assert wav.dtype == torch.int16 and wav.shape[1] % 4 == 0 and lens.tolist() == [1016, 349]
    assert torch.equal(wav[0, 8:1008], torch.from_numpy(a)) and int(wav[0, :8].abs().sum()) == 0 and int(wav[1, 341:].abs().sum()) == 0
    f = (b.astype(np.float32) / 32768.0)
    wav2, lens2 = st.stage([a, f], pad=8)
    assert wav2.dtype == torch.float32 and lens2.tolist() == [1016, 349]
    assert np.array_equal(wav2[0, 8:1008].numpy(), a.astype(np.float32) / 32768.0) and np.array_equal(wav2[1, 8:341].numpy(), f)
    wav3, _ = st.stage([a, b], pad=8)                        # the int16 buffer is reused after a float batch
    assert wav3.dtype == torch.int16 and torch.equal(wav3[1, 8:341], torch.from_numpy(b))


def test_audio_from_path_pcm16_and_fallback_errors(tmp_path):
    import numpy as np
    from scipy.io import wavfile
    from reazonspeech_b200.nemo.asr.audio import audio_from_path, norm_audio, pad_audio
    x = (np.sin(np.arange(1600) * 0.05) * 20000).astype(np.int16)
    p = tmp_path / "m.wav"
    wavfile.write(p, 16000, x)
    f = audio_from_path(str(p))
    q = audio_from_path(str(p), pcm16=True)
    assert f.waveform.dtype == np.float32 and q.waveform.dtype == np.int16 and q.samplerate == 16000
    assert np.array_equal(f.waveform, x.astype(np.float32) / 32768.0)           # the values the device computes from the int16 samples
    assert norm_audio(q).waveform.dtype == np.int16 and pad_audio(norm_audio(q), 0.5).waveform.shape == (1600 + 16000,)
    wavfile.write(p, 8000, x)                                                    # resampling leaves PCM behind
    assert norm_audio(audio_from_path(str(p), pcm16=True)).waveform.dtype == np.float32
    bad = tmp_path / "x.webm"
    bad.write_bytes(b"not audio at all")
    import pytest
    with pytest.raises(RuntimeError, match="cannot decode"):
        audio_from_path(str(bad))
