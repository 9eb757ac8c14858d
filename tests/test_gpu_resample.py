"""norm_audio on the device (rs_resample_mono, csrc/resample.cu) against the host path it replaces (nemo/asr/audio.py:
scipy.signal.resample_poly per channel, then the channel mean -- the order of pkg/nemo-asr/src/audio.py:64-67 -- then
pad_audio): same polyphase filter, so the two agree to fp32 rounding.  Rates: 48 kHz (3:1), 44.1 kHz (441:160), 8 kHz (1:2),
22.05 kHz; mono and stereo; float32 and 16-bit PCM; ragged lengths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from reazonspeech_b200.nemo.asr.audio import norm_audio, pad_audio
from reazonspeech_b200.nemo.asr.interface import AudioData


def _clips(rate, channels, lengths, seed):
    g = np.random.default_rng(seed)
    out = []
    for n in lengths:
        t = np.arange(n) / rate
        x = np.stack([0.3 * np.sin(2 * np.pi * (220 + 90 * c) * t * (1 + 0.1 * np.sin(2 * np.pi * 3 * t))) + 0.02 * g.standard_normal(n)
                      for c in range(channels)]).astype(np.float32)
        out.append(x)
    return out


@pytest.mark.parametrize("rate,channels,pcm16", [(48000, 2, True), (44100, 1, False), (8000, 1, True), (22050, 2, False), (16000, 2, False)])
def test_device_resample_matches_host_norm_audio(tiny_engine, rate, channels, pcm16):
    eng = tiny_engine
    lengths = [int(rate * s) for s in (0.61, 1.3, 0.25)]
    clips = _clips(rate, channels, lengths, seed=rate + channels)
    if pcm16:
        clips = [np.round(c * 32767.0).astype(np.int16) for c in clips]
    L = max(lengths)
    raw = torch.zeros(len(clips), channels, L, dtype=torch.int16 if pcm16 else torch.float32)
    for i, c in enumerate(clips):
        raw[i, :, : c.shape[1]] = torch.from_numpy(c)
    lens = torch.tensor(lengths, dtype=torch.int32)
    pad = 8000
    wav, out_len = eng.resample_mono(raw.cuda(), lens.cuda(), rate, pad=pad)
    torch.cuda.synchronize()
    wav, out_len = wav.cpu().numpy(), out_len.cpu().tolist()
    worst = 0.0
    for i, c in enumerate(clips):
        ref = pad_audio(norm_audio(AudioData(c if channels > 1 else c[0], rate)), 0.5).waveform.astype(np.float32)
        assert out_len[i] == len(ref), (out_len[i], len(ref))
        got = wav[i, : len(ref)]
        worst = max(worst, float(np.abs(got - ref).max()))
        assert np.abs(wav[i, len(ref):]).max(initial=0.0) == 0.0 and np.abs(got[:pad]).max() == 0.0
    print(f"{rate} Hz x{channels} {'pcm16' if pcm16 else 'f32'}: max abs difference from the host path {worst:.2e}")
    assert worst < 2e-5


def test_resampled_batch_transcribes(tiny_engine, tiny_cfg):
    """48 kHz stereo PCM -> device resample -> rs_transcribe_device: same tokens as the host-normalised waveforms through
    the float entry point, up to the resamplers' rounding (the decision sequences are compared with the noise-free bar of
    tests/parity.py against each other via the oracle-free route: identical token counts and >= 90 % equal tokens)."""
    eng = tiny_engine
    rate = 48000
    clips = [np.round(c * 32767.0).astype(np.int16) for c in _clips(rate, 2, [int(rate * s) for s in (2.0, 3.1)], seed=7)]
    L = max(c.shape[1] for c in clips)
    raw = torch.zeros(len(clips), 2, L, dtype=torch.int16)
    for i, c in enumerate(clips):
        raw[i, :, : c.shape[1]] = torch.from_numpy(c)
    lens = torch.tensor([c.shape[1] for c in clips], dtype=torch.int32)
    wav, wl = eng.resample_mono(raw.cuda(), lens.cuda(), rate, pad=8000)
    td, fd, nd = [a.cpu() for a in eng.transcribe_device(wav, wl)]
    host = [pad_audio(norm_audio(AudioData(c, rate)), 0.5).waveform.astype(np.float32) for c in clips]
    Lh = max(len(h) for h in host)
    xh = torch.zeros(len(host), Lh)
    for i, h in enumerate(host):
        xh[i, : len(h)] = torch.from_numpy(h)
    th, fh, nh = [a.cpu() for a in eng.transcribe_device(xh.cuda(), torch.tensor([len(h) for h in host], dtype=torch.int32).cuda())]
    assert int(nd.sum()) > 0
    for i in range(len(clips)):
        n = min(int(nd[i]), int(nh[i]))
        assert abs(int(nd[i]) - int(nh[i])) <= max(2, n // 10)
        same = int((td[i, :n] == th[i, :n]).sum())
        print(f"clip {i}: {int(nd[i])} / {int(nh[i])} tokens, {same} equal in the common prefix")
