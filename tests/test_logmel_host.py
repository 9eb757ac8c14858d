"""CPU replay of the fused log-mel kernel (csrc/logmel.cu): its per-frame arithmetic -- the register FFT (csrc/fft16.cuh), the
16 x 16 decomposition with the inter-pass twiddle table, the PAIRED real-FFT split with the partner-lane exchange
(csrc/logmel_frame.cuh), and the slot-dealt mel tables of reazonspeech_b200/logmel_tables.py -- compiled with g++ and run
lane by lane on the host against the oracle's un-normalised log-mel.  What this cannot cover is CUDA itself (staging,
barriers, shuffles, the statistics ticket); that is tests/test_gpu_kernels.py on a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.logmel_tables import LANES, META_OUT, META_START, logmel_tables, mel_slots
from reazonspeech_b200.synth import synth_clip
from reazonspeech_b200.weights import hann_window, mel_filterbank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    out = tmp_path_factory.mktemp("logmel") / "replay.so"
    src = os.path.join(ROOT, "tests", "host", "logmel_replay.cpp")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "reazonspeech_b200", "csrc"), src, "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("g++ could not build the replay:\n" + r.stderr[-2000:])
    return C.CDLL(str(out))


def test_slot_tables_cover_every_tap_once():
    cfg = ModelConfig()
    fb = mel_filterbank(cfg).numpy().astype(np.float32)
    w, meta, slots = mel_slots(fb)
    assert meta[0] == 5 and meta[1] == w.shape[0] == int(meta[8:8 + 5].sum())
    assert w.shape[0] <= 40, "taps per lane (31.25 unpadded)"
    rebuilt = np.zeros_like(fb)
    seen, base = [], 0
    for s in range(meta[0]):
        c = int(meta[8 + s])
        for t in range(LANES):
            m, first = int(meta[META_OUT + s * LANES + t]), int(meta[META_START + s * LANES + t])
            assert 0 <= first and first + c <= 257                     # every tap reads a real bin
            if m < cfg.n_mels:
                rebuilt[m, first:first + c] += w[base:base + c, t]
                seen.append(m)
            else:
                assert not w[base:base + c, t].any()
        base += c
    assert sorted(seen) == list(range(cfg.n_mels))                     # every filter on exactly one (slot, lane)
    assert np.array_equal(rebuilt, fb)                                 # bit-identical weights
    tb = logmel_tables(cfg)
    tw_b = tb["fe.tw_b"].numpy().reshape(16, 16, 2)
    assert np.allclose(tw_b[3, 5, 0] + 1j * tw_b[3, 5, 1], np.exp(-2j * np.pi * 15 / 256), atol=1e-7)
    tw_x = tb["fe.tw_x"].numpy().reshape(8, 16, 2)
    assert np.allclose(tw_x[2, 7, 0] + 1j * tw_x[2, 7, 1], np.exp(-2j * np.pi * 39 / 512), atol=1e-7)
    assert np.array_equal(tb["fe.mel_w"].numpy().reshape(-1, 16) * 4, w)


def test_slot_tables_with_a_partial_last_slot():
    """n_mels not a multiple of 16: lanes without a filter in the last slot carry zero weights and park their result."""
    cfg = ModelConfig()
    fb = mel_filterbank(cfg).numpy().astype(np.float32)[:70]
    w, meta, slots = mel_slots(fb)
    assert meta[0] == 5 and len(slots[-1]) == 6
    outs = meta[META_OUT:META_OUT + 5 * LANES].reshape(5, LANES)
    assert sorted(int(m) for m in outs.reshape(-1) if m < 70) == list(range(70))
    assert all(int(m) >= 128 for m in outs[4, 6:])


def test_replayed_frames_match_the_oracle_formula(replay):
    cfg = ModelConfig()
    tb = logmel_tables(cfg)
    wave = np.pad(synth_clip(11, 1.7), 8000).astype(np.float32)
    n = len(wave)
    x = torch.from_numpy(wave)
    xe = torch.cat((x[:1], x[1:] - cfg.preemph * x[:-1]))
    spec = torch.stft(xe, n_fft=cfg.n_fft, hop_length=cfg.n_window_stride, win_length=cfg.n_window_size, window=hann_window(cfg),
                      center=True, pad_mode="constant", return_complex=True)
    power = torch.view_as_real(spec).pow(2).sum(-1)                                  # [257, F]
    ref_mel = torch.log(mel_filterbank(cfg) @ power + cfg.log_zero_guard).numpy()    # [80, F]
    power = power.numpy()
    arr = lambda a, ty: np.ascontiguousarray(a, dtype=ty)
    window = arr(tb["fe.window"].numpy(), np.float32)
    tw_b, tw_x = arr(tb["fe.tw_b"].numpy(), np.float32), arr(tb["fe.tw_x"].numpy(), np.float32)
    mel_w, meta = arr(tb["fe.mel_w"].numpy(), np.float32), arr(tb["fe.mel_meta"].numpy(), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    replay.replay_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
    n_frames = cfg.mel_valid(n)
    worst_pw = worst_mel = 0.0
    for f in (0, 1, 2, 49, 50, 51, 123, n_frames - 3, n_frames - 2, n_frames - 1, n_frames):   # edges, pad / signal boundary, interior, the masked frame
        pw = np.zeros(257, np.float32); mel = np.zeros(cfg.n_mels, np.float32)
        replay.replay_frame(p(wave), n, f, cfg.n_window_stride, cfg.preemph, cfg.log_zero_guard, p(window), p(tw_b), p(tw_x),
                            p(mel_w), p(meta), cfg.n_mels, p(pw), p(mel))
        assert not np.isnan(pw).any(), "a power bin was written zero or several times"
        assert not np.isnan(mel).any(), "a mel filter was left unwritten"
        scale = max(float(power[:, f].max()), 1e-20)
        worst_pw = max(worst_pw, float(np.abs(pw / 4 - power[:, f]).max() / scale))
        worst_mel = max(worst_mel, float(np.abs(mel - ref_mel[:, f]).max()))
    print(f"power spectrum: max error {worst_pw:.2e} of the frame's peak; log-mel max-abs {worst_mel:.2e}")
    assert worst_pw < 2e-6
    assert worst_mel < 5e-3           # un-normalised log of near-silent bins; the normalised feature divides this by sigma ~ 4
