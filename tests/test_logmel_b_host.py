"""CPU replay of the log-mel EXPERIMENT kernel (RS_LOGMEL_VARIANT=B, csrc/frontend.cu logmel_b_kernel): the kernel's
per-frame arithmetic -- its own register FFT (csrc/fft16.cuh, compiled here with g++), the 16 x 16 decomposition with
the inter-pass twiddle table, the real-FFT split with the partner-lane mapping, and the lane-balanced mel tables of
engine.py::frontend_tables_b -- is run lane by lane on the host and held to the oracle's un-normalised log-mel.  What
this cannot cover is CUDA itself (staging, barriers, shuffles); that is tests/experiments on a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.engine import LOGMEL_B_LANE_BINS, LOGMEL_B_LANE_TAPS, frontend_tables, frontend_tables_b
from reazonspeech_b200.synth import synth_clip
from reazonspeech_b200.weights import hann_window, mel_filterbank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    out = tmp_path_factory.mktemp("logmel_b") / "replay.so"
    src = os.path.join(ROOT, "tests", "host", "logmel_b_replay.cpp")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "reazonspeech_b200", "csrc"), src, "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("g++ could not build the replay:\n" + r.stderr[-2000:])
    return C.CDLL(str(out))


def test_lane_tables_cover_every_tap_once():
    cfg = ModelConfig()
    base = frontend_tables(cfg)
    tb = frontend_tables_b(cfg, base)
    fb = mel_filterbank(cfg).numpy()
    bins = tb["fe.b.lane_bins"].numpy().reshape(16, LOGMEL_B_LANE_BINS)
    nb = tb["fe.b.lane_nb"].numpy()
    lw = tb["fe.b.lane_w"].numpy().reshape(16, LOGMEL_B_LANE_TAPS)
    rebuilt = np.zeros_like(fb)
    seen = []
    taps = []
    for t in range(16):
        pos = 0
        for bi in range(nb[t]):
            e = int(bins[t, bi]); m, s0, c = e & 255, (e >> 8) & 1023, e >> 18
            rebuilt[m, s0:s0 + c] = lw[t, pos:pos + c]
            pos += c
            seen.append(m)
        taps.append(pos)
        assert pos <= LOGMEL_B_LANE_TAPS
    assert sorted(seen) == list(range(cfg.n_mels))                   # every filter on exactly one lane
    assert np.array_equal(rebuilt, fb)                               # and bit-identical weights
    assert max(taps) - min(taps) <= 3, taps                          # balanced: today's kernel has 26 taps on its busiest lane of 32
    tw_b = tb["fe.b.tw_b"].numpy().reshape(16, 16, 2)
    assert np.allclose(tw_b[3, 5, 0] + 1j * tw_b[3, 5, 1], np.exp(-2j * np.pi * 15 / 256), atol=1e-7)
    tw_x = tb["fe.b.tw_x"].numpy().reshape(16, 16, 2)
    assert np.allclose(tw_x[2, 7, 0] + 1j * tw_x[2, 7, 1], np.exp(-2j * np.pi * 39 / 512), atol=1e-7)


def test_replayed_frames_match_the_oracle_formula(replay):
    cfg = ModelConfig()
    base = frontend_tables(cfg)
    tb = frontend_tables_b(cfg, base)
    wave = np.pad(synth_clip(11, 1.7), 8000).astype(np.float32)
    n = len(wave)
    # the oracle's un-normalised log-mel (oracle/nemo_restated.py::log_mel up to the normalisation)
    x = torch.from_numpy(wave)
    xe = torch.cat((x[:1], x[1:] - cfg.preemph * x[:-1]))
    spec = torch.stft(xe, n_fft=cfg.n_fft, hop_length=cfg.n_window_stride, win_length=cfg.n_window_size, window=hann_window(cfg),
                      center=True, pad_mode="constant", return_complex=True)
    power = torch.view_as_real(spec).pow(2).sum(-1)                                  # [257, F]
    ref_mel = torch.log(mel_filterbank(cfg) @ power + cfg.log_zero_guard).numpy()    # [80, F]
    power = power.numpy()
    arr = lambda a, ty: np.ascontiguousarray(a, dtype=ty)
    window = arr(base["fe.window"].numpy(), np.float32)
    tw_b, tw_x = arr(tb["fe.b.tw_b"].numpy(), np.float32), arr(tb["fe.b.tw_x"].numpy(), np.float32)
    lane_w = arr(tb["fe.b.lane_w"].numpy(), np.float32)
    lane_bins, lane_nb = arr(tb["fe.b.lane_bins"].numpy(), np.int32), arr(tb["fe.b.lane_nb"].numpy(), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    replay.replay_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p]
    n_frames = cfg.mel_valid(n)
    worst_pw = worst_mel = 0.0
    for f in (0, 1, 2, 49, 50, 51, 123, n_frames - 3, n_frames - 2, n_frames - 1):     # utterance edges, pad / signal boundary, interior
        pw = np.zeros(257, np.float32); mel = np.zeros(cfg.n_mels, np.float32)
        replay.replay_frame(p(wave), n, f, cfg.n_window_stride, cfg.preemph, cfg.log_zero_guard, p(window), p(tw_b), p(tw_x),
                            p(lane_w), p(lane_bins), p(lane_nb), cfg.n_mels, p(pw), p(mel))
        assert not np.isnan(mel).any(), "a mel filter was left unwritten"
        scale = max(float(power[:, f].max()), 1e-20)
        worst_pw = max(worst_pw, float(np.abs(pw - power[:, f]).max() / scale))
        worst_mel = max(worst_mel, float(np.abs(mel - ref_mel[:, f]).max()))
    print(f"power spectrum: max error {worst_pw:.2e} of the frame's peak; log-mel max-abs {worst_mel:.2e}")
    assert worst_pw < 2e-6
    assert worst_mel < 5e-3           # un-normalised log of near-silent bins; the normalised feature divides this by sigma ~ 4
