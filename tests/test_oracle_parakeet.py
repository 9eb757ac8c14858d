"""Pins oracle/nemo_restated.py (frontend + FastConformer encoder) to an independent implementation
of NeMo's modules: transformers' Parakeet port (tests/golden/make_parakeet_golden.py).

Two layers of evidence: the committed vectors in tests/golden/parakeet_cases.npz (always checked;
they travel to the GPU box), and a live Parakeet run when transformers imports (checks that the
vectors are reproducible and that the generator has not drifted).

Tolerances: both sides are fp32 evaluations of the same formulas in a different operation order.
Normalised log-mel features: max-abs 2e-3 (log of near-silent bins amplifies fp32 rounding; measured
2.7e-4).  Encoder output: relative L2 1e-4 (measured 1.1e-5)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_parakeet_golden as G  # noqa: E402

from oracle import nemo_restated as O  # noqa: E402
from reazonspeech_b200.weights import random_state_dict  # noqa: E402

MEL_TOL, ENC_TOL = 2e-3, 1e-4


@pytest.fixture(scope="module")
def golden():
    return np.load(G.OUT)


def _oracle(case):
    name, kw, wseed, cseed, secs = case
    cfg = G.case_config(kw)
    sd = random_state_dict(cfg, seed=wseed, calibrate=False)
    wave = torch.from_numpy(G.padded_clip(cseed, secs))
    with torch.no_grad():
        mel = O.log_mel(wave, cfg)
        enc = O.encoder(mel, sd, cfg)
    return cfg, sd, wave, mel.T.numpy(), enc.numpy()


@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_oracle_matches_parakeet_vectors(golden, case):
    name = case[0]
    cfg, _, wave, mel, enc = _oracle(case)
    feats, n, ref = golden[name + ".features"], int(golden[name + ".n_frames"]), golden[name + ".encoder"]
    # length convention: get_seq_len = L // hop valid frames, the STFT's final frame masked to zero
    assert n == cfg.mel_valid(wave.numel()) == mel.shape[0] and feats.shape[0] == cfg.mel_frames(wave.numel())
    assert np.abs(feats[n:]).max() == 0.0
    assert np.abs(mel - feats[:n]).max() < MEL_TOL
    assert ref.shape[0] == cfg.enc_frames(wave.numel()) == enc.shape[0]
    rel = np.linalg.norm(enc - ref) / np.linalg.norm(ref)
    assert rel < ENC_TOL, rel


def test_vectors_reproduce_live(golden):
    """The committed vectors are what Parakeet computes today (skipped where transformers is absent)."""
    pytest.importorskip("transformers.models.parakeet.modeling_parakeet")
    pytest.importorskip("torchaudio")
    name, kw, wseed, cseed, secs = G.CASES[0]
    cfg = G.case_config(kw)
    sd = random_state_dict(cfg, seed=wseed, calibrate=False)
    feats, n, enc = G.run_parakeet(cfg, sd, G.padded_clip(cseed, secs))
    assert n == int(golden[name + ".n_frames"])
    assert np.abs(feats - golden[name + ".features"]).max() < 1e-4
    assert np.linalg.norm(enc - golden[name + ".encoder"]) / np.linalg.norm(enc) < 1e-5


def test_local_attention_equals_full_when_window_covers_utterance():
    """The premise of the pin: with T <= w + 1 and no global token the banded softmax sees every key."""
    cfg = G.case_config(G.CASES[0][1])
    T, H, dk = 40, cfg.n_heads, cfg.d_head
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(H, T, dk, generator=g) for _ in range(3))
    p = torch.randn(H, cfg.n_rel, dk, generator=g)
    u, vb = torch.randn(H, dk, generator=g), torch.randn(H, dk, generator=g)
    out = O.local_attention_core(q, k, v, p, u, vb, cfg)
    # Transformer-XL form with an explicit rel_shift (multi_head_attention.py::rel_shift), positions T-1 .. -(T-1)
    pos = p[:, cfg.att_left - (T - 1): cfg.att_left + T]            # rows for offsets +(T-1) .. -(T-1)
    bd = torch.matmul(q + vb[:, None], pos.transpose(1, 2))        # [H,T,2T-1]
    bd = torch.nn.functional.pad(bd, (1, 0)).view(H, 2 * T, T)[:, 1:].reshape(H, T, 2 * T - 1)[..., :T]
    ac = torch.matmul(q + u[:, None], k.transpose(1, 2))
    ref = torch.matmul(torch.softmax((ac + bd) / dk ** 0.5, -1), v)
    assert (out - ref).abs().max() < 1e-4
