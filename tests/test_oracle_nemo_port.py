"""Pins oracle/nemo_restated.py (subsampling + Conformer layers with LIMITED attention context, ragged batches) to NeMo's
ConformerEncoder as ported into vLLM with NeMo's own parameter names (tests/golden/make_nemo_port_golden.py).

The committed vectors in tests/golden/nemo_port_cases.npz are always checked (they travel to the GPU box, where the
engine is held to them as well: tests/test_gpu_nemo_port.py); a live run of the port re-derives one batch when vllm
imports.  Both sides are fp32 evaluations of the same formulas: relative L2 <= 1e-5 (measured 3e-7)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_nemo_port_golden as G  # noqa: E402

from oracle import nemo_restated as O  # noqa: E402
from reazonspeech_b200.weights import random_state_dict  # noqa: E402

TOL = 1e-5


@pytest.fixture(scope="module")
def golden():
    return np.load(G.OUT)


@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_oracle_matches_nemo_port_vectors(golden, case):
    name, kw, wseed, clips = case
    cfg = G.case_config(kw)
    sd = random_state_dict(cfg, seed=wseed, calibrate=False)
    assert int(golden[f"{name}/n"]) == len(clips)
    longest = 0
    for i, (cseed, secs) in enumerate(clips):
        wave = torch.from_numpy(G.padded_clip(cseed, secs))
        with torch.no_grad():
            enc = O.encoder(O.log_mel(wave, cfg), sd, cfg).numpy()
        ref = golden[f"{name}/enc{i}"]
        assert ref.shape == enc.shape == (cfg.enc_frames(wave.numel()), cfg.d_model)
        rel = np.linalg.norm(enc - ref) / np.linalg.norm(ref)
        assert rel < TOL, (name, i, rel)
        longest = max(longest, ref.shape[0])
    assert longest > 3 * (cfg.att_left + cfg.att_right + 1), "the case no longer exercises the attention window"


def test_nemo_named_state_dict_loads_strictly_and_vectors_reproduce_live(golden):
    """The port accepts the oracle's NeMo-named encoder tensors with strict=True, and computes the stored vectors today
    (skipped where vllm is absent, e.g. on a minimal box)."""
    pytest.importorskip("vllm.model_executor.models.cohere_asr")
    name, kw, wseed, clips = G.CASES[1]
    _, _, _, outs, out_len = G.run_case(kw, wseed, clips)
    for i, o in enumerate(outs):
        ref = golden[f"{name}/enc{i}"]
        assert out_len[i] == ref.shape[0]
        assert np.linalg.norm(o - ref) / np.linalg.norm(ref) < 1e-6


def test_oracle_matches_filterbank_port_and_end_to_end(golden):
    """NeMo's FilterbankFeatures (vLLM's copy, fp32 buffers) on a ragged batch, then the encoder port on ITS features:
    the oracle's batch-of-one frontend must give the same valid frames (max-abs 2e-3: fp32 log of near-silent bins;
    measured 2.6e-4) and the oracle's whole N1-N7 path the same encoder output (relative L2 1e-4; measured 1.2e-5)."""
    name, kw, wseed, _ = G.CASES[0]
    cfg = G.case_config(kw)
    sd = random_state_dict(cfg, seed=wseed, calibrate=False)
    assert int(golden["frontend/n"]) == len(G.FRONTEND_CLIPS)
    for i, (cseed, secs) in enumerate(G.FRONTEND_CLIPS):
        wave = torch.from_numpy(G.padded_clip(cseed, secs))
        with torch.no_grad():
            mel = O.log_mel(wave, cfg)
            enc = O.encoder(mel, sd, cfg).numpy()
        ref_mel, ref_enc = golden[f"frontend/mel{i}"], golden[f"frontend/enc{i}"]
        assert ref_mel.shape == (cfg.mel_valid(wave.numel()), cfg.n_mels) == tuple(mel.T.shape)      # get_seq_len = L // hop
        assert np.abs(mel.T.numpy() - ref_mel).max() < 2e-3
        assert ref_enc.shape == enc.shape
        assert np.linalg.norm(enc - ref_enc) / np.linalg.norm(ref_enc) < 1e-4
