"""RS_LN_FOLD=1 (EXPERIMENT, written without GPU time left in round 1 -- see DESIGN.md section 8): three of the five
LayerNorms of a Conformer layer folded into their consumer GEMMs.  Not part of the default `-m gpu` suite on purpose:
run explicitly with

    RS_RUN_EXPERIMENTS=1 python -m pytest tests/experiments/test_gpu_ln_fold.py -m gpu -q -s

It checks the folded encoder against the default engine (same weights) and against the fp32 oracle at the encoder's
usual tolerance (relative L2 <= 2e-2)."""
import os

import numpy as np
import pytest
import torch

from reazonspeech_b200.synth import synth_clip

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("RS_RUN_EXPERIMENTS") != "1", reason="experiment: set RS_RUN_EXPERIMENTS=1")]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _pad(w):
    hop = 160
    n = (len(w) + hop - 1) // hop * hop
    return np.pad(w, (0, n - len(w))).astype(np.float32)


def test_folded_layernorm_encoder(tiny_engine, tiny_cfg, tiny_sd, monkeypatch):
    from oracle import nemo_restated as O
    from reazonspeech_b200.engine import Engine
    monkeypatch.setenv("RS_LN_FOLD", "1")
    folded = Engine(tiny_cfg, tiny_sd, "cuda:0")
    monkeypatch.delenv("RS_LN_FOLD")
    assert any(k.endswith(".fold") for k in folded.weights)
    waves = [_pad(synth_clip(90, 6.1)), _pad(synth_clip(91, 0.9)), _pad(synth_clip(92, 3.3))]
    L = max(len(w) for w in waves)
    x = torch.zeros(len(waves), L)
    lens = torch.tensor([len(w) for w in waves], dtype=torch.int32)
    for i, w in enumerate(waves):
        x[i, : len(w)] = torch.from_numpy(w)
    x, lens = x.cuda(), lens.cuda()
    mel, mel_len = tiny_engine.log_mel(x, lens)
    base, enc_len = tiny_engine.encode(mel, mel_len)
    got, enc_len2 = folded.encode(mel, mel_len)
    torch.cuda.synchronize()
    assert torch.equal(enc_len.cpu(), enc_len2.cpu())
    for i, w in enumerate(waves):
        T = int(enc_len[i])
        with torch.no_grad():
            ref = O.encoder(O.log_mel(torch.from_numpy(w), tiny_cfg), tiny_sd, tiny_cfg)
        r_base, r_fold, r_pair = _rel(base[i, :T].cpu(), ref), _rel(got[i, :T].cpu(), ref), _rel(got[i, :T].cpu(), base[i, :T].cpu())
        print(f"utt{i} T={T}: default vs oracle {r_base:.3e}, folded vs oracle {r_fold:.3e}, folded vs default {r_pair:.3e}")
        assert r_fold < 2e-2
        assert got[i, T:].abs().max().item() == 0.0 if T < got.shape[1] else True
