"""RS_LOGMEL_VARIANT=B (EXPERIMENT, written without GPU time left in round 1): the log-mel kernel with a register-resident
16 x 16 FFT, sixteen lanes per frame.  Its per-frame arithmetic is replayed on the CPU by tests/test_logmel_b_host.py;
this is the GPU half: the variant against the default kernel and against the oracle on a ragged batch.

    RS_RUN_EXPERIMENTS=1 timeout 300 python -m pytest tests/experiments/test_gpu_logmel_b.py -m gpu -q -s"""
import os

import numpy as np
import pytest
import torch

from reazonspeech_b200.synth import synth_clip

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("RS_RUN_EXPERIMENTS") != "1", reason="experiment: set RS_RUN_EXPERIMENTS=1")]


def test_logmel_variant_b(tiny_engine, tiny_cfg, tiny_sd):
    from oracle import nemo_restated as O
    from reazonspeech_b200.engine import Engine
    os.environ["RS_LOGMEL_VARIANT"] = "B"
    try:
        eng_b = Engine(tiny_cfg, tiny_sd, "cuda:0")
    finally:
        del os.environ["RS_LOGMEL_VARIANT"]
    assert "fe.b.tw_b" in eng_b.weights
    waves = [np.pad(synth_clip(130 + i, s), 8000).astype(np.float32) for i, s in enumerate((0.3, 2.0, 11.3, 5.05, 30.0))]
    L = max(len(w) for w in waves)
    x = torch.zeros(len(waves), L)
    for i, w in enumerate(waves):
        x[i, : len(w)] = torch.from_numpy(w)
    lens = torch.tensor([len(w) for w in waves], dtype=torch.int32)
    mel_a, len_a = tiny_engine.log_mel(x.cuda(), lens.cuda())
    mel_b, len_b = eng_b.log_mel(x.cuda(), lens.cuda())
    torch.cuda.synchronize()
    assert torch.equal(len_a, len_b)
    for i, w in enumerate(waves):
        n = int(len_a[i])
        with torch.no_grad():
            ref = O.log_mel(torch.from_numpy(w), tiny_cfg).T
        d_ab = (mel_a[i, :n] - mel_b[i, :n]).abs().max().item()
        d_bo = (mel_b[i, :n].cpu() - ref).abs().max().item()
        print(f"utt{i}: {n} frames; variant B vs default {d_ab:.2e}, variant B vs oracle {d_bo:.2e}")
        assert d_ab < 1e-3 and d_bo < 1e-3
        if n < mel_b.shape[1]:
            assert mel_b[i, n:].abs().max().item() == 0.0
