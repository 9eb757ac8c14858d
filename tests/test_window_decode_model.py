"""Host-side model of the windowed greedy loop of csrc/decode_spec.cu: scoring a window of K frames against
the current prediction-network state and consuming it up to the first non-blank must reproduce the
sequential greedy loop of the oracle (NeMo GreedyRNNTInfer) decision for decision, including the
max_symbols forced advance.  Pure CPU; the CUDA kernel is held to the oracle by the -m gpu tests."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nemo_restated as O
from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.weights import random_state_dict


def windowed_greedy(ep, sd, cfg, K):
    hp = cfg.pred_hidden
    emb = sd["decoder.prediction.embed.weight"]
    W, b = sd["joint.joint_net.2.weight"], sd["joint.joint_net.2.bias"]
    h, c = O.lstm_step(torch.zeros(hp), torch.zeros(hp), torch.zeros(hp), sd)
    pp = F.linear(h, sd["joint.pred.weight"], sd["joint.pred.bias"])
    t, sym, T = 0, 0, ep.shape[0]
    tokens, frames, iters = [], [], 0
    while t < T:
        iters += 1
        nv = min(K, T - t)
        ks = F.linear(torch.relu(ep[t:t + nv] + pp), W, b).argmax(-1).tolist()     # the whole window, one state
        for k in ks:
            if k == cfg.blank:
                t += 1; sym = 0
                continue
            tokens.append(k); frames.append(t)
            h, c = O.lstm_step(emb[k], h, c, sd)
            pp = F.linear(h, sd["joint.pred.weight"], sd["joint.pred.bias"])
            sym += 1
            if sym >= cfg.max_symbols:
                t += 1; sym = 0
            break
    return tokens, frames, iters


@pytest.mark.parametrize("max_symbols", [1, 2, 10])
@pytest.mark.parametrize("K", [1, 4, 7])
def test_windowed_loop_equals_sequential(K, max_symbols):
    cfg = ModelConfig.tiny().replace(max_symbols=max_symbols)
    sd = random_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(7)
    enc = torch.randn(90, cfg.d_model, generator=g) * 2.0
    with torch.no_grad():
        ref = O.rnnt_greedy(enc, sd, cfg)
        tok, fr, iters = windowed_greedy(O.joint_enc_proj(enc, sd), sd, cfg, K)
    assert len(ref.tokens) > 5
    assert tok == ref.tokens and fr == ref.frames
    if K > 1:                                       # blanks are what a window saves iterations on
        assert iters <= len(ref.decisions)
        if ref.decisions.count(cfg.blank) >= 8:
            assert iters < len(ref.decisions)
