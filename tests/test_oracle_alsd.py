"""CPU checks of the ALSD beam-search restatement (oracle/alsd_restated.py; SURVEY.md section 8(f).3 -- the reference's
default decoding, pkg/nemo-asr/src/decode.py:29,38-40,48).  NeMo is absent offline, so these are internal consistency
properties plus the Hypothesis shape the reference's own decode.py consumes."""
import math

import numpy as np
import pytest
import torch

from oracle import nemo_restated as O
from oracle.alsd_restated import alsd_beam
from reazonspeech_b200.synth import synth_clip


@pytest.fixture(scope="module")
def enc_case(tiny_cfg, tiny_sd):
    w = torch.from_numpy(np.pad(synth_clip(3, 3.0), 8000))
    with torch.no_grad():
        return O.encoder(O.log_mel(w, tiny_cfg), tiny_sd, tiny_cfg)


def _path_score(enc, sd, cfg, tokens, frames):
    """log-probability of ONE alignment (the tokens at their frames, blanks elsewhere) under the transducer."""
    import torch.nn.functional as F
    ep = O.joint_enc_proj(enc, sd)
    emb = sd["decoder.prediction.embed.weight"]
    h = torch.zeros(cfg.pred_hidden); c = torch.zeros(cfg.pred_hidden)
    h2, c2 = O.lstm_step(torch.zeros(cfg.pred_hidden), h, c, sd)
    pp = F.linear(h2, sd["joint.pred.weight"], sd["joint.pred.bias"])
    W, b = sd["joint.joint_net.2.weight"], sd["joint.joint_net.2.bias"]
    score, i = 0.0, 0
    for t in range(ep.shape[0]):
        while True:
            logp = torch.log_softmax(F.linear(torch.relu(ep[t] + pp), W, b), dim=-1)
            if i < len(tokens) and frames[i] == t:
                score += float(logp[tokens[i]])
                h, c = h2, c2
                h2, c2 = O.lstm_step(emb[tokens[i]], h, c, sd)
                pp = F.linear(h2, sd["joint.pred.weight"], sd["joint.pred.bias"])
                i += 1
            else:
                score += float(logp[cfg.blank])
                break
    assert i == len(tokens)
    return score


def test_beam_one_is_greedy_without_the_symbol_cap(enc_case, tiny_cfg, tiny_sd):
    g = O.rnnt_greedy(enc_case, tiny_sd, tiny_cfg)
    counts = np.bincount(g.frames, minlength=enc_case.shape[0])
    capped = np.nonzero(counts >= tiny_cfg.max_symbols)[0]
    T = int(capped[0]) if len(capped) else enc_case.shape[0]          # greedy's max_symbols cap has no counterpart in ALSD
    assert T > 8
    g = O.rnnt_greedy(enc_case[:T], tiny_sd, tiny_cfg)
    r = alsd_beam(enc_case[:T], tiny_sd, tiny_cfg, beam=1)
    assert len(g.tokens) > 5 and r.tokens == g.tokens and r.frames == g.frames


def test_hypothesis_shape_and_score(enc_case, tiny_cfg, tiny_sd):
    r = alsd_beam(enc_case, tiny_sd, tiny_cfg, beam=4, recombine_returns_input=False)
    T = enc_case.shape[0]
    assert r.y_sequence[0] == tiny_cfg.blank and len(r.timestamp) == len(r.y_sequence) - 1      # decode.py:38-40 / pack_hypotheses
    assert all(0 <= f < T for f in r.frames) and r.frames == sorted(r.frames)
    assert r.timestamp == [f + i for i, f in enumerate(r.frames)]                                # alignment step = t + u
    # without recombination gains the reported score is the log-probability of that very alignment
    best_unmerged = alsd_beam(enc_case, tiny_sd, tiny_cfg, beam=4, recombine_returns_input=True)
    s = _path_score(enc_case, tiny_sd, tiny_cfg, best_unmerged.tokens, best_unmerged.frames)
    assert s <= best_unmerged.score + 1e-3                             # logaddexp of equal sequences can only add mass
    # a wider beam never ends with a worse length-normalised score than greedy's own path
    g = O.rnnt_greedy(enc_case, tiny_sd, tiny_cfg)
    gs = _path_score(enc_case, tiny_sd, tiny_cfg, g.tokens, g.frames) / (len(g.tokens) + 1)
    wide = alsd_beam(enc_case, tiny_sd, tiny_cfg, beam=8)
    assert wide.score / len(wide.y_sequence) >= gs - 1e-4 or wide.tokens != g.tokens


def test_reference_decode_py_consumes_the_hypothesis(enc_case, tiny_cfg, tiny_sd):
    """The ALSD result goes through our decode_hypothesis (byte-identical to the reference's decode.py on the same input,
    tests/test_decode_golden.py): seconds = max(0.08 (step - idx - 1) - 0.5, 0) with step = t + idx, i.e. frame t - 1."""
    from reazonspeech_b200.nemo.asr.decode import decode_hypothesis
    from reazonspeech_b200.tokenizer import PieceTableTokenizer, synthetic_pieces

    class M:
        tokenizer = PieceTableTokenizer(synthetic_pieces(tiny_cfg.vocab_size))

    class H:
        pass

    r = alsd_beam(enc_case, tiny_sd, tiny_cfg, beam=4)
    h = H(); h.y_sequence = torch.tensor(r.y_sequence); h.timestamp = r.timestamp
    res = decode_hypothesis(M, h)
    assert len(res.subwords) <= len(r.tokens)
    frames = {sw.token_id: None for sw in res.subwords}
    for sw in res.subwords:
        assert sw.seconds >= 0
    first = res.subwords[0]
    idx = r.tokens.index(first.token_id)
    assert math.isclose(first.seconds, max(0.08 * (r.frames[idx] - 1) - 0.5, 0), abs_tol=1e-9)
