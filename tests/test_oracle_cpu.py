"""CPU checks of the oracle restatement itself (no GPU): shape arithmetic, independent
constructions of the same quantity, and the properties the engine relies on."""
import math

import numpy as np
import torch

from oracle import nemo_restated as O
from reazonspeech_b200.config import ModelConfig, conv_out_len
from reazonspeech_b200.engine import glu_interleave_index, pack_weights
from reazonspeech_b200.synth import synth_clip
from reazonspeech_b200.weights import mel_filterbank, random_state_dict, rel_pos_table, state_dict_shapes


def test_shape_arithmetic_matches_survey():
    c = ModelConfig()
    assert c.mel_frames(496000) == 3101 and c.mel_valid(496000) == 3100 and c.enc_frames(496000) == 388
    assert [conv_out_len(n) for n in (3101, 1551, 776)] == [1551, 776, 388]
    assert c.sub_freq == 10 and c.sub_out_dim == 2560 and c.n_rel == 257


def test_param_count_is_619m():
    n = sum(int(np.prod(s)) for k, s in state_dict_shapes(ModelConfig()).items() if "running" not in k)
    assert 600e6 < n < 640e6, n                                                # README.rst:34-35: 619 M


def test_mel_filterbank_matches_torchaudio(tiny_cfg):
    w = torch.from_numpy(np.pad(synth_clip(0, 1.0), 8000))
    a, b = O.log_mel(w, tiny_cfg), O.log_mel(w, tiny_cfg, independent_fb=True)
    assert (a - b).abs().max() < 1e-3
    fb = mel_filterbank(tiny_cfg)
    assert fb.shape == (80, 257) and (fb >= 0).all() and (fb.sum(1) > 0).all()


def test_log_mel_statistics(tiny_cfg):
    m = O.log_mel(torch.from_numpy(np.pad(synth_clip(1, 2.0), 8000)), tiny_cfg)
    assert m.shape == (80, tiny_cfg.mel_valid(48000)) and tiny_cfg.mel_valid(48000) == tiny_cfg.mel_frames(48000) - 1
    assert m.mean(1).abs().max() < 1e-4 and (m.std(1) - 1).abs().max() < 1e-3


def test_stft_reflect_equals_constant_on_padded_audio(tiny_cfg):
    """with transcribe()'s 0.5 s silence the STFT pad mode is immaterial."""
    x = torch.from_numpy(np.pad(synth_clip(2, 1.0), 8000))
    from reazonspeech_b200.weights import hann_window
    kw = dict(n_fft=512, hop_length=160, win_length=400, window=hann_window(tiny_cfg), center=True, return_complex=True)
    y = torch.cat((x[:1], x[1:] - 0.97 * x[:-1]))
    assert torch.equal(torch.stft(y, pad_mode="constant", **kw), torch.stft(y, pad_mode="reflect", **kw))


def test_local_attention_matches_bruteforce(tiny_cfg):
    """Dense gather formulation vs an explicit per-query loop over the band + global column."""
    cfg = tiny_cfg
    g = torch.Generator().manual_seed(0)
    H, T, dk = cfg.n_heads, 45, cfg.d_head
    q, k, v = (torch.randn(H, T, dk, generator=g) for _ in range(3))
    p = torch.randn(H, cfg.n_rel, dk, generator=g)
    u, vb = torch.randn(H, dk, generator=g) * 0.1, torch.randn(H, dk, generator=g) * 0.1
    got = O.local_attention_core(q, k, v, p, u, vb, cfg)
    s = 1 / math.sqrt(dk)
    for h in range(H):
        for i in range(T):
            if i == 0:
                sc = (q[h, 0] * s) @ k[h].T
                ref = torch.softmax(sc, 0) @ v[h]
            else:
                js = [j for j in range(T) if -cfg.att_left <= j - i <= cfg.att_right]
                sc = [(q[h, i] * s) @ k[h, 0]] + [((q[h, i] + u[h]) @ k[h, j] + (q[h, i] + vb[h]) @ p[h, cfg.att_left - (i - j)]) * s for j in js]
                pr = torch.softmax(torch.stack(sc), 0)
                ref = pr[0] * v[h, 0] + sum(pr[1 + n] * v[h, j] for n, j in enumerate(js))
            assert (got[h, i] - ref).abs().max() < 1e-4


def test_rel_pos_table_orientation(tiny_cfg):
    t = rel_pos_table(tiny_cfg)
    assert t.shape == (tiny_cfg.n_rel, tiny_cfg.d_model)
    assert torch.allclose(t[tiny_cfg.att_left, 0::2], torch.zeros(tiny_cfg.d_model // 2))   # centre row = position 0
    assert t[0, 0] == torch.sin(torch.tensor(float(tiny_cfg.att_left)))                       # row 0 = +left


def test_greedy_respects_max_symbols(tiny_cfg, tiny_sd):
    sd = dict(tiny_sd)
    b = sd["joint.joint_net.2.bias"].clone(); b[tiny_cfg.blank] = -1e4                       # blank never wins
    sd["joint.joint_net.2.bias"] = b
    enc = torch.randn(7, tiny_cfg.d_model, generator=torch.Generator().manual_seed(1))
    r = O.rnnt_greedy(enc, sd, tiny_cfg)
    assert len(r.tokens) == 7 * tiny_cfg.max_symbols
    assert r.frames == [t for t in range(7) for _ in range(tiny_cfg.max_symbols)]


def test_synthetic_weights_are_bf16_exact_and_seeded(tiny_cfg, tiny_sd):
    again = random_state_dict(tiny_cfg, seed=0)
    for k, v in tiny_sd.items():
        assert torch.equal(v, again[k])
        assert torch.equal(v, v.to(torch.bfloat16).float()), k
    assert torch.count_nonzero(tiny_sd["decoder.prediction.embed.weight"][tiny_cfg.blank]) == 0


def test_emission_rate_is_bounded(tiny_cfg, tiny_sd):
    """The calibrated blank bias keeps greedy decoding away from the max_symbols-per-frame regime."""
    n_tok = n_frames = 0
    for i, s in ((3, 8.0), (6, 4.0), (8, 6.0)):
        r = O.transcribe_tokens(torch.from_numpy(np.pad(synth_clip(i, s), 8000)), tiny_sd, tiny_cfg)
        n_tok += len(r.tokens); n_frames += tiny_cfg.enc_frames(int(s * 16000) + 16000)
    assert 0.02 * n_frames < n_tok < 2 * n_frames, (n_tok, n_frames)


def test_pack_weights_layouts(tiny_cfg, tiny_sd):
    p = pack_weights(tiny_sd, tiny_cfg)
    d = tiny_cfg.d_model
    idx = glu_interleave_index(d)
    assert sorted(idx.tolist()) == list(range(2 * d))
    assert idx[:16].tolist() == list(range(16)) and idx[16:32].tolist() == list(range(d, d + 16))
    # sub.out.w column permutation: engine column f*C + c  <->  NeMo column c*F3 + f
    W = tiny_sd["encoder.pre_encode.out.weight"]; C, F3 = tiny_cfg.sub_channels, tiny_cfg.sub_freq
    Wp = p["sub.out.w"].float()
    assert torch.equal(Wp[:, 3 * C + 5], W[:, 5 * F3 + 3])
    # BatchNorm folding reproduces conv + BN on a random input
    pre = "encoder.layers.0.conv."
    x = torch.randn(1, d, 20)
    y = torch.nn.functional.conv1d(x, tiny_sd[pre + "depthwise_conv.weight"], tiny_sd[pre + "depthwise_conv.bias"], padding=4, groups=d)
    y = torch.nn.functional.batch_norm(y, tiny_sd[pre + "batch_norm.running_mean"], tiny_sd[pre + "batch_norm.running_var"],
                                       tiny_sd[pre + "batch_norm.weight"], tiny_sd[pre + "batch_norm.bias"], False, eps=tiny_cfg.bn_eps)
    w = p["L0.conv.dw.w"]                                                                    # [k, d]
    y2 = torch.nn.functional.conv1d(x, w.T.unsqueeze(1).contiguous(), p["L0.conv.dw.shift"], padding=4, groups=d)
    assert (y - y2).abs().max() < 1e-4
    assert p["pred.lstm.w"].shape == (4 * tiny_cfg.pred_hidden, 2 * tiny_cfg.pred_hidden)


class _ModuleRnnt(torch.nn.Module):
    """The prediction and joint networks built from stock torch.nn modules under the CHECKPOINT's parameter names
    (decoder.prediction.embed / .dec_rnn.lstm, joint.enc / .pred / .joint_net = [ReLU, Dropout, Linear]),
    so that the state dict loads with strict=True.  torch.nn.LSTM is the very module NeMo wraps; this pins the oracle's
    hand-written cell, gate order and bias handling to the library implementation and its key layout."""

    def __init__(self, cfg):
        super().__init__()
        pred = torch.nn.Module()
        pred.embed = torch.nn.Embedding(cfg.n_classes, cfg.pred_hidden, padding_idx=cfg.blank)
        rnn = torch.nn.Module()
        rnn.lstm = torch.nn.LSTM(cfg.pred_hidden, cfg.pred_hidden, num_layers=1)
        pred.dec_rnn = rnn
        self.decoder = torch.nn.Module()
        self.decoder.prediction = pred
        self.joint = torch.nn.Module()
        self.joint.enc = torch.nn.Linear(cfg.d_model, cfg.joint_hidden)
        self.joint.pred = torch.nn.Linear(cfg.pred_hidden, cfg.joint_hidden)
        self.joint.joint_net = torch.nn.Sequential(torch.nn.ReLU(), torch.nn.Dropout(0.2), torch.nn.Linear(cfg.joint_hidden, cfg.n_classes))


def test_greedy_matches_stock_module_implementation(tiny_cfg, tiny_sd):
    cfg = tiny_cfg
    m = _ModuleRnnt(cfg).eval()
    m.load_state_dict({k: v for k, v in tiny_sd.items() if k.startswith(("decoder.", "joint."))}, strict=True)
    enc = 2.0 * torch.randn(60, cfg.d_model, generator=torch.Generator().manual_seed(5))
    ref = O.rnnt_greedy(enc, tiny_sd, cfg)
    assert 0 < len(ref.tokens) < 60 * cfg.max_symbols
    with torch.no_grad():
        f = m.joint.enc(enc)
        tokens, frames, state, last = [], [], None, cfg.blank            # SOS = blank, whose embedding row is zero
        for t in range(enc.shape[0]):
            for _ in range(cfg.max_symbols):
                g, new_state = m.decoder.prediction.dec_rnn.lstm(m.decoder.prediction.embed(torch.tensor([[last]])), state)
                logits = m.joint.joint_net(f[t] + m.joint.pred(g[0, 0]))
                k = int(logits.argmax())
                if k == cfg.blank:
                    break
                tokens.append(k); frames.append(t); state, last = new_state, k
    assert tokens == ref.tokens and frames == ref.frames


def test_greedy_follow_resynchronises(tiny_cfg, tiny_sd):
    """oracle.greedy_follow (the re-synchronising comparison used by every token-identity test, tests/parity.py): the
    oracle's own decisions give no difference; a flipped decision is reported with the amount it loses by and the walk
    goes on teacher-forced to the last frame; truncated / over-long sequences are flagged; a decision sequence rebuilt
    from tokens + frames equals the recorded one."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from parity import check_decisions, decisions_from
    import pytest
    cfg, sd = tiny_cfg, tiny_sd
    w = torch.from_numpy(np.pad(synth_clip(3, 2.5), 8000))
    with torch.no_grad():
        enc = O.encoder(O.log_mel(w, cfg), sd, cfg)
    ref = O.rnnt_greedy(enc, sd, cfg)
    assert len(ref.tokens) > 3
    assert decisions_from(ref.tokens, ref.frames, enc.shape[0], cfg.max_symbols, cfg.blank) == ref.decisions
    r = O.greedy_follow(enc, sd, cfg, ref.decisions)
    assert r.complete and not r.gaps and r.n_decisions == len(ref.decisions)
    assert abs(r.min_margin - min(ref.margins)) < 1e-6
    assert check_decisions(ref.tokens, ref.frames, enc, sd, cfg, "self", emulate=False) == 0
    # flip one emission to another token: reported once, with a positive gap, and the walk continues to the end
    j = next(i for i, k in enumerate(ref.decisions) if k != cfg.blank)
    bad = list(ref.decisions)
    bad[j] = (bad[j] + 1) % cfg.vocab_size
    r = O.greedy_follow(enc, sd, cfg, bad)
    assert r.complete and r.gaps and r.gaps[0][0] == j and r.gaps[0][2] == bad[j] and r.gaps[0][3] == ref.decisions[j] and r.gaps[0][4] > 0
    # a wrong token changes the predictor state, later oracle decisions may differ too -- but every one is examined
    assert r.n_decisions == len(bad)
    tok = [k for k in bad if k != cfg.blank]
    with pytest.raises(AssertionError):
        check_decisions(tok, ref.frames, enc, sd, cfg, "flipped", emulate=False, tol=0.0)
    assert not O.greedy_follow(enc, sd, cfg, ref.decisions[:-3]).complete
    assert not O.greedy_follow(enc, sd, cfg, ref.decisions + [cfg.blank]).complete
