"""Edge regimes of the greedy decode kernels that the default suite only reaches by chance: (1) the blank can never win, so
every frame emits exactly max_symbols tokens (the cap) and the output buffer fills to T x max_symbols;
(2) the blank always wins (no token at all); (3) U_max smaller than the true emission count (n_tok reports the true count,
only U_max entries are stored).  All four kernels (RS_DECODE_MODE 1-4) against the CPU oracle.  Written after the round's
GPU minutes were spent, hence parked here (RS_RUN_EXPERIMENTS=1) until it has been seen green once."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("RS_RUN_EXPERIMENTS") != "1", reason="experiment: set RS_RUN_EXPERIMENTS=1")]


def _engine(tiny_cfg, tiny_sd, blank_bias):
    from reazonspeech_b200.engine import Engine
    sd = dict(tiny_sd)
    b = sd["joint.joint_net.2.bias"].clone()
    b[tiny_cfg.blank] = blank_bias
    sd["joint.joint_net.2.bias"] = b
    return sd, Engine(tiny_cfg, sd, "cuda:0")


@pytest.mark.parametrize("mode", ["1", "2", "3", "4"])
def test_blank_never_wins_hits_the_symbol_cap(tiny_cfg, tiny_sd, mode, monkeypatch):
    from oracle import nemo_restated as O
    monkeypatch.setenv("RS_DECODE_MODE", mode)
    sd, eng = _engine(tiny_cfg, tiny_sd, -1e4)
    g = torch.Generator().manual_seed(3)
    lens = [9, 4, 13]
    T = max(lens)
    enc = torch.randn(len(lens), T, tiny_cfg.d_model, generator=g)
    tokens, frames, ntok = eng.greedy(enc.cuda(), torch.tensor(lens, dtype=torch.int32).cuda())
    torch.cuda.synchronize()
    for i, n_frames in enumerate(lens):
        ref = O.rnnt_greedy(enc[i, :n_frames], sd, tiny_cfg, emulate=True)
        n = int(ntok[i])
        assert n == n_frames * tiny_cfg.max_symbols == len(ref.tokens)
        assert frames[i, :n].cpu().tolist() == [t for t in range(n_frames) for _ in range(tiny_cfg.max_symbols)]
        got = tokens[i, :n].cpu().tolist()
        if got != ref.tokens:                       # tolerate a near-tie only: report the first divergence and its oracle margin
            k = next(j for j, (a, b) in enumerate(zip(got, ref.tokens)) if a != b)
            assert ref.margins[k] < 5e-2, f"utt {i}: token {k} differs at oracle margin {ref.margins[k]:.3e}"


@pytest.mark.parametrize("mode", ["1", "2", "3", "4"])
def test_blank_always_wins_and_small_output_buffer(tiny_cfg, tiny_sd, mode, monkeypatch):
    monkeypatch.setenv("RS_DECODE_MODE", mode)
    g = torch.Generator().manual_seed(4)
    enc = torch.randn(2, 11, tiny_cfg.d_model, generator=g)
    lens = torch.tensor([11, 6], dtype=torch.int32)
    _, eng = _engine(tiny_cfg, tiny_sd, 1e4)
    tokens, frames, ntok = eng.greedy(enc.cuda(), lens.cuda())
    torch.cuda.synchronize()
    assert ntok.cpu().tolist() == [0, 0]
    _, eng = _engine(tiny_cfg, tiny_sd, -1e4)
    tokens, frames, ntok = eng.greedy(enc.cuda(), lens.cuda(), U_max=7)      # room for 7 of 110 / 60 emissions
    torch.cuda.synchronize()
    assert ntok.cpu().tolist() == [11 * tiny_cfg.max_symbols, 6 * tiny_cfg.max_symbols]
    assert frames[0, :7].cpu().tolist() == [0] * 7 and tokens.shape[1] == 7
