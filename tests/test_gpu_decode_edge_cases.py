"""Edge regimes of the greedy decode kernel that ordinary clips only reach by chance (the ``max_symbols`` cap
is part of the greedy loop's contract): (1) the blank can never win, so every frame emits exactly max_symbols tokens and the
output fills to T x max_symbols; (2) the blank always wins (no token at all); (3) U_max smaller than the true emission
count (n_tok reports the true count, only U_max entries are stored); (4) random encoder outputs at ragged batch sizes
(zero-length utterances, batches that do not divide into the kernel's utterance groups and span several passes) against
the CPU oracle's sequential loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(tiny_cfg, tiny_sd, blank_bias):
    from reazonspeech_b200.engine import Engine
    sd = dict(tiny_sd)
    b = sd["joint.joint_net.2.bias"].clone()
    b[tiny_cfg.blank] = blank_bias
    sd["joint.joint_net.2.bias"] = b
    return sd, Engine(tiny_cfg, sd, "cuda:0")


def test_blank_never_wins_hits_the_symbol_cap(tiny_cfg, tiny_sd):
    from oracle import nemo_restated as O
    from parity import check_decisions
    sd, eng = _engine(tiny_cfg, tiny_sd, -1e4)
    g = torch.Generator().manual_seed(3)
    lens = [9, 4, 13]
    T = max(lens)
    enc = torch.randn(len(lens), T, tiny_cfg.d_model, generator=g)
    tokens, frames, ntok = eng.greedy(enc.cuda(), torch.tensor(lens, dtype=torch.int32).cuda())
    torch.cuda.synchronize()
    for i, n_frames in enumerate(lens):
        n = int(ntok[i])
        assert n == n_frames * tiny_cfg.max_symbols
        assert frames[i, :n].cpu().tolist() == [t for t in range(n_frames) for _ in range(tiny_cfg.max_symbols)]
        # the encoder output is shared with the oracle: identical decisions up to exact near-ties (gap < 1e-3 in fp32 logits)
        check_decisions(tokens[i, :n].cpu().tolist(), frames[i, :n].cpu().tolist(), enc[i, :n_frames], sd, tiny_cfg, f"utt{i}", tol=1e-3)


def test_blank_always_wins_and_small_output_buffer(tiny_cfg, tiny_sd):
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


@pytest.mark.parametrize("B", [1, 3, 5, 33, 70, 130])
def test_windowed_decode_equals_the_sequential_loop(tiny_engine, tiny_cfg, tiny_sd, B):
    """The windowed kernel (4 frames per iteration, utterance groups x vocabulary slices) against the oracle's one-decision-
    at-a-time loop on random encoder outputs, ragged lengths (incl. a zero-length utterance), batch sizes that are not
    multiples of the group count and span several passes: decision sequences identical."""
    from parity import check_decisions
    eng = tiny_engine
    g = torch.Generator().manual_seed(B)
    T = 61
    enc = torch.randn(B, T, tiny_cfg.d_model, generator=g) * 2.0
    enc_len = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32)
    enc_len[0] = T
    if B > 2:
        enc_len[2] = 0
    t, f, n = [a.cpu() for a in eng.greedy(enc.cuda(), enc_len.cuda())]
    print("tokens per utterance:", n.tolist()[:12], "lens", enc_len.tolist()[:12])
    ties = 0
    for b in range(B):
        k, L = int(n[b]), int(enc_len[b])
        if L == 0:
            assert k == 0
            continue
        ties += check_decisions(t[b, :k].tolist(), f[b, :k].tolist(), enc[b, :L], tiny_sd, tiny_cfg, f"utt{b}", tol=1e-3, verbose=False)
    assert ties <= max(1, B // 10)          # exact fp32 near-ties (gap < 1e-3) are the only admissible differences
