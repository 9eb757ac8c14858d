"""Full-size (619 M, 24-layer) parity: short clips, then BASELINE.json configs[1] itself (32 x 30 s in one batch) against
the CPU oracle, then size-independent properties of the same batch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.synth import synth_clip


@pytest.fixture(scope="module")
def full():
    from reazonspeech_b200.engine import Engine
    from reazonspeech_b200.weights import random_state_dict
    cfg = ModelConfig()
    sd = random_state_dict(cfg, seed=0)
    return cfg, sd, Engine(cfg, sd, "cuda:0")


def _batch(waves):
    L = max(len(w) for w in waves)
    x = torch.zeros(len(waves), L)
    for i, w in enumerate(waves):
        x[i, : len(w)] = torch.from_numpy(w)
    return x.cuda(), torch.tensor([len(w) for w in waves], dtype=torch.int32).cuda()


def _oracle_check(cfg, sd, w, enc_row, enc_len_i, tok, frm, n, tag, with_fp32):
    """One clip: encoder relative L2 <= 2e-2 vs the fp32 oracle when ``with_fp32``; decision sequence vs the
    oracle with the bf16 storage points emulated, re-synchronising (tests/parity.py: no difference at a logit gap >= 1e-2,
    at most 3 near-ties)."""
    from oracle import nemo_restated as O
    from parity import check_decisions
    wt = torch.from_numpy(w)
    with torch.no_grad():
        mel = O.log_mel(wt, cfg)
        emu = O.encoder(mel, sd, cfg, emulate=True)
        T = emu.shape[0]
        assert enc_len_i == T == cfg.enc_frames(len(w))
        rel = None
        if with_fp32:
            ref = O.encoder(mel, sd, cfg)
            rel = ((enc_row[:T].double() - ref.double()).norm() / ref.double().norm()).item()
            assert rel < 2e-2, f"{tag}: encoder relative L2 {rel:.3e}"
    ties = check_decisions(tok[:n].tolist(), frm[:n].tolist(), emu, sd, cfg, tag)
    print(f"{tag}: T={T} enc rel-L2 {'-' if rel is None else format(rel, '.3e')}; {n} tokens; {ties} near-tie differences")
    return ties, rel


def test_full_model_encoder_and_tokens(full):
    """Short clips (one query tile): encoder and decision sequence against the oracle."""
    cfg, sd, eng = full
    waves = [np.pad(synth_clip(40, 2.5), 8000), np.pad(synth_clip(41, 4.0), 8000)]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    tokens, frames, ntok = [a.cpu() for a in eng.transcribe_device(x, lens)]
    enc = enc.cpu()
    for i, w in enumerate(waves):
        _oracle_check(cfg, sd, w, enc[i], int(enc_len[i]), tokens[i], frames[i], int(ntok[i]), f"utt{i}", True)


def test_production_geometry_parity(full):
    """BASELINE.json configs[1] itself -- the bench's own 32 x 30 s clip set in ONE batch (M = 32 x 392 rows, T = 388 valid
    frames = 4 query tiles of the tensor-core attention, +-128 window, global token, d = 1024 x 8 heads) -- against the CPU
    oracle (full 619 M model; about 1 s of CPU per clip and pass):
      * every clip: the engine's decision sequence walked through the oracle (bf16 storage points emulated) to the last
        frame; any difference at an oracle logit gap >= 1e-2 fails, more than 3 near-ties in a clip fail;
      * every fourth clip: encoder output relative L2 <= 2e-2 against the fp32 oracle;
    then a ragged batch (5 / 10 / 20 s next to 30 s clips) to the same bar."""
    cfg, sd, eng = full
    waves = [np.pad(synth_clip(i, 30.0), 8000) for i in range(32)]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    tokens, frames, ntok = [a.cpu() for a in eng.transcribe_device(x, lens)]
    enc = enc.cpu()
    assert enc.shape[1] == 392 and cfg.enc_frames(len(waves[0])) == 388
    ties, rels, failures = [], [], []

    def check(*a):          # every clip is examined before the test fails: one GPU run reports all of them
        try:
            t, r = _oracle_check(cfg, sd, *a)
        except AssertionError as exc:
            failures.append(str(exc)[:400]); print("FAIL", failures[-1])
            return
        ties.append(t)
        if r is not None:
            rels.append(r)

    for i, w in enumerate(waves):
        check(w, enc[i], int(enc_len[i]), tokens[i], frames[i], int(ntok[i]), f"clip{i}", i % 4 == 0)
    print(f"32 x 30 s: identical decision sequences {sum(t == 0 for t in ties)}/32, near-ties {ties}, worst encoder rel-L2 {max(rels) if rels else float('nan'):.3e}")
    waves = [np.pad(synth_clip(50 + i, s), 8000) for i, s in enumerate((5.0, 10.0, 20.0))] + [waves[3], waves[17]]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    t2, f2, n2 = [a.cpu() for a in eng.transcribe_device(x, lens)]
    enc = enc.cpu()
    for i, w in enumerate(waves[:3]):
        check(w, enc[i], int(enc_len[i]), t2[i], f2[i], int(n2[i]), f"ragged{i}", True)
    assert not failures, f"{len(failures)} clips fail parity: {failures[:3]}"
    for j, i in ((3, 3), (4, 17)):       # the 30 s clips decode identically next to shorter ones
        n = int(ntok[i])
        assert int(n2[j]) == n and torch.equal(t2[j, :n], tokens[i, :n]) and torch.equal(f2[j, :n], frames[i, :n])


def test_full_batch_properties(full):
    """BASELINE.json configs[1] shape (32 x 30 s): size-independent properties instead of the oracle:
    a clip decodes identically alone, inside the batch and at a different batch position; repeated runs
    are bit-identical; frames are non-decreasing, within range, at most max_symbols per frame."""
    cfg, sd, eng = full
    waves = [np.pad(synth_clip(i, 30.0), 8000) for i in range(8)] + [np.pad(synth_clip(50 + i, s), 8000) for i, s in enumerate((5.0, 10.0, 20.0))]
    x, lens = _batch(waves)
    t1, f1, n1 = [a.cpu() for a in eng.transcribe_device(x, lens)]
    t2, f2, n2 = [a.cpu() for a in eng.transcribe_device(x, lens)]
    assert torch.equal(t1, t2) and torch.equal(f1, f2) and torch.equal(n1, n2)
    perm = torch.arange(len(waves) - 1, -1, -1)
    t3, f3, n3 = [a.cpu() for a in eng.transcribe_device(x[perm.cuda()].contiguous(), lens[perm.cuda()].contiguous())]
    for i in range(len(waves)):
        j = int((perm == i).nonzero()[0])
        n = int(n1[i])
        assert int(n3[j]) == n and torch.equal(t1[i, :n], t3[j, :n]) and torch.equal(f1[i, :n], f3[j, :n]), f"utt {i} depends on batch position"
        fr = f1[i, :n]
        assert (fr[1:] >= fr[:-1]).all() and (n == 0 or int(fr.max()) < cfg.enc_frames(len(waves[i])))
        assert n == 0 or int(torch.bincount(fr).max()) <= cfg.max_symbols
    xs, ls = _batch([waves[9]])
    ta, fa, na = [a.cpu() for a in eng.transcribe_device(xs, ls)]
    n = int(na[0])
    assert n == int(n1[9]) and torch.equal(ta[0, :n], t1[9, :n])
    print("tokens per clip:", n1.tolist())
