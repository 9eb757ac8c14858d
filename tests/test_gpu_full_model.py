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


def _oracle_check(cfg, sd, w, enc_row, enc_len_i, tok, frm, n, tag):
    """One clip of the whole path against the oracle: encoder relative L2 <= 2e-2 vs the fp32 oracle, and the
    decision sequence walked through the oracle (bf16 storage points emulated) to the last frame under the noise-aware bar
    of tests/parity.py: every difference within 4 sigma of the storage noise measured in the oracle itself for this clip,
    no more differences than 3 + 3 x what that noise is expected to overturn."""
    from oracle import nemo_restated as O
    from parity import check_decisions_noise_aware
    with torch.no_grad():
        mel = O.log_mel(torch.from_numpy(w), cfg)
        emu = O.encoder(mel, sd, cfg, emulate=True)
        ref = O.encoder(mel, sd, cfg)
    T = emu.shape[0]
    assert enc_len_i == T == cfg.enc_frames(len(w))
    rel = ((enc_row[:T].double() - ref.double()).norm() / ref.double().norm()).item()
    frame_rel = ((enc_row[:T].double() - ref.double()).norm(dim=1) / ref.double().norm(dim=1)).max().item()
    assert rel < 2e-2 and frame_rel < 2e-2, f"{tag}: encoder relative L2 {rel:.3e} (worst single frame {frame_rel:.3e})"
    r = check_decisions_noise_aware(tok[:n].tolist(), frm[:n].tolist(), emu, ref, sd, cfg, tag)
    print(f"{tag}: T={T} enc rel-L2 {rel:.3e} (worst frame {frame_rel:.3e}); {n} tokens, {r['decisions']} decisions: {r['differences']} differ "
          f"(storage noise sigma {r['sigma']:.2e} -> {r['expected']:.1f} expected, largest gap {r['max_gap']:.2e})")
    return r, rel, emu


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
        _oracle_check(cfg, sd, w, enc[i], int(enc_len[i]), tokens[i], frames[i], int(ntok[i]), f"utt{i}")


def test_production_geometry_parity(full):
    """BASELINE.json configs[1] itself -- the bench's own 32 x 30 s clip set in ONE batch (M = 32 x 392 rows, T = 388 valid
    frames = 4 query tiles of the tensor-core attention, +-128 window, global token, d = 1024 x 8 heads) -- against the CPU
    oracle (full 619 M model, about 1 s of CPU per clip and encoder pass), every clip, nothing sampled:
      * encoder output: relative L2 <= 2e-2 against the fp32 oracle, over the clip AND for its worst single frame (the
        first and last frames of an utterance are where windows clip, tiles end and pad rows begin);
      * the DECODE KERNEL ALONE, fed the oracle's encoder output: decision sequence identical to the oracle's, all 32 clips
        in one batch (the encoder output is shared, so only an exact fp32 near-tie -- a logit gap below 1e-3, at most one
        per clip -- may come out the other way);
      * the whole path: decision sequence walked through the oracle to the last frame, noise-aware bar (tests/parity.py);
    then a ragged batch (5 / 10 / 20 s next to 30 s clips) to the same bars."""
    from parity import check_decisions
    cfg, sd, eng = full
    waves = [np.pad(synth_clip(i, 30.0), 8000) for i in range(32)]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    tokens, frames, ntok = [a.cpu() for a in eng.transcribe_device(x, lens)]
    enc = enc.cpu()
    assert enc.shape[1] == 392 and cfg.enc_frames(len(waves[0])) == 388
    results, rels, emus, failures = [], [], [], []

    def check(*a):          # every clip is examined before the test fails: one GPU run reports all of them
        try:
            r, rel, emu = _oracle_check(cfg, sd, *a)
        except AssertionError as exc:
            failures.append(str(exc)[:400]); print("FAIL", failures[-1])
            return None
        results.append(r); rels.append(rel)
        return emu

    for i, w in enumerate(waves):
        emus.append(check(w, enc[i], int(enc_len[i]), tokens[i], frames[i], int(ntok[i]), f"clip{i}"))
    same = sum(r["differences"] == 0 for r in results)
    print(f"32 x 30 s whole path: identical decision sequences {same}/32; differing decisions {sum(r['differences'] for r in results)} of "
          f"{sum(r['decisions'] for r in results)} (storage noise predicts {sum(r['expected'] for r in results):.0f}); "
          f"worst encoder rel-L2 {max(rels) if rels else float('nan'):.3e}")
    # the decode kernel alone on the oracle's encoder outputs: identical, not merely close
    ok = [i for i, e in enumerate(emus) if e is not None]
    eb = torch.zeros(len(ok), 392, cfg.d_model)
    for j, i in enumerate(ok):
        eb[j, :388] = emus[i]
    tk, fr, nt = [a.cpu() for a in eng.greedy(eb.cuda(), torch.full((len(ok),), 388, dtype=torch.int32).cuda())]
    for j, i in enumerate(ok):
        n = int(nt[j])
        try:
            check_decisions(tk[j, :n].tolist(), fr[j, :n].tolist(), emus[i], sd, cfg, f"decode-alone clip{i}", tol=1e-3, max_near_ties=1)
        except AssertionError as exc:
            failures.append(str(exc)[:400]); print("FAIL", failures[-1])
    print(f"decode kernel alone on the oracle's encoder output: {len(ok)} clips, identical decision sequences required (fp32 near-ties < 1e-3 aside)")
    waves = [np.pad(synth_clip(50 + i, s), 8000) for i, s in enumerate((5.0, 10.0, 20.0))] + [waves[3], waves[17]]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    t2, f2, n2 = [a.cpu() for a in eng.transcribe_device(x, lens)]
    enc = enc.cpu()
    for i, w in enumerate(waves[:3]):
        check(w, enc[i], int(enc_len[i]), t2[i], f2[i], int(n2[i]), f"ragged{i}")
    assert not failures, f"{len(failures)} checks fail parity: {failures[:3]}"
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


def test_alsd_full_size_matches_the_oracle():
    """ALSD beam search at the production size (V = 3000, 640-wide predictor / joint, tripled-weight GEMMs with K = 1920 /
    3840): the winner equals the oracle's on the same encoder output."""
    from oracle.alsd_restated import alsd_beam
    from reazonspeech_b200.engine import Engine
    from reazonspeech_b200.weights import random_state_dict
    cfg = ModelConfig()
    sd = random_state_dict(cfg, seed=0)
    eng = Engine(cfg, sd, "cuda:0", alsd=True)
    waves = [np.pad(synth_clip(60 + i, s), 8000) for i, s in enumerate((4.0, 6.5, 2.0))]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    y, steps, n, score = [a.cpu() for a in eng.alsd(enc, enc_len, beam=4)]
    enc = enc.cpu()
    same = 0
    for i in range(len(waves)):
        T = int(enc_len[i])
        ref = alsd_beam(enc[i, :T], sd, cfg, beam=4, emulate=True)
        k = int(n[i])
        ok = y[i, : k + 1].tolist() == ref.y_sequence and steps[i, :k].tolist() == ref.timestamp
        print(f"utt{i}: T={T}, {k} tokens (oracle {len(ref.tokens)}), score {float(score[i]):.3f} (oracle {ref.score:.3f}), identical={ok}")
        same += ok
        assert abs(float(score[i]) / (k + 1) - ref.score / len(ref.y_sequence)) < 1e-3
    assert same >= len(waves) - 1


def test_long_form_clip_at_full_size(full):
    """SURVEY.md section 8(f).2 at the production size: a 150 s clip (1 888 encoder frames = 15 query tiles, ~95 k mel frames) in
    ONE call next to a short one, as the reference feeds audio of any length to the model (transcribe.py:44-53): encoder within
    2e-2 of the fp32 oracle over the clip and for its worst frame, decisions under the noise-aware bar of tests/parity.py."""
    cfg, sd, eng = full
    waves = [np.pad(synth_clip(95, 150.0), 8000), np.pad(synth_clip(96, 3.0), 8000)]
    x, lens = _batch(waves)
    mel, mel_len = eng.log_mel(x, lens)
    enc, enc_len = eng.encode(mel, mel_len)
    tokens, frames, ntok = [a.cpu() for a in eng.transcribe_device(x, lens)]
    enc = enc.cpu()
    assert int(enc_len[0]) == cfg.enc_frames(len(waves[0])) == 1888
    for i, w in enumerate(waves):
        _oracle_check(cfg, sd, w, enc[i], int(enc_len[i]), tokens[i], frames[i], int(ntok[i]), f"long{i}")
