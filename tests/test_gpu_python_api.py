"""The drop-in Python surface on a GPU (pkg/nemo-asr/src/__init__.py:1-3 names + transcribe_batch): load_model ->
transcribe / transcribe_batch / the NeMo-shaped model.transcribe, on seeded synthetic weights of the tiny config.
Engine-level parity lives in test_gpu_kernels.py / test_gpu_full_model.py; here the host path around it is held to
the engine's own outputs: padding written in place, batching by length, pinned staging reuse, the worker-thread
pipeline and decode_hypothesis must not change a single token."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from reazonspeech_b200.synth import synth_clip


@pytest.fixture(scope="module")
def model(tiny_cfg):
    from reazonspeech_b200.nemo import asr
    return asr.load_model("cuda:0", synthetic=True, config=tiny_cfg, seed=0, max_batch=4)


def test_transcribe_batch_equals_transcribe_and_the_engine(model, tiny_cfg):
    from reazonspeech_b200.nemo import asr
    cfgv = asr.TranscribeConfig(verbose=False, raw_hypothesis=True)
    clips = [synth_clip(100 + i, s) for i, s in enumerate((2.0, 0.7, 3.1, 1.3, 2.6, 0.9, 4.2))]
    audios = [asr.audio_from_numpy(c, 16000) for c in clips]
    single = [asr.transcribe(model, a, cfgv) for a in audios]
    for rep in range(2):                                              # second pass reuses the (dirty) staging buffers
        batch = asr.transcribe_batch(model, audios, cfgv)
        assert len(batch) == len(audios)
        for i, (b, s) in enumerate(zip(batch, single)):
            assert b.hypothesis.y_sequence.tolist() == s.hypothesis.y_sequence.tolist(), f"clip {i}, pass {rep}"
            assert list(b.hypothesis.timestamp) == list(s.hypothesis.timestamp)
            assert b.text == s.text and b.subwords == s.subwords and b.segments == s.segments
    # the engine called directly on the reference's padded waveform (pad_audio: 0.5 s both sides)
    eng = model.engine
    n_tok = 0
    for i, c in enumerate(clips):
        w = np.pad(c.astype(np.float32), 8000)
        x = torch.from_numpy(w)[None].cuda()
        t, f, n = eng.transcribe_device(x, torch.tensor([len(w)], dtype=torch.int32).cuda())
        n = int(n[0])
        hyp = single[i].hypothesis
        assert hyp.y_sequence.tolist() == [tiny_cfg.blank] + t[0, :n].cpu().tolist()
        assert [ts - k - 1 for k, ts in enumerate(hyp.timestamp)] == f[0, :n].cpu().tolist()     # decode.py:48 recovers the frame
        n_tok += n
    assert n_tok > 0, "the synthetic checkpoint emitted nothing: the comparison above is vacuous"
    for r in single:
        assert all(s.seconds >= 0 for s in r.subwords)
        assert r.text == model.tokenizer.ids_to_text(r.hypothesis.y_sequence.tolist()[1:])


def test_nemo_call_shape(model, tiny_cfg):
    """model.transcribe(list_of_tensors, batch_size, return_hypotheses, verbose) as pkg/nemo-asr/src/transcribe.py:48-53 calls it."""
    w = torch.from_numpy(np.pad(synth_clip(120, 1.5), 8000).astype(np.float32))
    hyps = model.transcribe([w], batch_size=1, return_hypotheses=True, verbose=False)
    assert len(hyps) == 1 and int(hyps[0].y_sequence[0]) == tiny_cfg.blank
    assert len(hyps[0].timestamp) == len(hyps[0].y_sequence) - 1
    texts = model.transcribe([w, w[: len(w) // 2]], batch_size=2, return_hypotheses=False, verbose=False)
    assert len(texts) == 2 and all(isinstance(t, str) for t in texts)
