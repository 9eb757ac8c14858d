"""Second pin of the oracle: NeMo's ConformerEncoder as ported (nearly verbatim, with NeMo's own parameter names)
into vLLM for its Cohere ASR model -- ``vllm.model_executor.models.cohere_asr`` (vllm 0.22 in this image):
ConvSubsampling with MaskedConvSequential, RelPositionalEncoding + rel_shift, ConformerLayer, and the
``att_context_size`` band mask of ``ConformerEncoder._create_masks``.

What this adds to the Parakeet pin (make_parakeet_golden.py):

  * the seeded NeMo-named state dict loads into the port with ``strict=True`` -- the oracle's key names, tensor shapes and
    layouts (conv.N indices, linear_pos, pos_bias_u/v, batch_norm buffers) are the checkpoint's, not a guess;
  * LIMITED attention context with T far beyond the window: the port evaluates full relative-position attention under a
    band mask (att_context_style "regular"), which is the same function as NeMo's Longformer-style
    ``rel_pos_local_attn`` without a global token (same sinusoid per relative offset, same positions excluded);
    Parakeet only covered T <= w + 1;
  * a PADDED, RAGGED BATCH: the port is run on a zero-padded batch with a length vector (pad masks in attention, the
    masking between the subsampling convolutions, masked_fill before the depthwise convolution), and every utterance's
    valid frames must equal the oracle's batch-of-one result -- the padding-invariance contract the engine is built on.

  * the FRONTEND as well: the same vLLM package carries a copy of NeMo's ``FilterbankFeatures`` class
    (``vllm.transformers_utils.processors.cohere_asr``: get_seq_len, preemphasis with the time mask, constant-padded centred
    STFT, Slaney mel, log(x + 2^-24), per-feature normalisation with N-1, masking); a ragged batch goes through it and on
    into the encoder port, so one case is NeMo's own classes end to end against the oracle's batch-of-one path.

Still not covered by any third-party implementation in the image: the global token of
RelPositionMultiHeadAttentionLongformer and the RNN-T greedy loop.

The features fed to the port are the oracle's log-mel (pinned by the Parakeet vectors), laid out as NeMo's preprocessor
hands them over: [B, n_mels, L // hop + 1] with frames at and beyond ``length = L // hop`` zero.

Usage:  python tests/golden/make_nemo_port_golden.py     (writes tests/golden/nemo_port_cases.npz; needs vllm importable)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import nemo_restated as O  # noqa: E402
from reazonspeech_b200.config import ModelConfig  # noqa: E402
from reazonspeech_b200.synth import synth_clip  # noqa: E402
from reazonspeech_b200.weights import random_state_dict  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nemo_port_cases.npz")
PAD = 8000

# (name, config overrides on ModelConfig.tiny(), weight seed, [(clip seed, seconds)] = one ragged batch)
CASES = [
    ("w16_batch3", dict(att_left=16, att_right=16, global_tokens=0), 7, [(60, 6.3), (61, 1.1), (62, 11.7)]),
    ("w8_24_batch2", dict(att_left=8, att_right=24, global_tokens=0), 8, [(63, 9.4), (64, 4.0)]),
]


def case_config(kw: dict) -> ModelConfig:
    return ModelConfig.tiny().replace(**kw)


def padded_clip(seed: int, seconds: float) -> np.ndarray:
    return np.pad(synth_clip(seed, seconds), PAD).astype(np.float32)      # the reference's pad_audio (0.5 s both sides)


def port_encoder(cfg: ModelConfig, sd: dict):
    """vLLM's port of NeMo's ConformerEncoder carrying the NeMo-named weights of ``sd`` (strict load)."""
    from vllm.model_executor.models.cohere_asr import ConformerEncoder
    enc_cfg = dict(feat_in=cfg.n_mels, n_layers=cfg.n_layers, d_model=cfg.d_model, feat_out=-1, causal_downsampling=False,
                   subsampling="dw_striding", subsampling_factor=cfg.sub_factor, subsampling_conv_channels=cfg.sub_channels,
                   ff_expansion_factor=cfg.ff_expansion, self_attention_model="rel_pos", n_heads=cfg.n_heads,
                   att_context_size=[cfg.att_left, cfg.att_right], att_context_style="regular", xscaling=cfg.xscaling,
                   untie_biases=True, pos_emb_max_len=5000, conv_kernel_size=cfg.conv_kernel, conv_norm_type="batch_norm",
                   conv_context_size=None)
    vc = types.SimpleNamespace(model_config=types.SimpleNamespace(hf_config=types.SimpleNamespace(encoder=enc_cfg)))
    enc = ConformerEncoder(vllm_config=vc).eval()
    own = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    for k in enc.state_dict():
        if k.endswith("num_batches_tracked"):
            own[k] = torch.tensor(0)
    enc.load_state_dict(own, strict=True)
    return enc


def port_frontend(cfg: ModelConfig):
    """vLLM's copy of NeMo's FilterbankFeatures, configured as model.transcribe runs it (dither 0, pad_to 0)."""
    from vllm.transformers_utils.processors.cohere_asr import FilterbankFeatures
    import torch._dynamo
    torch._dynamo.config.disable = True          # the port's forward is decorated with torch.compile; run it eagerly (same arithmetic)
    fe = FilterbankFeatures(sample_rate=cfg.sample_rate, n_window_size=cfg.n_window_size, n_window_stride=cfg.n_window_stride,
                            window="hann", normalize="per_feature", n_fft=cfg.n_fft, preemph=cfg.preemph, nfilt=cfg.n_mels,
                            lowfreq=0, highfreq=None, log=True, log_zero_guard_type="add", log_zero_guard_value=cfg.log_zero_guard,
                            dither=0.0, pad_to=0, frame_splicing=1, mag_power=2.0, mel_norm="slaney").eval()
    # The vLLM copy ends its constructor by rounding its two buffers to bfloat16 (a choice tied to Cohere's checkpoint);
    # NeMo keeps them in fp32.  Re-create both with the class's own expressions, minus that cast.
    from torchaudio.functional import melscale_fbanks
    assert fe.window.dtype == torch.bfloat16 and fe.fb.dtype == torch.bfloat16, "the port changed: re-read its constructor"
    fe.window = torch.hann_window(fe.win_length, periodic=False)
    fe.fb = melscale_fbanks(n_freqs=fe.n_fft // 2 + 1, f_min=0, f_max=cfg.sample_rate / 2, n_mels=cfg.n_mels,
                            sample_rate=cfg.sample_rate, norm="slaney", mel_scale="slaney").T.unsqueeze(0)
    return fe


FRONTEND_CLIPS = [(70, 1.3), (71, 3.9), (72, 0.6)]        # one ragged batch through the frontend port (and on into the encoder port)


def run_frontend(cfg: ModelConfig, sd: dict):
    """Ragged zero-padded batch -> FilterbankFeatures port -> ConformerEncoder port: NeMo's own classes end to end."""
    waves = [padded_clip(s, sec) for s, sec in FRONTEND_CLIPS]
    L = max(len(w) for w in waves)
    x = torch.zeros(len(waves), L)
    for i, w in enumerate(waves):
        x[i, : len(w)] = torch.from_numpy(w)
    lens = torch.tensor([len(w) for w in waves], dtype=torch.float)
    with torch.no_grad():
        feats, flen = port_frontend(cfg)(x, lens)                                      # [B, n_mels, F_max], [B]
        enc, elen = port_encoder(cfg, sd)(audio_signal=feats, length=flen)
    return waves, feats, flen.tolist(), [enc[i, :, : int(elen[i])].T.contiguous().numpy() for i in range(len(waves))]


def run_case(kw: dict, wseed: int, clips):
    cfg = case_config(kw)
    sd = random_state_dict(cfg, seed=wseed, calibrate=False)
    enc = port_encoder(cfg, sd)
    waves = [padded_clip(s, sec) for s, sec in clips]
    mels = [O.log_mel(torch.from_numpy(w), cfg) for w in waves]                        # [n_mels, F_valid] each
    lengths = torch.tensor([m.shape[1] for m in mels], dtype=torch.int64)
    F_max = max(len(w) // cfg.n_window_stride + 1 for w in waves)
    feats = torch.zeros(len(waves), cfg.n_mels, F_max)
    for i, m in enumerate(mels):
        feats[i, :, : m.shape[1]] = m
    with torch.no_grad():
        out, out_len = enc(audio_signal=feats, length=lengths)                         # [B, d, T], [B]
    return cfg, sd, waves, [out[i, :, : int(out_len[i])].T.contiguous().numpy() for i in range(len(waves))], out_len.tolist()


def main():
    store = {}
    for name, kw, wseed, clips in CASES:
        cfg, sd, waves, outs, out_len = run_case(kw, wseed, clips)
        store[f"{name}/n"] = np.int64(len(clips))
        for i, o in enumerate(outs):
            store[f"{name}/enc{i}"] = o.astype(np.float32)
            with torch.no_grad():
                ref = O.encoder(O.log_mel(torch.from_numpy(waves[i]), cfg), sd, cfg)
            rel = float((torch.from_numpy(o).double() - ref.double()).norm() / ref.double().norm())
            print(f"{name} utt{i}: T={o.shape[0]} (window {cfg.att_left}+{cfg.att_right}+1), oracle vs port rel-L2 {rel:.3e}")
    # frontend port (+ encoder port on its features), first case's config and weights
    name, kw, wseed, _ = CASES[0]
    cfg = case_config(kw)
    sd = random_state_dict(cfg, seed=wseed, calibrate=False)
    waves, feats, flen, encs = run_frontend(cfg, sd)
    for i, w in enumerate(waves):
        n = int(flen[i])
        assert float(feats[i, :, n:].abs().max()) == 0.0 if n < feats.shape[2] else True
        store[f"frontend/mel{i}"] = feats[i, :, :n].T.contiguous().numpy().astype(np.float32)      # [F_valid, n_mels]
        store[f"frontend/enc{i}"] = encs[i].astype(np.float32)
        with torch.no_grad():
            mel = O.log_mel(torch.from_numpy(w), cfg)
            ref = O.encoder(mel, sd, cfg)
        err = float((mel.T - feats[i, :, :n].T).abs().max())
        rel = float((torch.from_numpy(encs[i]).double() - ref.double()).norm() / ref.double().norm())
        print(f"frontend utt{i}: {n} frames (oracle {mel.shape[1]}), log-mel max-abs {err:.3e}; encoder on port features vs oracle end to end rel-L2 {rel:.3e}")
    store["frontend/n"] = np.int64(len(waves))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
