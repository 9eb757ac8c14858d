"""Pin the oracle against an INDEPENDENT implementation of NeMo's FastConformer encoder.

NeMo itself (``nemo_toolkit[asr] >= 2.6.1``, pkg/nemo-asr/pyproject.toml:13) cannot be imported
offline, but the image ships ``transformers.models.parakeet``: NVIDIA's own port of NeMo's
``FilterbankFeatures`` + ``ConformerEncoder`` (dw_striding subsampling, relative-position MHSA,
convolution module, macaron FFN), validated upstream against NeMo checkpoints.  This script

  1. maps a seeded NeMo-named state dict (``reazonspeech_b200.weights.random_state_dict``) onto a
     ``ParakeetEncoder`` of the same shape,
  2. runs ``ParakeetFeatureExtractor`` + ``ParakeetEncoder`` (batch = 1, like the reference calls
     NeMo: pkg/nemo-asr/src/transcribe.py:48-50) on seeded clips,
  3. stores inputs' seeds and the third-party outputs in ``parakeet_cases.npz``.

``tests/test_oracle_parakeet.py`` checks ``oracle/nemo_restated.py`` against the stored vectors
(always) and against a live Parakeet run (when transformers is importable).  Parakeet implements
full relative-position attention (``self_attention_model: rel_pos``); the oracle's Longformer-style
local attention with window +-w and no global token is the same function whenever T <= w + 1,
which is how the cases are chosen.  What stays (R) after this pin: the global-token wiring of
``RelPositionMultiHeadAttentionLongformer`` and the greedy loop's ``max_symbols``.

The feature extractor builds its mel matrix with librosa, which is absent here; the matrix is
injected from ``torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")`` (the same
Slaney construction) -- every other step of the frontend is Parakeet's own code.

Usage:  python tests/golden/make_parakeet_golden.py   (writes tests/golden/parakeet_cases.npz)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from reazonspeech_b200.config import ModelConfig  # noqa: E402
from reazonspeech_b200.synth import synth_clip  # noqa: E402
from reazonspeech_b200.weights import random_state_dict  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parakeet_cases.npz")

# (name, config, weight seed, clip seed, seconds).  T = enc_frames must be <= att_left + 1.
CASES = [
    ("tiny_2s", dict(n_layers=2, d_model=256, n_heads=2, sub_channels=64, att_left=128, att_right=128,
                     global_tokens=0, vocab_size=127, pred_hidden=128, joint_hidden=128), 3, 40, 2.0),
    ("tiny_7s", dict(n_layers=2, d_model=256, n_heads=2, sub_channels=64, att_left=128, att_right=128,
                     global_tokens=0, vocab_size=127, pred_hidden=128, joint_hidden=128), 4, 41, 7.3),
    ("mid_4s", dict(n_layers=3, d_model=512, n_heads=4, sub_channels=128, att_left=128, att_right=128,
                    global_tokens=0, vocab_size=127, pred_hidden=128, joint_hidden=128), 5, 42, 4.1),
]


def case_config(kw: dict) -> ModelConfig:
    return ModelConfig(**kw)


def parakeet_feature_extractor(cfg: ModelConfig):
    """ParakeetFeatureExtractor without its librosa import (see module docstring)."""
    import torchaudio
    from transformers.models.parakeet.feature_extraction_parakeet import ParakeetFeatureExtractor
    from transformers.feature_extraction_sequence_utils import SequenceFeatureExtractor

    fe = ParakeetFeatureExtractor.__new__(ParakeetFeatureExtractor)
    SequenceFeatureExtractor.__init__(fe, feature_size=cfg.n_mels, sampling_rate=cfg.sample_rate, padding_value=0.0)
    fe.hop_length, fe.n_fft, fe.win_length, fe.preemphasis = cfg.n_window_stride, cfg.n_fft, cfg.n_window_size, cfg.preemph
    fe.mel_filters = torchaudio.functional.melscale_fbanks(
        cfg.n_freq, 0.0, cfg.sample_rate / 2, cfg.n_mels, cfg.sample_rate, norm="slaney", mel_scale="slaney").T.contiguous()
    return fe


def parakeet_encoder(cfg: ModelConfig, sd: dict):
    """A ParakeetEncoder carrying the NeMo-named weights of ``sd``."""
    from transformers.models.parakeet.configuration_parakeet import ParakeetEncoderConfig
    from transformers.models.parakeet.modeling_parakeet import ParakeetEncoder

    pc = ParakeetEncoderConfig(
        hidden_size=cfg.d_model, num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_heads,
        intermediate_size=cfg.d_ff, conv_kernel_size=cfg.conv_kernel, subsampling_factor=cfg.sub_factor,
        subsampling_conv_channels=cfg.sub_channels, num_mel_bins=cfg.n_mels, scale_input=cfg.xscaling,
        dropout=0.0, dropout_positions=0.0, layerdrop=0.0, activation_dropout=0.0, attention_dropout=0.0)
    pc._attn_implementation = "eager"
    enc = ParakeetEncoder(pc).eval()
    hf = {}
    pre = "encoder.pre_encode."
    for i in (0, 2, 3, 5, 6):
        hf[f"subsampling.layers.{i}.weight"] = sd[pre + f"conv.{i}.weight"]
        hf[f"subsampling.layers.{i}.bias"] = sd[pre + f"conv.{i}.bias"]
    hf["subsampling.linear.weight"] = sd[pre + "out.weight"]
    hf["subsampling.linear.bias"] = sd[pre + "out.bias"]
    ren = {"self_attn.linear_q": "self_attn.q_proj", "self_attn.linear_k": "self_attn.k_proj",
           "self_attn.linear_v": "self_attn.v_proj", "self_attn.linear_out": "self_attn.o_proj",
           "self_attn.linear_pos": "self_attn.relative_k_proj", "self_attn.pos_bias_u": "self_attn.bias_u",
           "self_attn.pos_bias_v": "self_attn.bias_v", "conv.batch_norm": "conv.norm"}
    for k, v in sd.items():
        if not k.startswith("encoder.layers."):
            continue
        name = k[len("encoder."):]
        for a, b in ren.items():
            name = name.replace(a, b)
        hf[name] = v
    own = enc.state_dict()
    missing = [k for k in own if k not in hf and not k.endswith("num_batches_tracked")]
    extra = [k for k in hf if k not in own]
    assert not missing and not extra, (missing, extra)
    enc.load_state_dict(hf, strict=False)
    return enc


def run_parakeet(cfg: ModelConfig, sd: dict, wave: np.ndarray):
    """-> (features [F, n_mels] as the extractor returns them, feature length, encoder output [T, d])."""
    fe = parakeet_feature_extractor(cfg)
    enc = parakeet_encoder(cfg, sd)
    with torch.no_grad():
        feats = fe(wave, sampling_rate=cfg.sample_rate, return_tensors="pt")
        out = enc(input_features=feats["input_features"], attention_mask=feats["attention_mask"])
    n = int(feats["attention_mask"].sum())
    t = int(out.attention_mask.sum())
    return feats["input_features"][0].numpy(), n, out.last_hidden_state[0, :t].numpy()


def padded_clip(seed: int, seconds: float) -> np.ndarray:
    """What transcribe() hands to NeMo: the clip with 0.5 s of silence on both sides (audio.py:70-83)."""
    return np.pad(synth_clip(seed, seconds), 8000).astype(np.float32)


def main() -> None:
    blob = {}
    for name, kw, wseed, cseed, secs in CASES:
        cfg = case_config(kw)
        sd = random_state_dict(cfg, seed=wseed, calibrate=False)
        wave = padded_clip(cseed, secs)
        feats, n, enc = run_parakeet(cfg, sd, wave)
        assert enc.shape[0] <= cfg.att_left + 1
        blob[name + ".features"] = feats.astype(np.float32)
        blob[name + ".n_frames"] = np.int64(n)
        blob[name + ".encoder"] = enc.astype(np.float32)
        print(name, "features", feats.shape, "valid", n, "encoder", enc.shape)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
