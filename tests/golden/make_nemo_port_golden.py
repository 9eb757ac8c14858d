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
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
