"""CPU study (not a test): how much does the one semantic of the global attention token that no implementation in the image
can confirm matter?  The oracle and the kernels keep global key 0 ALSO inside the local band of the queries near it
(recalled from NeMo's RelPositionMultiHeadAttentionLongformer, where the band mask is the padding mask only); Hugging
Face's Longformer, which that class was adapted from, removes global positions from the band.  This script evaluates
both on the same seeded model and prints the relative L2 between the two encoder outputs -- the size of the error the
engine would carry IF NeMo followed the Hugging Face variant.

Measured (seeded random weights):   2 x 256, w = 16, 6 s: 2.0e-2 (4.5e-2 on the frames within the window of token 0)
                                    6 x 512, w = 128, 14 s: 7.2e-3 -- the same order as the bf16 storage noise (3e-3), well inside the 2e-2 tolerance
Usage: python tests/studies/global_token_band_study.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nemo_restated as O  # noqa: E402
from reazonspeech_b200.config import ModelConfig  # noqa: E402
from reazonspeech_b200.synth import synth_clip  # noqa: E402
from reazonspeech_b200.weights import random_state_dict  # noqa: E402


def attention_without_global_in_band(q, k, v, p, u, vb, cfg, emulate=False):
    """local_attention_core with the global keys masked out of the local band (the Hugging Face Longformer variant)."""
    import math
    H, T, dk = q.shape
    wl, wr, G = cfg.att_left, cfg.att_right, cfg.global_tokens
    scale = 1.0 / math.sqrt(dk)
    ac = torch.matmul(q + u[:, None, :], k.transpose(1, 2))
    bd_rel = torch.matmul(q + vb[:, None, :], p.transpose(1, 2))
    i = torch.arange(T)[:, None]; j = torch.arange(T)[None, :]
    rel = j - i
    band = (rel >= -wl) & (rel <= wr) & (j >= G)                                   # <- the only difference
    bd = torch.gather(bd_rel, 2, (rel + wl).clamp(0, cfg.n_rel - 1).unsqueeze(0).expand(H, T, T))
    s_local = ((ac + bd) * scale).masked_fill(~band.unsqueeze(0), float("-inf"))
    s_glob = torch.matmul(q * scale, k[:, :G].transpose(1, 2))
    probs = torch.softmax(torch.cat((s_glob, s_local), dim=-1), dim=-1)
    out = torch.matmul(probs[..., :G], v[:, :G]) + torch.matmul(probs[..., G:], v)
    sg = torch.matmul(q[:, :G] * scale, k.transpose(1, 2))
    out[:, :G] = torch.matmul(torch.softmax(sg, dim=-1), v)
    return out


def run(cfg: ModelConfig, seconds: float):
    sd = random_state_dict(cfg, seed=0, calibrate=False)
    wave = torch.from_numpy(np.pad(synth_clip(7, seconds), 8000).astype(np.float32))
    with torch.no_grad():
        mel = O.log_mel(wave, cfg)
        a = O.encoder(mel, sd, cfg)
        keep = O.local_attention_core
        O.local_attention_core = attention_without_global_in_band
        try:
            b = O.encoder(mel, sd, cfg)
        finally:
            O.local_attention_core = keep
    T = a.shape[0]
    near = min(T, cfg.att_left + 1)
    rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())
    print(f"{cfg.n_layers} x {cfg.d_model}, w = {cfg.att_left}, T = {T}: whole output rel-L2 {rel(a, b):.3e}; "
          f"frames within the window of token 0 {rel(a[:near], b[:near]):.3e}; frames beyond {rel(a[near:], b[near:]) if T > near else 0.0:.3e}")


if __name__ == "__main__":
    run(ModelConfig.tiny(), 6.0)
    run(ModelConfig(n_layers=6, d_model=512, n_heads=4, sub_channels=128, vocab_size=127, pred_hidden=128, joint_hidden=128), 14.0)
