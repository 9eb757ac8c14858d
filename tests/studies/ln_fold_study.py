"""CPU study (not a test): how much encoder accuracy would folding the three single LayerNorms of a layer into their
consumer GEMMs cost?  y = r (bf16(x) @ bf16(W * gamma)^T) - r mu c + d, per-row (mu, r) in fp32.
Measured (4 s clip, seeded weights, relative L2 against the fp32 oracle; today = bf16 storage points emulated):
    2 x 256: 3.6e-3 -> 3.8e-3     6 x 512: 2.8e-3 -> 3.9e-3     12 x 1024: 2.7e-3 -> 5.1e-3   (tolerance 2e-2)
Usage: python tests/studies/ln_fold_study.py"""
import sys, math, numpy as np, torch, torch.nn.functional as F
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nemo_restated as O
from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.weights import random_state_dict
from reazonspeech_b200.synth import synth_clip

def q(x): return x.to(torch.bfloat16).float()

def folded_linear(x, ln_w, ln_b, W, b, eps):
    """y = LN(x) @ W^T + b evaluated as the folded GEMM: bf16(x) @ bf16(W*gamma)^T with per-row (mu, r) applied after."""
    mu = x.mean(-1, keepdim=True); var = x.var(-1, unbiased=False, keepdim=True); r = torch.rsqrt(var + eps)
    Wp = q(W * ln_w[None, :])
    c = Wp.sum(1)                        # consistent with the rounded weights
    d = W @ ln_b + (b if b is not None else 0)
    return r * (q(x) @ Wp.T) - r * mu * c[None, :] + d[None, :]

def layer_folded(x, sd, i, cfg):
    p = f"encoder.layers.{i}."
    dd = (cfg.d_model,)
    ln = lambda t, n: F.layer_norm(t, dd, sd[p+n+".weight"], sd[p+n+".bias"], cfg.ln_eps)
    em = True
    # FFN1 (its LN is the chained norm_out+ff1 kernel: stays a real LN)
    x = x + 0.5 * O.feed_forward(q(ln(x, "norm_feed_forward1")), sd, p+"feed_forward1.", em)
    # attention: QKV projection folded
    a = p + "self_attn."
    T, d = x.shape; H, dk = cfg.n_heads, cfg.d_head
    Wqkv = torch.cat([sd[a+"linear_q.weight"], sd[a+"linear_k.weight"], sd[a+"linear_v.weight"]], 0)
    bqkv = torch.cat([sd[a+"linear_q.bias"], sd[a+"linear_k.bias"], sd[a+"linear_v.bias"]], 0)
    qkv = q(folded_linear(x, sd[p+"norm_self_att.weight"], sd[p+"norm_self_att.bias"], Wqkv, bqkv, cfg.ln_eps))
    heads = lambda t: t.view(T, H, dk).transpose(0, 1)
    pos = F.linear(O.rel_pos_table(cfg), sd[a+"linear_pos.weight"])
    pp = q(pos).view(cfg.n_rel, H, dk).transpose(0, 1)
    o = O.local_attention_core(heads(qkv[:, :d]), heads(qkv[:, d:2*d]), heads(qkv[:, 2*d:]), pp, sd[a+"pos_bias_u"], sd[a+"pos_bias_v"], cfg, em)
    o = q(o.transpose(0, 1).reshape(T, d))
    x = x + F.linear(o, sd[a+"linear_out.weight"], sd[a+"linear_out.bias"])
    # conv module: pointwise_conv1 folded
    c = p + "conv."
    y = folded_linear(x, sd[p+"norm_conv.weight"], sd[p+"norm_conv.bias"], sd[c+"pointwise_conv1.weight"][:, :, 0], sd[c+"pointwise_conv1.bias"], cfg.ln_eps)
    y = q(F.glu(y, dim=-1)).T.unsqueeze(0)
    pad = (cfg.conv_kernel - 1) // 2
    y = F.conv1d(y, sd[c+"depthwise_conv.weight"], sd[c+"depthwise_conv.bias"], padding=pad, groups=y.shape[1])
    y = F.batch_norm(y, sd[c+"batch_norm.running_mean"], sd[c+"batch_norm.running_var"], sd[c+"batch_norm.weight"], sd[c+"batch_norm.bias"], training=False, eps=cfg.bn_eps)
    y = q(F.silu(y)[0].T)
    x = x + F.linear(y, sd[c+"pointwise_conv2.weight"][:, :, 0], sd[c+"pointwise_conv2.bias"])
    # FFN2: linear1 folded
    f = p + "feed_forward2."
    h = F.silu(folded_linear(x, sd[p+"norm_feed_forward2.weight"], sd[p+"norm_feed_forward2.bias"], sd[f+"linear1.weight"], sd[f+"linear1.bias"], cfg.ln_eps))
    x = x + 0.5 * F.linear(q(h), sd[f+"linear2.weight"], sd[f+"linear2.bias"])
    return ln(x, "norm_out")

for name, cfg in (("tiny", ModelConfig.tiny()), ("6x512", ModelConfig(n_layers=6, d_model=512, n_heads=4, sub_channels=128, vocab_size=127, pred_hidden=128, joint_hidden=128)),
                  ("12x1024", ModelConfig(n_layers=12, vocab_size=127, pred_hidden=128, joint_hidden=128))):
    sd = random_state_dict(cfg, seed=0, calibrate=False)
    w = torch.from_numpy(np.pad(synth_clip(7, 4.0), 8000))
    with torch.no_grad():
        mel = O.log_mel(w, cfg)
        ref = O.encoder(mel, sd, cfg)                         # fp32
        emu = O.encoder(mel, sd, cfg, emulate=True)           # today's storage points
        x = O.subsample(mel, sd, cfg, True) * math.sqrt(cfg.d_model)
        for i in range(cfg.n_layers): x = layer_folded(x, sd, i, cfg)
    rel = lambda a: float((a - ref).norm() / ref.norm())
    rowmean = None
    print(f"{name}: today (emulated) rel-L2 {rel(emu):.3e}   LN folded into consumer GEMMs rel-L2 {rel(x):.3e}")
