"""Host-side contracts between engine.py::pack_weights and the attention kernels, checked on CPU against the oracle's
own formulas (oracle/nemo_restated.py::local_attention_core):

* pos_bias_u is folded into the q bias of the fused QKV projection, and the positional GEMM's bias takes it out again:
  q' k = (q + u) k  and  q' p[c] + bdbias[c] = (q + v) p[c];
* the rel_shift of the positional term on the read side of the attention kernel's shared-memory BD tile."""
import torch

from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.engine import pack_weights
from reazonspeech_b200.weights import random_state_dict, rel_pos_table


def test_pos_bias_u_fold_is_algebraically_neutral():
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, seed=1)
    pk = pack_weights(sd, cfg)
    H, dk, d = cfg.n_heads, cfg.d_head, cfg.d_model
    a = "encoder.layers.0.self_attn."
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, d, generator=g)
    w, b = pk["L0.att.wqkv"].float(), pk["L0.att.bqkv"]
    qkv = x @ w.T + b
    qp = qkv[:, :d].view(-1, H, dk)                                # what the engine's QKV buffer holds in its q columns
    k = qkv[:, d:2 * d].view(-1, H, dk)
    q = (x @ sd[a + "linear_q.weight"].T + sd[a + "linear_q.bias"]).view(-1, H, dk)
    u, v = sd[a + "pos_bias_u"], sd[a + "pos_bias_v"]
    # content term: (q + u) . k
    ac_engine = torch.einsum("ihd,jhd->hij", qp, k)
    ac_oracle = torch.einsum("ihd,jhd->hij", q + u, k)
    assert (ac_engine - ac_oracle).abs().max() < 1e-3
    # positional term: (q + v) . p[c], with p = linear_pos(table) as the oracle builds it
    pos = torch.nn.functional.linear(rel_pos_table(cfg), sd[a + "linear_pos.weight"]).view(cfg.n_rel, H, dk)
    bd_oracle = torch.einsum("ihd,chd->ihc", q + v, pos)
    p_packed = pk["L0.att.pos"].float()[:, : cfg.n_rel]            # [H, n_rel, dk] (bf16-rounded table)
    bias = pk["L0.att.bdbias"].view(H, -1)[:, : cfg.n_rel]
    bd_engine = torch.einsum("ihd,hcd->ihc", qp, p_packed) + bias
    assert (bd_engine - bd_oracle).abs().max() < 5e-2 * bd_oracle.abs().max()     # table rounded to bf16
    # global token: scored against q, not q + u
    assert ((qp - u) - q).abs().max() < 1e-4


def test_rel_shift_read_index_contract():
    """attention_tc.cu keeps the positional product BD[r][c] (row r of a 128-query tile, relative offset index c) un-shifted in
    shared memory and applies NeMo's rel_shift on the read side: for window column jj (key j = q0 - 128 + jj) row r reads
    BD[r][jj - (128 - w_left) - r].  That must be the oracle's index (j - i) + w_left for every in-band (i, j), and the 17
    32-bit words a chunk loads must cover its 32 columns for either parity of the start."""
    for w_left, w_right in ((128, 128), (16, 16), (8, 24), (64, 32)):
        for q0 in (0, 128, 384):
            for r in (0, 1, 77, 127):
                i = q0 + r
                for rel in range(-w_left, w_right + 1):
                    j = i + rel
                    if j < 0:
                        continue
                    jj = j - (q0 - 128)
                    assert 0 <= jj < 384
                    m, e = jj // 32, jj % 32                       # chunk and element, as pass 1 walks them
                    cs = 32 * m - (128 - w_left) - r                # first BD column of the chunk for this row
                    assert cs + e == rel + w_left                   # the oracle's `idx`
                    word, sh = cs >> 1, (cs & 1) * 16               # floor division also for negative starts
                    pair = e >> 1                                   # funnelshift(w[pair], w[pair + 1], sh) holds elements e, e + 1 of the chunk
                    first_half = 2 * (word + pair) + (1 if sh else 0)
                    assert first_half == cs + 2 * pair and pair + 1 <= 16


def test_gate_table_is_the_input_half_of_the_lstm_step():
    """decode_spec.cu adds pred.gate_tab[token] (built once by pack_weights) to W_hh . h instead of multiplying W_ih with the
    embedding at every step: one LSTM step through the table equals the oracle's lstm_step, for ordinary tokens and for the
    start-of-sequence step (blank row = the zero padding embedding -> the bias alone)."""
    from oracle import nemo_restated as O
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, seed=3)
    pk = pack_weights(sd, cfg)
    hp = cfg.pred_hidden
    tab = pk["pred.gate_tab"]
    assert tab.shape == (cfg.vocab_size + 1, 4 * hp) and tab.dtype == torch.float32
    l = "decoder.prediction.dec_rnn.lstm."
    w_hh = sd[l + "weight_hh_l0"]
    emb = sd["decoder.prediction.embed.weight"]
    g = torch.Generator().manual_seed(0)
    for k in (0, 5, cfg.vocab_size - 1, cfg.blank):
        h, c = torch.randn(hp, generator=g) * 0.5, torch.randn(hp, generator=g) * 0.5
        gates = tab[k] + w_hh @ h
        i, f, gg, o = gates.split(hp)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        x = torch.zeros(hp) if k == cfg.blank else emb[k]          # the oracle's SOS step feeds zeros (blank as padding)
        h_ref, c_ref = O.lstm_step(x, h, c, sd)
        assert (h2 - h_ref).abs().max() < 1e-5 and (c2 - c_ref).abs().max() < 1e-5
