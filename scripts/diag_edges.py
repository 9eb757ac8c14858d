"""GPU diagnostic (not a test): where do the engine's decisions leave the oracle's?  Round-2 production-geometry parity
showed differences only in the first ~25 and the last ~4 encoder frames of a clip.  For a few clips this prints, stage by
stage, the engine's per-frame error against the fp32 oracle at the head, the middle and the tail of the utterance, and
replays the decode kernel alone on the ORACLE's encoder output.   gpurun -- python scripts/diag_edges.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import nemo_restated as O
from oracle.cpu_threads import physical_threads
from parity import decisions_from
from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.engine import Engine
from reazonspeech_b200.synth import synth_clip
from reazonspeech_b200.weights import random_state_dict

print("threads", physical_threads())
cfg = ModelConfig()
sd = random_state_dict(cfg, seed=0)
eng = Engine(cfg, sd, "cuda:0")
clips = {"clip5": np.pad(synth_clip(5, 30.0), 8000), "clip15": np.pad(synth_clip(15, 30.0), 8000), "short40": np.pad(synth_clip(40, 2.5), 8000)}
names = list(clips)
L = max(len(w) for w in clips.values())
x = torch.zeros(len(clips), L)
for i, w in enumerate(clips.values()):
    x[i, : len(w)] = torch.from_numpy(w)
lens = torch.tensor([len(w) for w in clips.values()], dtype=torch.int32)
xd, ld = x.cuda(), lens.cuda()
mel, mel_len = eng.log_mel(xd, ld)


def fmt(v):
    return " ".join(f"{a:.1e}" for a in v)


def frame_err(got, ref):
    return ((got.double() - ref.double()).norm(dim=1) / ref.double().norm(dim=1).clamp_min(1e-20)).tolist()


with torch.no_grad():
    for i, (name, w) in enumerate(clips.items()):
        wt = torch.from_numpy(w)
        mo = O.log_mel(wt, cfg).T
        F = mo.shape[0]
        d = (mel[i, :F].cpu() - mo).abs().max(dim=1).values
        print(f"{name}: log-mel max-abs per frame head {fmt(d[:6].tolist())} | mid max {float(d[60:F-60].max()) if F > 200 else float(d.max()):.1e} | tail {fmt(d[-6:].tolist())}")
    for nl in (0, 1, 2, 4, 8, 24):
        enc, enc_len = eng.encode(mel, mel_len, n_layers=nl)
        enc = enc.cpu()
        for i, (name, w) in enumerate(clips.items()):
            wt = torch.from_numpy(w)
            ref = O.encoder(O.log_mel(wt, cfg), sd, cfg, n_layers=nl)
            emu = O.encoder(O.log_mel(wt, cfg), sd, cfg, emulate=True, n_layers=nl)
            T = ref.shape[0]
            e = frame_err(enc[i, :T], ref)
            ee = frame_err(emu, ref)
            mid = e[30:T - 10] if T > 60 else e
            print(f"layers={nl:2d} {name}: T={T} engine-vs-fp32 head {fmt(e[:10])} | mid mean {np.mean(mid):.1e} max {np.max(mid):.1e} | tail {fmt(e[-6:])}   [emulated oracle: head {fmt(ee[:3])} tail {fmt(ee[-3:])}]")
    # decode kernel alone on the oracle's (emulated) encoder output
    encs = [O.encoder(O.log_mel(torch.from_numpy(w), cfg), sd, cfg, emulate=True) for w in clips.values()]
    Tm = max(e.shape[0] for e in encs)
    Tcap = (Tm + 7) // 8 * 8
    eb = torch.zeros(len(encs), Tcap, cfg.d_model)
    for i, e in enumerate(encs):
        eb[i, : e.shape[0]] = e
    el = torch.tensor([e.shape[0] for e in encs], dtype=torch.int32)
    tok, frm, ntok = [a.cpu() for a in eng.greedy(eb.cuda(), el.cuda())]
    for i, name in enumerate(names):
        n = int(ntok[i])
        got = decisions_from(tok[i, :n].tolist(), frm[i, :n].tolist(), encs[i].shape[0], cfg.max_symbols, cfg.blank)
        r = O.greedy_follow(encs[i], sd, cfg, got, emulate=True)
        print(f"decode alone {name}: {n} tokens, complete={r.complete}, differences {[(g[1], round(g[4], 4)) for g in r.gaps]}")
    # whole path, for reference
    tok, frm, ntok = [a.cpu() for a in eng.transcribe_device(xd, ld)]
    for i, name in enumerate(names):
        n = int(ntok[i])
        got = decisions_from(tok[i, :n].tolist(), frm[i, :n].tolist(), encs[i].shape[0], cfg.max_symbols, cfg.blank)
        r = O.greedy_follow(encs[i], sd, cfg, got, emulate=True)
        print(f"whole path {name}: {n} tokens, differences {[(g[1], round(g[4], 4)) for g in r.gaps]}")
