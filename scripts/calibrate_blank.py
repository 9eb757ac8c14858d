"""Measure the blank-bias shift that gives the synthetic checkpoint a speech-like greedy emission
rate (target: one token per three encoder frames).  The result is pasted into
reazonspeech_b200/weights.py::CALIBRATED_BLANK_SHIFT -- a one-parameter fit that only shapes the
decode LOAD of the seeded random weights; it is not part of the engine.

    python scripts/calibrate_blank.py --config tiny            # CPU, through the oracle
    python scripts/calibrate_blank.py --config full            # on a B200, through the engine
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.synth import synth_clip
from reazonspeech_b200.weights import _bf16_round, random_state_dict

TARGET = 1.0 / 3.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="tiny", choices=["tiny", "full"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cfg = ModelConfig.tiny() if a.config == "tiny" else ModelConfig()
    sd = random_state_dict(cfg, a.seed, blank_shift=0.0)
    secs = (6.0, 9.0, 12.0) if a.config == "tiny" else (30.0, 30.0, 30.0, 30.0)
    waves = [np.pad(synth_clip(100 + i, s), 8000) for i, s in enumerate(secs)]
    frames = sum(cfg.enc_frames(len(w)) for w in waves)
    base = sd["joint.joint_net.2.bias"].clone()

    if a.config == "tiny":
        from oracle import nemo_restated as O
        torch.set_num_threads(1)
        with torch.no_grad():
            encs = [O.encoder(O.log_mel(torch.from_numpy(w), cfg), sd, cfg) for w in waves]

        def rate(shift):
            b = base.clone(); b[cfg.blank] += shift
            sd["joint.joint_net.2.bias"] = _bf16_round(b)
            return sum(len(O.rnnt_greedy(e, sd, cfg).tokens) for e in encs) / frames
    else:
        from reazonspeech_b200.engine import Engine
        eng = Engine(cfg, sd, "cuda:0")
        L = max(len(w) for w in waves)
        x = torch.zeros(len(waves), L)
        for i, w in enumerate(waves):
            x[i, : len(w)] = torch.from_numpy(w)
        lens = torch.tensor([len(w) for w in waves], dtype=torch.int32).cuda()
        x = x.cuda()
        bias_dev = eng.weights["joint.out.b"]

        def rate(shift):
            b = base.clone(); b[cfg.blank] += shift
            bias_dev.copy_(_bf16_round(b))
            _, _, ntok = eng.transcribe_device(x, lens)
            torch.cuda.synchronize()
            return float(ntok.sum()) / frames

    lo, hi = -4.0, 8.0            # emission rate decreases monotonically with the shift
    r_lo, r_hi = rate(lo), rate(hi)
    print(f"rate({lo})={r_lo:.3f} rate({hi})={r_hi:.3f}")
    for _ in range(24):
        mid = 0.5 * (lo + hi)
        r = rate(mid)
        if r > TARGET:
            lo = mid
        else:
            hi = mid
    shift = float(_bf16_round(torch.tensor(0.5 * (lo + hi))))
    res = {"config": a.config, "seed": a.seed, "shift": shift, "rate": rate(shift), "frames": frames}
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"))


if __name__ == "__main__":
    main()
