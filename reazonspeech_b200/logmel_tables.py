"""Tables of the fused log-mel kernel (csrc/logmel.cu): window, FFT twiddles in the kernel's lane order, and the mel
filterbank dealt to sixteen lanes in uniform "slots".

Replaces the buffers of NeMo's ``FilterbankFeatures`` (window, ``fb``) reached through ``model.transcribe``
(pkg/nemo-asr/src/transcribe.py:48-53)."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .config import ModelConfig
from .weights import hann_window, mel_filterbank

LANES = 16                 # lanes per frame (csrc/logmel_frame.cuh kLanes)
PAIRS = 8                  # bin pairs per lane
MAX_SLOTS = 8
META_INTS = 272            # [0] n_slots, [1] n_taps, [8 + s] taps of slot s, [16 + 16 s + t] first bin, [144 + 16 s + t] filter index
META_START, META_OUT = 16, 16 + LANES * MAX_SLOTS
N_BINS = 257


def mel_slots(fb: np.ndarray):
    """Deal the filters to the lanes: sorted by width (widest first), sixteen at a time = one slot; lane t gets the t-th
    filter of every slot.  Within a slot every lane runs as many taps as the slot's widest filter; a narrower filter is
    padded with zero weights, and its first bin is moved down where the padded range would leave the spectrum
    (the taps in front are then the zeros).  Returns (weights [n_taps, 16], meta int32[META_INTS], slots)."""
    n_mels = fb.shape[0]
    assert fb.shape[1] == N_BINS and n_mels <= LANES * MAX_SLOTS
    spans = []
    for m in range(n_mels):
        nz = np.nonzero(fb[m])[0]
        spans.append((int(nz[0]), int(nz[-1]) - int(nz[0]) + 1) if len(nz) else (0, 0))
    order = sorted(range(n_mels), key=lambda m: (-spans[m][1], m))
    slots = [order[i:i + LANES] for i in range(0, n_mels, LANES)]
    meta = np.zeros(META_INTS, dtype=np.int32)
    meta[0] = len(slots)
    rows = []
    for s, ms in enumerate(slots):
        c = max(max(spans[m][1] for m in ms), 1)
        meta[8 + s] = c
        w = np.zeros((c, LANES), dtype=np.float32)
        for t in range(LANES):
            if t < len(ms):
                m = ms[t]
                s0, cnt = spans[m]
                first = min(s0, N_BINS - c)                     # keep first + c <= 257
                w[s0 - first:s0 - first + cnt, t] = fb[m, s0:s0 + cnt]
                meta[META_START + s * LANES + t] = first
                meta[META_OUT + s * LANES + t] = m
            else:                                               # no filter: zero weights, result parked behind the 128 real rows
                meta[META_START + s * LANES + t] = 0
                meta[META_OUT + s * LANES + t] = LANES * MAX_SLOTS + t
        rows.append(w)
    weights = np.concatenate(rows, axis=0)
    meta[1] = weights.shape[0]
    return weights, meta, slots


def logmel_tables(cfg: ModelConfig) -> Dict[str, torch.Tensor]:
    if cfg.n_fft != 512:
        raise ValueError("the log-mel kernel is built for n_fft == 512")
    if cfg.n_window_stride % 2:
        raise ValueError("the log-mel kernel needs an even hop (8-byte aligned frame starts)")
    win = torch.zeros(cfg.n_fft, dtype=torch.float32)
    lo = (cfg.n_fft - cfg.n_window_size) // 2
    win[lo:lo + cfg.n_window_size] = hann_window(cfg)
    t = np.arange(LANES, dtype=np.float64)[None, :]
    k = np.arange(16, dtype=np.float64)[:, None]
    ang_b = 2 * np.pi * (t * k) / 256.0                          # inter-pass twiddles W256^(t k1) at [k1][t]
    tw_b = np.stack([np.cos(ang_b), -np.sin(ang_b)], axis=-1).astype(np.float32)
    k2 = np.arange(PAIRS, dtype=np.float64)[:, None]
    ang_x = 2 * np.pi * (t + 16 * k2) / 512.0                    # split twiddles W512^(t + 16 k2) at [k2][t]
    tw_x = np.stack([np.cos(ang_x), -np.sin(ang_x)], axis=-1).astype(np.float32)
    fb = mel_filterbank(cfg).numpy().astype(np.float32)
    weights, meta, _ = mel_slots(fb)
    # the kernel's powers are 4 |X|^2 (logmel_frame.cuh drops the 1/2 of the real-FFT split); a power of two, so exact
    return {"fe.window": win, "fe.tw_b": torch.from_numpy(tw_b).reshape(-1), "fe.tw_x": torch.from_numpy(tw_x).reshape(-1),
            "fe.mel_w": torch.from_numpy(weights * np.float32(0.25)).reshape(-1), "fe.mel_meta": torch.from_numpy(meta)}
