"""Where a 2-CTA GEMM launch spends its time: clock64 timeline of the first and last cluster (Engine.gemm_cycles) for the
encoder's shapes at the bench geometry (M = 32 x 388), next to the CUDA-event time of the same launch (warm, median of 20).

    RS_BUILD_FLAGS=-DRS_PROF python -m reazonspeech_b200.build --force     # the stamps are not in the shipped build
    python scripts/diag_gemm_timeline.py            # on a B200
    python -m reazonspeech_b200.build --force       # back to the shipped build
"""
import json
import statistics
import sys

import torch

sys.path.insert(0, ".")
from reazonspeech_b200 import engine as E                     # noqa: E402
from reazonspeech_b200.config import ModelConfig              # noqa: E402
from reazonspeech_b200.weights import random_state_dict       # noqa: E402


def main():
    cfg = ModelConfig.tiny()
    eng = E.Engine(cfg, random_state_dict(cfg, 0), "cuda:0")
    M = 32 * 388
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = [("out-proj / pw2 (in-place residual)", 1024, 1024, E.EPI_RESID_F32, True),
              ("same, residual read in registers", 1024, 1024, E.EPI_RESID_F32, False),
              ("FFN W2 (in-place residual)", 1024, 4096, E.EPI_RESID_F32, True),
              ("FFN W1 (swish)", 4096, 1024, E.EPI_BIAS_SWISH_BF16, None),
              ("pw1 (GLU)", 2048, 1024, E.EPI_BIAS_GLU_BF16, None)]
    for name, N, K, epi, in_place in shapes:
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda", generator=g)
        x = torch.randn(M, N, device="cuda", generator=g) if in_place is not None else None
        out = x if in_place else (torch.empty(M, N, device="cuda") if in_place is False else None)
        times = {"warm": [], "flushed": []}
        for mode in ("warm", "flushed"):
            for _ in range(12):
                if mode == "flushed":
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.gemm(a, w, bias, epi, resid=x, alpha=0.5, out=out)
                e1.record()
                torch.cuda.synchronize()
                times[mode].append(e0.elapsed_time(e1) * 1e3)
        tl = eng.gemm_cycles()
        flops = 2.0 * M * N * K
        us = {k: statistics.median(v[2:]) for k, v in times.items()}
        print(f"== {name}: N={N} K={K}  warm {us['warm']:.1f} us ({flops / us['warm'] / 1e6:.0f} TFLOP/s)  L2-flushed {us['flushed']:.1f} us")
        print(json.dumps(tl))


if __name__ == "__main__":
    main()
