"""RS_GEMM_SPLITK=1 (EXPERIMENT, written without GPU time left in round 1 -- DESIGN.md section 8): split-K of the last,
partial wave of the persistent 2-CTA GEMM.  Not part of the default `-m gpu` suite:

    RS_RUN_EXPERIMENTS=1 timeout 300 python -m pytest tests/experiments/test_gpu_splitk.py -m gpu -q -s

(run it under `timeout`: an owner tile waits on flags, a scheduling bug would hang rather than fail).  Shapes are the
N = 1024 GEMMs of the full model at 32 and 5 clips; the split-K result must equal the default kernel's up to the fp32
re-association of three partial sums, for every epilogue of the common group, and repeated launches must reuse the
flag workspace correctly."""
import math
import os

import pytest
import torch

from reazonspeech_b200 import engine as E

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("RS_RUN_EXPERIMENTS") != "1", reason="experiment: set RS_RUN_EXPERIMENTS=1")]


@pytest.fixture(scope="module")
def splitk_engine(tiny_cfg, tiny_sd):
    os.environ["RS_GEMM_SPLITK"] = "1"
    try:
        return E.Engine(tiny_cfg, tiny_sd, "cuda:0")
    finally:
        del os.environ["RS_GEMM_SPLITK"]


@pytest.mark.parametrize("M,N,K,epi", [(12544, 1024, 4096, E.EPI_RESID_F32), (12416, 1024, 4096, E.EPI_BIAS_F32),
                                       (1960, 1024, 4096, E.EPI_RESID_F32), (12544, 1024, 2560, E.EPI_BIAS_F32),
                                       (12544, 2048, 4096, E.EPI_BIAS_SWISH_BF16), (12544, 1024, 2048, E.EPI_BIAS_BF16)])
def test_splitk_tail_equals_default_kernel(tiny_engine, splitk_engine, M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == E.EPI_RESID_F32 else None
    kw = dict(alpha=0.5) if epi in (E.EPI_RESID_F32, E.EPI_BIAS_F32) else {}
    ref = tiny_engine.gemm(a, w, bias, epi, resid=None if resid is None else resid.clone(), **kw).float()
    for rep in range(3):                                            # the flags must come back to zero after every launch
        got = splitk_engine.gemm(a, w, bias, epi, resid=None if resid is None else resid.clone(), **kw).float()
        torch.cuda.synchronize()
        err = ((got - ref).abs() / (ref.abs() + 1.0)).max().item()
        print(f"M={M} N={N} K={K} epi={epi} rep={rep}: max scaled diff {err:.2e}")
        assert err < (1e-5 if got.dtype == torch.float32 and epi in (E.EPI_RESID_F32, E.EPI_BIAS_F32) else 8e-3)
    truth = a.float() @ w.float().T + bias
    if epi == E.EPI_BIAS_F32:
        assert ((got - 0.5 * truth).abs().max() < 2e-3 * max(1.0, truth.abs().max().item()))
