"""Where does LayerNorm stand?  Times the engine's LayerNorm seam against torch's own elementwise conversion
(50 MB fp32 read + 25 MB bf16 write: the same traffic) with a warm and a flushed L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.engine import Engine
from reazonspeech_b200.weights import random_state_dict

cfg = ModelConfig.tiny()
eng = Engine(cfg, random_state_dict(cfg, 0), "cuda:0")
M, d = 12544, 1024
x = torch.randn(M, d, device="cuda")
g, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = torch.empty(M, d, dtype=torch.bfloat16, device="cuda")

def timeit(fn, warm, n=20):
    ts = []
    for _ in range(n):
        if not warm:
            flush.zero_()
        else:
            x.add_(0.0)            # touch x: leaves it (dirty) in L2 like the producing GEMM does
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

def ln_variant(v):
    def run():
        os.environ["RS_LN_VARIANT"] = v
        eng.layernorm(x, g, b)
    return run

for name, fn in (("engine layernorm A persistent", ln_variant("A")), ("engine layernorm B row/warp x32", ln_variant("B")),
                 ("torch x.to(bf16) copy_", lambda: out.copy_(x)),
                 ("torch layer_norm fp32->fp32", lambda: torch.nn.functional.layer_norm(x, (d,), g, b))):
    try:
        print(f"{name:32s} cold L2 {timeit(fn, False):6.1f} us   warm L2 {timeit(fn, True):6.1f} us")
    except Exception as ex:
        print(name, "failed:", ex)
