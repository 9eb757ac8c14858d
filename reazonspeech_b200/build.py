"""Build librs_engine.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles without a GPU.  Objects are cached under reazonspeech_b200/csrc/build/
and rebuilt when the source (or any header) is newer.  Usage: ``python -m reazonspeech_b200.build``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "librs_engine.so")
SOURCES = ["gemm_tcgen05.cu", "logmel.cu", "subsample.cu", "elementwise.cu", "attention.cu", "attention_tc.cu", "decode.cu", "decode_batched.cu", "decode_spec.cu", "engine.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "rs_engine.h"))
    return max(os.path.getmtime(h) for h in hs)


# Compile-time variants of the library (EXPERIMENTS; the default library is what every test and the bench load unless
# RS_ENGINE_VARIANT names one).  "pdl": programmatic dependent launch in the encoder's kernels, see csrc/common.cuh.
VARIANTS = {"pdl": ["-DRS_PDL=1"]}


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    if variant and variant not in VARIANTS:
        raise ValueError(f"unknown variant {variant!r} (known: {sorted(VARIANTS)})")
    obj_dir = OBJ + ("_" + variant if variant else "")
    lib = LIB.replace(".so", f"_{variant}.so") if variant else LIB
    defines = VARIANTS.get(variant, [])
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    hm = _headers_mtime()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc, *ARCH, *FLAGS, *defines, "-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(lib):
        cmd = [nvcc, *ARCH, "-shared", "-o", lib, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    _variant = ""
    if "--variant" in sys.argv:
        _variant = sys.argv[sys.argv.index("--variant") + 1]
    print(build(force="--force" in sys.argv, verbose=True, variant=_variant))
