"""The C-ABI library builds, loads and exports every symbol include/rs_engine.h declares.
No compute calls are made (there is no GPU here); error paths that do not need one are checked."""
import ctypes as C
import os
import re

import pytest

from reazonspeech_b200 import engine as E
from reazonspeech_b200.config import ModelConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rs_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = E.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rs_engine.h but not exported"
    assert sorted(E.EXPORTS) == names


def test_config_struct_matches_header():
    src = open(os.path.join(ROOT, "include", "rs_engine.h")).read()
    body = src[src.index("typedef struct rs_model_config {"):src.index("} rs_model_config;")]
    fields = re.findall(r"\b([a-z_0-9]+)\s*[,;]", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f[0] for f in E.RsModelConfig._fields_]
    assert C.sizeof(E.RsModelConfig) == 4 * len(fields)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = E.load_library()
    cfg = E.to_rs_config(ModelConfig.tiny())
    arr = (E.RsTensor * 1)()
    h = C.c_void_p()
    rc = lib.rs_engine_create(C.byref(cfg), arr, 0, 0, C.byref(h))
    assert rc != 0 and not h
    assert b"CUDA" in lib.rs_last_error(None) or b"device" in lib.rs_last_error(None)
    with pytest.raises(RuntimeError):
        E.Engine(ModelConfig.tiny(), {}, "cuda:0")


def test_load_model_contract():
    from reazonspeech_b200.nemo import asr
    with pytest.raises(RuntimeError):
        asr.load_model("cpu")
    os.environ.pop("REAZONSPEECH_B200_SYNTHETIC", None)
    os.environ.pop("REAZONSPEECH_NEMO_CHECKPOINT", None)
    with pytest.raises(FileNotFoundError):
        asr.load_model("cuda")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "reazonspeech_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "import_module(\"oracle" not in src and "__import__(\"oracle" not in src, f


def test_create_rejects_unsupported_configs_before_touching_a_device():
    """Configuration checks come first in rs_engine_create, so they are observable without a GPU."""
    lib = E.load_library()
    arr = (E.RsTensor * 1)()
    for field, value, needle in (("global_tokens", 2, b"global_tokens=2"), ("conv_kernel", 31, b"conv_kernel=31"),
                                 ("n_fft", 1024, b"n_fft=1024"), ("att_left", -1, b"att context")):
        cfg = E.to_rs_config(ModelConfig.tiny())
        setattr(cfg, field, value)
        h = C.c_void_p()
        rc = lib.rs_engine_create(C.byref(cfg), arr, 0, 0, C.byref(h))
        assert rc == -5 and not h, field                             # RS_ERR_UNSUPPORTED
        assert needle in lib.rs_last_error(None), lib.rs_last_error(None)


def test_stage_rows_host_helper():
    """rs_stage_rows: pad | samples | zeros rows, float and PCM16, threads or not, argument errors -- no GPU involved."""
    import ctypes as C
    import numpy as np
    from reazonspeech_b200.engine import load_library
    lib = load_library()
    rng = np.random.default_rng(0)
    waves = [rng.standard_normal(n).astype(np.float32) for n in (0, 1, 17, 4000, 999)]
    pcm = [(w * 1000).astype(np.int16) for w in waves]
    B, pad, L = len(waves), 5, 4012

    def call(dst, srcs, dst16, threads, is16=None, L_=L):
        ptr = (C.c_void_p * B)(*[a.ctypes.data for a in srcs])
        n = (C.c_int64 * B)(*[len(a) for a in srcs])
        flags = (C.c_int32 * B)(*is16) if is16 is not None else None
        return lib.rs_stage_rows(dst.ctypes.data, L_, ptr, n, flags, dst16, B, pad, threads)

    for threads in (1, 3, 64):
        dst = np.full((B, L), 7.0, np.float32)
        assert call(dst, waves, 0, threads) == 0
        for r, w in enumerate(waves):
            assert np.array_equal(dst[r], np.pad(w, (pad, L - pad - len(w))))
        d16 = np.full((B, L), 7, np.int16)
        assert call(d16, pcm, 1, threads, [1] * B) == 0
        for r, w in enumerate(pcm):
            assert np.array_equal(d16[r], np.pad(w, (pad, L - pad - len(w))))
        mixed = [pcm[0], waves[1], pcm[2], waves[3], pcm[4]]
        assert call(dst, mixed, 0, threads, [1, 0, 1, 0, 1]) == 0
        assert np.array_equal(dst[2, pad:pad + 17], pcm[2].astype(np.float32) / np.float32(32768.0))
        assert np.array_equal(dst[3, pad:pad + 4000], waves[3])
    assert call(np.zeros((B, L), np.float32), waves, 0, 1, L_=4000) != 0          # a row does not fit
    assert call(np.zeros((B, L), np.int16), waves, 1, 1, [0] * B) != 0            # int16 rows from float sources


def test_header_is_plain_c(tmp_path):
    """include/rs_engine.h is the drop-in boundary for a cgo / JNI / ctypes binding: it must compile as C99 on its own, and a C
    program that references every declared entry point must link against the library (no C++ or torch types in the signatures)."""
    import re
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from reazonspeech_b200.engine import _LIB_PATH, load_library
    load_library()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "rs_engine.h")).read()
    names = sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", header)))
    assert "rs_transcribe_batch" in names and "rs_stage_rows" in names
    src = tmp_path / "use_all.c"
    src.write_text('#include "rs_engine.h"\n#include <stdio.h>\nint main(void) {\n  const void* fns[] = {' +
                   ", ".join(f"(const void*){n}" for n in names) + "};\n  printf(\"%d\\n\", (int)(sizeof fns / sizeof fns[0]));\n  return 0;\n}\n")
    exe = tmp_path / "use_all"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                        _LIB_PATH, "-Wl,-rpath," + os.path.dirname(_LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
