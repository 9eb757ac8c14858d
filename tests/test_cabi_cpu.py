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
