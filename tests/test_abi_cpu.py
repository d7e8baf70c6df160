"""CPU-side checks of the product boundary: the library loads, exports every symbol include/tsl.h declares, and
refuses to run without a GPU (no fallback path)."""
import ctypes as C
import os
import re

import pytest
import torch

import thinshelllab_b200 as tb
from thinshelllab_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tsl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in C.cast(_lib.lib().tsl_version(), C.c_char_p).value or b"sm_100a" in _lib.lib().tsl_version()


def test_config_struct_size_is_checked():
    L = _lib.lib()
    cfg = _lib.Config()
    cfg.struct_size = 4
    ctx = C.c_void_p()
    assert L.tsl_create(C.byref(cfg), C.byref(ctx)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        tb.ShellEngine(10, 5e-3, k_contact=1e4, eps_contact=4e-4)
    L = _lib.lib()
    cfg = _lib.Config()
    cfg.struct_size = C.sizeof(_lib.Config)
    cfg.n_verts = 10
    ctx = C.c_void_p()
    assert L.tsl_create(C.byref(cfg), C.byref(ctx)) == -2       # TSL_ERR_CUDA


def test_product_never_imports_the_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "thinshelllab_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dp, f)).read()
                assert "tsl_oracle" not in txt and "ti_emu" not in txt, os.path.join(dp, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), os.path.join(dp, f)
