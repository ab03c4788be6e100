"""The C-ABI library loads and exports every symbol include/tdr.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

import top_down_renderer_b200 as tdr
from top_down_renderer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tdr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdr_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    tdr.build()
    return C.CDLL(_lib.LIB_PATH)


def test_header_and_binding_list_agree():
    assert _declared() == sorted(_lib.SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in _declared() if not hasattr(lib, s)]
    assert not missing, missing


def test_abi_version_and_struct_sizes(lib):
    assert lib.tdr_abi_version() == 1
    assert C.sizeof(_lib.TdrState) == 28          # State, state_particle.h:9-17
    assert C.sizeof(_lib.TdrFilterParams) == 24 + 64


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.tdr_create(C.byref(h), 0)
    assert rc == _lib.TDR_ENOGPU and not h.value
    lib.tdr_last_error.restype = C.c_char_p
    assert b"no CPU fallback" in lib.tdr_last_error()
    from top_down_renderer_b200.core import Context
    with pytest.raises(tdr.TdrError):
        Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "top_down_renderer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tdr_oracle" not in txt, f
