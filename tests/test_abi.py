"""CPU test: the C-ABI library loads and exports every symbol include/citcomcu_b200.h declares
(no compute calls without a GPU) and fails loudly without a device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from conftest import ROOT, has_gpu

from citcomcu_b200 import _lib


def declared_symbols():
    txt = (ROOT / "include" / "citcomcu_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ccu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    _lib.build_library()
    lib = C.CDLL(str(_lib.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in citcomcu_b200.h but not exported"


@pytest.mark.skipif(has_gpu(), reason="checks the no-device failure path")
def test_create_fails_loudly_without_device():
    from citcomcu_b200.stokes import StokesContext
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        StokesContext(0, 1, {0: 3, 1: 5}, {0: 3, 1: 5}, {0: 3, 1: 5})
