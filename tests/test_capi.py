"""CPU: the C-ABI library loads, exports every symbol declared in include/rin_b200.h and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "robust-implicit-surface-networks_b200", "librin_b200.so")
HDR = os.path.join(ROOT, "include", "rin_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__  # noqa
        __graft_entry__.build()
    return C.CDLL(LIB)


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rin_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib.rin_last_error.restype = C.c_char_p
    h = C.c_void_p()
    rc = lib.rin_create(0, C.byref(h))
    assert rc == -1  # RIN_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.rin_last_error()
    assert lib.rin_device_count() == 0
