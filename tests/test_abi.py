"""CPU: the C-ABI library loads without a GPU and exports every symbol include/signerf_b200.h declares;
calling it without a device fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

from signerf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(sgn_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes out of sync with the header"
    assert lib.sgn_abi_version() >= 1


def test_struct_layouts_match_c_sizes():
    # sizes the C compiler produces for the header's structs (LP64)
    assert C.sizeof(_lib.SgnHashGrid) == 24
    assert C.sizeof(_lib.SgnLinear) == 24
    assert C.sizeof(_lib.SgnRenderOpts) == 40
    assert C.sizeof(_lib.SgnMaskOpts) == 52


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    from signerf_b200.field import HashGridParams, NerfactoFieldB200
    with pytest.raises(RuntimeError):
        NerfactoFieldB200(HashGridParams(torch.zeros(16 * 16, 2), torch.ones(16), 4), [], [], torch.zeros(32), 0.01)
    d = _lib.SgnFieldDesc()
    h = C.c_void_p()
    rc = _lib.load().sgn_field_create(C.byref(d), C.byref(h))
    assert rc < 0 and not h
    assert _lib.load().sgn_last_error()
