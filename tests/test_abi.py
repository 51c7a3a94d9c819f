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


def test_unet_operator_argument_errors_are_reported_without_a_gpu():
    """Argument validation happens before any CUDA call: bad shapes / alignments come back as SGN_ERR_INVALID_ARG with
    a message, on a box with or without a GPU."""
    lib = _lib.load()
    buf = (C.c_uint8 * 4096)()
    p = C.addressof(buf)
    p16 = (p + 15) & ~15
    ep = _lib.SgnEpilogue()
    assert lib.sgn_gemm_f16(p16, 60, p16, 60, 4, 16, 60, C.byref(ep), p16, None) == -1            # K % 8 != 0
    assert b"multiples of 8" in lib.sgn_last_error()
    assert lib.sgn_gemm_f16(p16 + 2, 64, p16, 64, 4, 16, 64, C.byref(ep), p16, None) == -1        # misaligned A
    assert lib.sgn_gemm_f16(None, 64, p16, 64, 4, 16, 64, C.byref(ep), p16, None) == -1           # null pointer
    assert lib.sgn_gemm_f16(None, 64, None, 64, 0, 16, 64, C.byref(ep), None, None) == 0          # empty M is a no-op
    ep.geglu, ep.d_residual = 1, p16
    assert lib.sgn_gemm_f16(p16, 64, p16, 64, 4, 16, 64, C.byref(ep), p16, None) == -1            # geglu + residual
    ep = _lib.SgnEpilogue()
    assert lib.sgn_conv3x3_f16(p16, p16, 1, 8, 8, 48, 16, C.byref(ep), p16, None) == -1           # Cin % 64 != 0
    assert b"Cin % 64" in lib.sgn_last_error()
    assert lib.sgn_attention_f16(p16, 60, p16, 64, p16, 64, 1, 1, 8, 8, 0.125, p16, 64, None) == -1   # stride % 8
    assert lib.sgn_attention_f16(p16, 64, p16, 64, p16, 64, 1, 2, 8, 8, 0.125, p16, 64, None) == -1   # stride < heads*64
    assert lib.sgn_group_norm_f16(p16, 1, 4, 30, 32, 1e-5, p16, p16, 0, p16, p16, None) == -1     # C % groups
    assert lib.sgn_layer_norm_f16(p16, 4, 30, 1e-5, p16, p16, p16, None) == -1                    # C % 4
    assert lib.sgn_linear_small(p16, p16, None, None, 9, 4, 4, 0, 0, p16, None) == -1             # more than 8 rows
    assert lib.sgn_cfg_euler_step(p16, p16, p16, None, None, 1, 4, 2, 2, 7.0, 1.0, 0.5, 0.5, p16, None, None) == -1  # init w/o mask
    assert lib.sgn_cfg_euler_step(p16, p16, None, None, None, 1, 4, 2, 2, 7.0, 0.0, 0.5, 0.5, p16, None, None) == -1  # sigma 0
    assert lib.sgn_sheet_to_conditioning(p16, p16, 12, 16, p16, p16, None) == -1                  # sheet not multiple of 8
    assert lib.sgn_group_norm_ws_doubles(2, 65536, 32) >= 2 * 32 * 3


def test_gemm_tile_schedule_plans():
    """sgn_gemm_plan (host-only probe, 148 SMs): the widths / tail splits the persistent GEMM picks for the UNet's shapes.
    A 2.16-wave GEMM gets its last round cut into column slices; short-K in-place-residual GEMMs get narrower tiles."""
    lib = _lib.load()

    def plan(M, N, K, res=0):
        p = (C.c_int * 6)()
        assert lib.sgn_gemm_plan(M, N, K, res, p) == 0
        return dict(zip(("bn", "cluster", "tiles", "split", "items", "units"), p))

    a = plan(8192, 1280, 1280)
    assert (a["bn"], a["cluster"], a["tiles"], a["units"]) == (256, 2, 160, 74) and a["split"] == 4 and a["items"] == 148 + 12 * 4
    assert plan(8192, 1280, 5120, 1)["split"] == 4
    b = plan(8192, 3840, 1280)                      # 6.9 waves: last round nearly full, nothing to split
    assert b["split"] == 1 and b["items"] == b["tiles"]
    c = plan(32768, 640, 640, 1)                    # epilogue-bound: narrower than the widest tile
    assert c["bn"] < 256 and c["bn"] % 16 == 0
    d = plan(100, 48, 64)                           # one M tile: single-CTA shape
    assert d["cluster"] == 1 and d["split"] == 1
    for M, N, K in ((8192, 10240, 1280), (32768, 5120, 640), (131072, 320, 320), (154, 153600, 2048)):
        q = plan(M, N, K)
        assert q["bn"] % 16 == 0 and 16 <= q["bn"] <= 256 and q["items"] >= q["tiles"]
        if q["split"] > 1:
            assert (q["bn"] // q["split"]) % 32 == 0 and q["bn"] // q["split"] >= 64
    assert lib.sgn_gemm_plan(0, 8, 8, 0, (C.c_int * 6)()) < 0
