"""Torch-tensor front end of the UNet operators of the C ABI (include/signerf_b200.h, K5-K9).  Like ops.py these
are argument marshallers only: tensors supply device pointers and the current stream, the arithmetic runs in the
sm_100a kernels.  Activations are NHWC; GEMM operands fp16, the residual stream fp32."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .ops import _ptr, _stream


def _chk(t: Optional[Tensor], dtype: torch.dtype, name: str) -> None:
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (signerf_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _epilogue(bias, rowbias, rows_per_batch, residual, out_f16, geglu, nchw, ldo=0) -> _lib.SgnEpilogue:
    _chk(bias, torch.float32, "bias")
    _chk(rowbias, torch.float32, "rowbias")
    _chk(residual, torch.float32, "residual")
    e = _lib.SgnEpilogue()
    e.d_bias = None if bias is None else bias.data_ptr()
    e.d_rowbias = None if rowbias is None else rowbias.data_ptr()
    e.rows_per_batch = int(rows_per_batch)
    e.d_residual = None if residual is None else residual.data_ptr()
    e.ldo = int(ldo)
    e.out_f16, e.geglu, e.nchw = int(out_f16), int(geglu), int(nchw)
    return e


def gemm_f16(a: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, rowbias: Optional[Tensor] = None,
             rows_per_batch: int = 0, residual: Optional[Tensor] = None, out_f16: bool = False, geglu: bool = False,
             out: Optional[Tensor] = None) -> Tensor:
    """out[M,N] = a[M,K] @ w[N,K]^T + bias (+ rowbias[m // rows_per_batch]) (+ residual); geglu: [M, N/2] fp16."""
    _chk(a, torch.float16, "a")
    _chk(w, torch.float16, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"a is [M,{K}] but w is {tuple(w.shape)}")
    odt = torch.float16 if (out_f16 or geglu) else torch.float32
    oshape = (M, N // 2 if geglu else N)
    if out is None:
        out = torch.empty(oshape, dtype=odt, device=a.device)
    elif out.dtype != odt or tuple(out.shape) != oshape or not out.is_contiguous():
        raise ValueError("bad `out` tensor")
    e = _epilogue(bias, rowbias, rows_per_batch, residual, out_f16, geglu, False)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().sgn_gemm_f16(_ptr(a), K, _ptr(w), K, M, N, K, C.byref(e), _ptr(out), _stream(a.device)))
    return out


def conv3x3_f16(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, rowbias: Optional[Tensor] = None,
                residual: Optional[Tensor] = None, out_f16: bool = False, nchw: bool = False,
                out: Optional[Tensor] = None) -> Tensor:
    """3x3/s1/p1 conv, x fp16 NHWC [B,H,W,C], w fp16 [N, 9C] (tap-major) -> [B*H*W, N] (or fp32 NCHW [B,N,H,W])."""
    _chk(x, torch.float16, "x")
    _chk(w, torch.float16, "w")
    B, H, W, Cin = x.shape
    N = w.shape[0]
    if w.shape[1] != 9 * Cin:
        raise ValueError(f"w must be [N, {9 * Cin}], got {tuple(w.shape)}")
    odt = torch.float16 if out_f16 else torch.float32
    oshape = (B, N, H, W) if nchw else (B * H * W, N)
    if out is None:
        out = torch.empty(oshape, dtype=odt, device=x.device)
    elif out.dtype != odt or tuple(out.shape) != oshape or not out.is_contiguous():
        raise ValueError("bad `out` tensor")
    e = _epilogue(bias, rowbias, H * W if rowbias is not None else 0, residual, out_f16, False, nchw)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().sgn_conv3x3_f16(_ptr(x), _ptr(w), B, H, W, Cin, N, C.byref(e), _ptr(out),
                                               _stream(x.device)))
    return out


def attention_f16(q: Tensor, k: Tensor, v: Tensor, batch: int, heads: int, out: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / 8) v per (image, head); q [B*Tq, >=64*heads], k / v [B*Tkv, ...] fp16 (column-slice views of a
    fused projection are fine: only the last dimension must be contiguous) -> [B*Tq, 64*heads] fp16."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.dtype != torch.float16 or t.stride(1) != 1:
            raise ValueError(f"{n} must be a CUDA fp16 matrix with unit column stride")
    Tq, Tkv = q.shape[0] // batch, k.shape[0] // batch
    if out is None:
        out = torch.empty((q.shape[0], heads * 64), dtype=torch.float16, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().sgn_attention_f16(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0),
                                                 batch, heads, Tq, Tkv, 0.125, _ptr(out), out.stride(0),
                                                 _stream(q.device)))
    return out
