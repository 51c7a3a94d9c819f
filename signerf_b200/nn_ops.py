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


# ---------------------------------------------------------------------------------------------- per-family profiling
# bench.py brackets every C-ABI call with CUDA events (eager pass, outside the timed region) to split a step's device
# time and algorithmic FLOPs by kernel family: the roofline numerators of the JSON line come from here.
_PROFILE: Optional[dict] = None
PROFILE_SHAPES = False   # scratch tooling: key GEMM / attention spans by shape


def start_profile() -> None:
    global _PROFILE
    _PROFILE = {}


def stop_profile() -> dict:
    """-> {family: {"ms": device time, "flops": algorithmic FLOPs, "calls": n}}"""
    global _PROFILE
    prof, _PROFILE = _PROFILE or {}, None
    torch.cuda.synchronize()
    out = {}
    for name, spans in prof.items():
        out[name] = {"ms": sum(a.elapsed_time(b) for a, b, _ in spans), "flops": float(sum(f for _, _, f in spans)),
                     "calls": len(spans)}
    return out


class _span:
    def __init__(self, name: str, flops: float = 0.0):
        self.name, self.flops = name, flops

    def __enter__(self):
        if _PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.b.record()
            _PROFILE.setdefault(self.name, []).append((self.a, self.b, self.flops))
        return False


def _chk(t: Optional[Tensor], dtype: torch.dtype, name: str) -> None:
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (signerf_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _epilogue(bias, rowbias, rows_per_batch, residual, out_f16, geglu, nchw, ldo=0, act_silu=False) -> _lib.SgnEpilogue:
    _chk(bias, torch.float32, "bias")
    _chk(rowbias, torch.float32, "rowbias")
    _chk(residual, torch.float32, "residual")
    e = _lib.SgnEpilogue()
    e.d_bias = None if bias is None else bias.data_ptr()
    e.d_rowbias = None if rowbias is None else rowbias.data_ptr()
    e.rows_per_batch = int(rows_per_batch)
    e.d_residual = None if residual is None else residual.data_ptr()
    e.ldo = int(ldo)
    e.out_f16, e.geglu, e.nchw, e.act_silu = int(out_f16), int(geglu), int(nchw), int(act_silu)
    return e


def gemm_f16(a: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, rowbias: Optional[Tensor] = None,
             rows_per_batch: int = 0, residual: Optional[Tensor] = None, out_f16: bool = False, geglu: bool = False,
             act_silu: bool = False, out: Optional[Tensor] = None) -> Tensor:
    """out[M,N] = a[M,K] @ w[N,K]^T + bias (+ rowbias[m // rows_per_batch]) (+ residual); geglu: [M, N/2] fp16.
    a / w may be column-slice views of wider matrices (unit column stride, row stride a multiple of 8)."""
    for t, n in ((a, "a"), (w, "w")):
        if not t.is_cuda:
            raise RuntimeError(f"{n} must be a CUDA tensor (signerf_b200 has no CPU path)")
        if t.dtype != torch.float16 or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 8 != 0:
            raise ValueError(f"{n} must be an fp16 matrix with unit column stride and a row stride that is a multiple of 8")
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
    e = _epilogue(bias, rowbias, rows_per_batch, residual, out_f16, geglu, False, act_silu=act_silu)
    with torch.cuda.device(a.device), _span(f"k_gemm_tc (linear) {M}x{N}x{K}{' geglu' if geglu else ''}{' res' if residual is not None else ''}{' f16' if out_f16 else ''}" if PROFILE_SHAPES else "k_gemm_tc (linear)", 2.0 * M * N * K):
        _lib.check(_lib.load().sgn_gemm_f16(_ptr(a), a.stride(0), _ptr(w), w.stride(0), M, N, K, C.byref(e), _ptr(out),
                                            _stream(a.device)))
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
    with torch.cuda.device(x.device), _span(f"k_gemm_tc (conv3x3) {B}x{H}x{W}x{Cin}->{N}" if PROFILE_SHAPES else "k_gemm_tc (conv3x3)", 2.0 * B * H * W * N * 9 * Cin):
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
    with torch.cuda.device(q.device), _span(f"k_attention_tc {batch}x{heads}x{Tq}x{Tkv}" if PROFILE_SHAPES else "k_attention_tc", 4.0 * batch * heads * Tq * Tkv * 64):
        lib = _lib.load()
        need = int(lib.sgn_attention_workspace_bytes(batch, heads, Tq, Tkv))   # > 0: tail items are split over the keys
        ws = torch.empty(need, dtype=torch.uint8, device=q.device) if need else None
        _lib.check(lib.sgn_attention_f16_ws(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0),
                                            batch, heads, Tq, Tkv, 0.125, _ptr(out), out.stride(0),
                                            _ptr(ws) if need else None, need, _stream(q.device)))
    return out


def _call(dev, fn, *args) -> None:
    with torch.cuda.device(dev), _span(fn.__name__):
        _lib.check(fn(*args, _stream(dev)))


def group_norm_f16(x: Tensor, B: int, HW: int, groups: int, eps: float, gamma: Tensor, beta: Tensor, act_silu: bool,
                   ws: Optional[Tensor] = None, split: bool = False) -> Tensor:
    """GroupNorm(groups) (+SiLU) of fp32 [B*HW, C] -> fp16 [B*HW, C]; ws: float64 scratch (allocated when None).
    split: fp16 [B*HW, 2C] = [hi | lo] (y = hi + lo to 2^-22)."""
    _chk(x, torch.float32, "x")
    need = int(_lib.load().sgn_group_norm_ws_doubles(B, HW, groups))
    if ws is None:
        ws = torch.empty(need, dtype=torch.float64, device=x.device)
    _chk(ws, torch.float64, "ws")
    if ws.numel() < need:
        raise ValueError(f"GroupNorm scratch needs {need} doubles, got {ws.numel()}")
    out = torch.empty((x.shape[0], x.shape[1] * (2 if split else 1)), dtype=torch.float16, device=x.device)
    fn = _lib.load().sgn_group_norm_split_f16 if split else _lib.load().sgn_group_norm_f16
    _call(x.device, fn, _ptr(x), B, HW, x.shape[1], groups, eps, _ptr(gamma), _ptr(beta), int(act_silu), _ptr(ws), _ptr(out))
    return out


def layer_norm_f16(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_layer_norm_f16, _ptr(x), x.shape[0], x.shape[1], eps, _ptr(gamma), _ptr(beta), _ptr(out))
    return out


def cast_f16(x: Tensor) -> Tensor:
    _chk(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_cast_f16, _ptr(x), x.numel(), _ptr(out))
    return out


def upsample2x_f16(x: Tensor, B: int, H: int, W: int) -> Tensor:
    """fp32 [B*H*W, C] -> fp16 NHWC [B, 2H, 2W, C] (nearest)."""
    _chk(x, torch.float32, "x")
    Cc = x.shape[-1]
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_upsample2x_f16, _ptr(x), B, H, W, Cc, _ptr(out))
    return out


def concat_f32(a: Tensor, b: Tensor, b2: Optional[Tensor] = None, scale: float = 1.0, with_f16: bool = False):
    """cat([a, b + scale*b2], dim=1) of fp32 [P, C] matrices; with_f16: -> (fp32, fp16 copy)."""
    _chk(a, torch.float32, "a")
    _chk(b, torch.float32, "b")
    _chk(b2, torch.float32, "b2")
    P = a.shape[0]
    out = torch.empty((P, a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
    if with_f16:
        out16 = torch.empty(out.shape, dtype=torch.float16, device=a.device)
        _call(a.device, _lib.load().sgn_concat_f32_f16, _ptr(a), a.shape[1], _ptr(b), _ptr(b2), float(scale), b.shape[1], P,
              _ptr(out), _ptr(out16))
        return out, out16
    _call(a.device, _lib.load().sgn_concat_f32, _ptr(a), a.shape[1], _ptr(b), _ptr(b2), float(scale), b.shape[1], P, _ptr(out))
    return out


def axpy_f32(x: Tensor, a: float, y: Tensor) -> None:
    _chk(x, torch.float32, "x")
    _chk(y, torch.float32, "y")
    _call(y.device, _lib.load().sgn_axpy_f32, _ptr(x), float(a), y.numel(), _ptr(y))


def im2col3x3_s2_f16(x: Tensor, B: int, H: int, W: int):
    """fp32 [B*H*W, C] -> (fp16 [B*Ho*Wo, 9C], Ho, Wo) for a 3x3 / stride 2 / pad 1 conv."""
    _chk(x, torch.float32, "x")
    Cc = x.shape[-1]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((B * Ho * Wo, 9 * Cc), dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_im2col3x3_s2_f16, _ptr(x), B, H, W, Cc, _ptr(out))
    return out, Ho, Wo


def conv3x3_direct(x: Tensor, in_nchw: bool, w: Tensor, bias: Optional[Tensor], residual: Optional[Tensor] = None,
                   stride: int = 1, act_silu: bool = False, out_f16: bool = False) -> Tensor:
    """fp32 direct conv; x NCHW [B,C,H,W] or NHWC [B,H,W,C]; w fp32 [Cout,3,3,Cin] -> NHWC [B,Ho,Wo,Cout]."""
    _chk(x, torch.float32, "x")
    _chk(w, torch.float32, "w")
    _chk(residual, torch.float32, "residual")
    if in_nchw:
        B, Cin, H, W = x.shape
    else:
        B, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float16 if out_f16 else torch.float32, device=x.device)
    res_batch = residual.shape[0] if residual is not None else 0
    _call(x.device, _lib.load().sgn_conv3x3_direct, _ptr(x), int(in_nchw), _ptr(w), _ptr(bias), _ptr(residual), res_batch,
          B, H, W, Cin, Cout, stride, int(act_silu), int(out_f16), _ptr(out))
    return out


def linear_small_segments(x: Tensor, w_cat: Tensor, b_cat: Tensor, seg_offsets: Tensor, silu_in: bool = False) -> Tensor:
    """Several small linears on the same input in one launch: w_cat [sum N_j, K], seg_offsets int32 [nseg + 1] (device);
    -> flat fp32 buffer in which segment j is the contiguous [B, N_j] matrix at B * seg[j]."""
    _chk(x, torch.float32, "x")
    _chk(w_cat, torch.float32, "w_cat")
    _chk(b_cat, torch.float32, "b_cat")
    _chk(seg_offsets, torch.int32, "seg_offsets")
    B, Kd = x.shape
    N = w_cat.shape[0]
    out = torch.empty(B * N, dtype=torch.float32, device=x.device)
    _call(x.device, _lib.load().sgn_linear_small_segments, _ptr(x), _ptr(w_cat), _ptr(b_cat), B, N, Kd, int(silu_in),
          _ptr(seg_offsets), seg_offsets.numel() - 1, _ptr(out))
    return out


def conv3x3_small_tc(x: Tensor, w16: Tensor, bias: Optional[Tensor], stride: int = 1, act_silu: bool = False,
                     out_f16: bool = False) -> Tensor:
    """fp32-exact 3x3 / p1 conv for 16 / 32 channels on mma.sync: x fp32 NHWC [B,H,W,Cin], w16 fp16 [Cout, 9*Cin]."""
    _chk(x, torch.float32, "x")
    _chk(w16, torch.float16, "w16")
    _chk(bias, torch.float32, "bias")
    B, H, W, Cin = x.shape
    Cout = w16.shape[0]
    if w16.shape[1] != 9 * Cin:
        raise ValueError(f"w16 must be [Cout, {9 * Cin}], got {tuple(w16.shape)}")
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float16 if out_f16 else torch.float32, device=x.device)
    _call(x.device, _lib.load().sgn_conv3x3_small_tc, _ptr(x), _ptr(w16), _ptr(bias), B, H, W, Cin, Cout, int(stride),
          int(act_silu), int(out_f16), _ptr(out))
    return out


def linear_small(x: Tensor, w: Tensor, bias: Optional[Tensor], residual: Optional[Tensor] = None, silu_in: bool = False,
                 silu_out: bool = False) -> Tensor:
    _chk(x, torch.float32, "x")
    _chk(w, torch.float32, "w")
    B, Kd = x.shape
    N = w.shape[0]
    out = torch.empty((B, N), dtype=torch.float32, device=x.device)
    _call(x.device, _lib.load().sgn_linear_small, _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), B, N, Kd, int(silu_in),
          int(silu_out), _ptr(out))
    return out


def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    _chk(t, torch.float32, "t")
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    _call(t.device, _lib.load().sgn_timestep_embedding, _ptr(t), t.shape[0], dim, _ptr(out))
    return out


def scale_cat2(x: Tensor, scale: float) -> Tensor:
    """cat([x, x]) * scale along the batch dimension."""
    _chk(x, torch.float32, "x")
    out = torch.empty((2 * x.shape[0],) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    _call(x.device, _lib.load().sgn_scale_repeat_f32, _ptr(x), x.numel(), float(scale), 2, _ptr(out))
    return out


def make_hint_and_latent_mask(cond_sheet: Tensor, mask_sheet: Tensor, hint: Tensor, lat_mask: Tensor) -> None:
    _chk(cond_sheet, torch.float32, "cond_sheet")
    _chk(mask_sheet, torch.float32, "mask_sheet")
    Hs, Ws = cond_sheet.shape[0], cond_sheet.shape[1]
    if hint.numel() != 3 * Hs * Ws or lat_mask.numel() != (Hs // 8) * (Ws // 8):
        raise ValueError("hint / lat_mask do not match the sheet size")
    _call(hint.device, _lib.load().sgn_sheet_to_conditioning, _ptr(cond_sheet), _ptr(mask_sheet), Hs, Ws, _ptr(hint),
          _ptr(lat_mask))


def cfg_euler_step(x: Tensor, eps: Tensor, init: Optional[Tensor], mask: Optional[Tensor], noise: Optional[Tensor],
                   cfg_scale: float, sigma: float, sigma_down: float, sigma_up: float):
    """Fused CFG + inpaint blend + Euler-ancestral update on latents [B,C,H,W]; eps [2B,C,H,W] = (cond, uncond)."""
    for t, n in ((x, "x"), (eps, "eps"), (init, "init"), (mask, "mask"), (noise, "noise")):
        _chk(t, torch.float32, n)
    B, Cc, H, W = x.shape
    x_out, den = torch.empty_like(x), torch.empty_like(x)
    _call(x.device, _lib.load().sgn_cfg_euler_step, _ptr(x), _ptr(eps), _ptr(init), _ptr(mask), _ptr(noise), B, Cc, H, W,
          float(cfg_scale), float(sigma), float(sigma_down), float(sigma_up), _ptr(x_out), _ptr(den))
    return x_out, den


def im2col3x3_split_f16(x: Tensor, in_nchw: bool, stride: int = 1):
    """fp32 image (NCHW or NHWC) -> (fp16 [B*Ho*Wo, 2*Kp] = [hi | lo] patches, Ho, Wo, Kp) for a 3x3 / pad 1 conv."""
    _chk(x, torch.float32, "x")
    if in_nchw:
        B, Cc, H, W = x.shape
    else:
        B, H, W, Cc = x.shape
    Ho, Wo, Kp = (H - 1) // stride + 1, (W - 1) // stride + 1, (9 * Cc + 7) // 8 * 8
    out = torch.empty((B * Ho * Wo, 2 * Kp), dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_im2col3x3_split_f16, _ptr(x), int(in_nchw), B, H, W, Cc, stride, _ptr(out))
    return out, Ho, Wo, Kp
