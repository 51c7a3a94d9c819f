"""SDXL first-stage autoencoder on B200: the VAE encode / decode the A1111 server wraps around the
denoising loop of the reference's img2img request (signerf/diffuser/diffuser.py:132-180; A1111 processing.py
`images_tensor_to_samples` / `decode_latent_batch`).  Host side only: this module walks the ldm `Encoder` / `Decoder`
graphs (sgm/modules/diffusionmodules/model.py) and issues C-ABI calls —

    3x3 convs, 1x1 shortcuts, q/k/v/proj        tcgen05 implicit-GEMM conv / GEMM  (sgn_conv3x3_f16, sgn_gemm_f16)
    GroupNorm(32, eps 1e-6) + swish              sgn_group_norm_f16
    Downsample (asymmetric pad, stride 2)        sgn_im2col3x3_s2_asym_f16 + GEMM
    Upsample (nearest x2 + conv)                 sgn_upsample2x_f16 + conv
    mid AttnBlock (1 head of 512, 65 536 tokens) scores and P.V as plain GEMMs around sgn_softmax_rows_f16: at head
                                                 dim 512 the O accumulator alone would fill tensor memory, and with
                                                 180 GB of HBM a 2 GB score slab per 8 192 queries is cheap
    conv_in / quant convs / latent sampling      sgn_conv3x3_direct, sgn_pointwise_nchw, sgn_vae_sample_latent

Weights: a state_dict with the upstream names (`encoder.down.0.block.0.norm1.weight`, ...; strip `first_stage_model.`
from an SDXL checkpoint).  Activations NHWC, fp32 residual stream, fp16 tensor-core operands (as unet.py).
No torch arithmetic, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Mapping, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from . import nn_ops as K
from .nn_ops import _call, _chk
from .ops import _ptr
from .unet import Act, RandomWeights, _Packed

SCALE_FACTOR = 0.13025


@dataclass
class VAEConfig:
    """ddconfig of the SDXL VAE (sd_xl_base.yaml first_stage_config)."""
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    in_channels: int = 3
    out_ch: int = 3
    z_channels: int = 4
    double_z: bool = True


# ---------------------------------------------------------------------------------------------- operator front ends
def im2col3x3_s2_asym_f16(x: Tensor, B: int, H: int, W: int, split: bool = False):
    """fp32 [B*H*W, C] -> (fp16 [B*Ho*Wo, 9C], Ho, Wo) for ldm Downsample (pad (0,1,0,1), 3x3, stride 2);
    split: rows [hi(9C) | lo(9C)]."""
    _chk(x, torch.float32, "x")
    Cc = x.shape[-1]
    Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    out = torch.empty((B * Ho * Wo, (18 if split else 9) * Cc), dtype=torch.float16, device=x.device)
    fn = _lib.load().sgn_im2col3x3_s2_asym_split_f16 if split else _lib.load().sgn_im2col3x3_s2_asym_f16
    _call(x.device, fn, _ptr(x), B, H, W, Cc, _ptr(out))
    return out, Ho, Wo


def split_f16(x: Tensor) -> Tensor:
    """fp32 [M,C] -> fp16 [M,2C] = [hi | lo]."""
    _chk(x, torch.float32, "x")
    out = torch.empty((x.shape[0], 2 * x.shape[1]), dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_split_f16, _ptr(x), x.shape[0], x.shape[1], _ptr(out))
    return out


def upsample2x_split_f16(x: Tensor, B: int, H: int, W: int) -> Tensor:
    """fp32 [B*H*W, C] -> fp16 NHWC [B, 2H, 2W, 2C] = nearest upsample, [hi | lo]."""
    _chk(x, torch.float32, "x")
    Cc = x.shape[-1]
    out = torch.empty((B, 2 * H, 2 * W, 2 * Cc), dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_upsample2x_split_f16, _ptr(x), B, H, W, Cc, _ptr(out))
    return out


def softmax_rows_f16(scores: Tensor, scale: float, out: Optional[Tensor] = None) -> Tensor:
    _chk(scores, torch.float32, "scores")
    M, N = scores.shape
    if out is None:
        out = torch.empty((M, N), dtype=torch.float16, device=scores.device)
    _call(scores.device, _lib.load().sgn_softmax_rows_f16, _ptr(scores), M, N, float(scale), _ptr(out))
    return out


def pointwise_nchw(x: Tensor, w_host: Tensor, b_host: Optional[Tensor], in_scale: float = 1.0) -> Tensor:
    """1x1 conv over <= 16 channels; x fp32 NCHW on the device, w [Cout,Cin] / bias [Cout] fp32 on the HOST."""
    _chk(x, torch.float32, "x")
    B, Cin, H, W = x.shape
    w = w_host.detach().to("cpu", torch.float32).contiguous()
    Cout = w.shape[0]
    b = None if b_host is None else b_host.detach().to("cpu", torch.float32).contiguous()
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    fp = C.POINTER(C.c_float)
    _call(x.device, _lib.load().sgn_pointwise_nchw, _ptr(x), C.cast(w.data_ptr(), fp),
          C.cast(b.data_ptr(), fp) if b is not None else None, B, Cin, Cout, H * W, float(in_scale), _ptr(out))
    return out


def vae_sample_latent(moments: Tensor, noise: Optional[Tensor], scale: float) -> Tensor:
    _chk(moments, torch.float32, "moments")
    _chk(noise, torch.float32, "noise")
    B, Z2, H, W = moments.shape
    out = torch.empty((B, Z2 // 2, H, W), dtype=torch.float32, device=moments.device)
    _call(moments.device, _lib.load().sgn_vae_sample_latent, _ptr(moments), _ptr(noise), B, Z2 // 2, H * W, float(scale),
          _ptr(out))
    return out


def u8_to_vae_input(img: Tensor) -> Tensor:
    """uint8 [H,W,3] -> fp32 [1,3,H,W] in [-1,1]."""
    _chk(img, torch.uint8, "img")
    H, W, _ = img.shape
    out = torch.empty((1, 3, H, W), dtype=torch.float32, device=img.device)
    _call(img.device, _lib.load().sgn_u8_to_vae_input, _ptr(img), H, W, _ptr(out))
    return out


def vae_output_to_u8(x: Tensor) -> Tensor:
    """fp32 [1,3,H,W] -> uint8 [H,W,3] = uint8(255 * clamp((x+1)/2, 0, 1))."""
    _chk(x, torch.float32, "x")
    _, _, H, W = x.shape
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=x.device)
    _call(x.device, _lib.load().sgn_vae_output_to_u8, _ptr(x), H, W, _ptr(out))
    return out


# ---------------------------------------------------------------------------------------------- parameter schema
def _res_schema(s: Dict[str, tuple], p: str, cin: int, cout: int) -> None:
    s[f"{p}.norm1.weight"], s[f"{p}.norm1.bias"] = (cin,), (cin,)
    s[f"{p}.conv1.weight"], s[f"{p}.conv1.bias"] = (cout, cin, 3, 3), (cout,)
    s[f"{p}.norm2.weight"], s[f"{p}.norm2.bias"] = (cout,), (cout,)
    s[f"{p}.conv2.weight"], s[f"{p}.conv2.bias"] = (cout, cout, 3, 3), (cout,)
    if cin != cout:
        s[f"{p}.nin_shortcut.weight"], s[f"{p}.nin_shortcut.bias"] = (cout, cin, 1, 1), (cout,)


def _mid_schema(s: Dict[str, tuple], p: str, c: int) -> None:
    _res_schema(s, f"{p}.block_1", c, c)
    s[f"{p}.attn_1.norm.weight"], s[f"{p}.attn_1.norm.bias"] = (c,), (c,)
    for n in ("q", "k", "v", "proj_out"):
        s[f"{p}.attn_1.{n}.weight"], s[f"{p}.attn_1.{n}.bias"] = (c, c, 1, 1), (c,)
    _res_schema(s, f"{p}.block_2", c, c)


def vae_param_schema(cfg: VAEConfig) -> "OrderedDict[str, tuple]":
    """name -> shape of every AutoencoderKL parameter, upstream (ldm / sgm) naming."""
    s: Dict[str, tuple] = OrderedDict()
    n = len(cfg.ch_mult)
    s["encoder.conv_in.weight"], s["encoder.conv_in.bias"] = (cfg.ch, cfg.in_channels, 3, 3), (cfg.ch,)
    in_mult = (1,) + tuple(cfg.ch_mult)
    cin = cfg.ch
    for i in range(n):
        cin, cout = cfg.ch * in_mult[i], cfg.ch * cfg.ch_mult[i]
        for j in range(cfg.num_res_blocks):
            _res_schema(s, f"encoder.down.{i}.block.{j}", cin, cout)
            cin = cout
        if i != n - 1:
            s[f"encoder.down.{i}.downsample.conv.weight"], s[f"encoder.down.{i}.downsample.conv.bias"] = (cin, cin, 3, 3), (cin,)
    _mid_schema(s, "encoder.mid", cin)
    zc = 2 * cfg.z_channels if cfg.double_z else cfg.z_channels
    s["encoder.norm_out.weight"], s["encoder.norm_out.bias"] = (cin,), (cin,)
    s["encoder.conv_out.weight"], s["encoder.conv_out.bias"] = (zc, cin, 3, 3), (zc,)
    cin = cfg.ch * cfg.ch_mult[-1]
    s["decoder.conv_in.weight"], s["decoder.conv_in.bias"] = (cin, cfg.z_channels, 3, 3), (cin,)
    _mid_schema(s, "decoder.mid", cin)
    for i in reversed(range(n)):
        cout = cfg.ch * cfg.ch_mult[i]
        for j in range(cfg.num_res_blocks + 1):
            _res_schema(s, f"decoder.up.{i}.block.{j}", cin, cout)
            cin = cout
        if i != 0:
            s[f"decoder.up.{i}.upsample.conv.weight"], s[f"decoder.up.{i}.upsample.conv.bias"] = (cin, cin, 3, 3), (cin,)
    s["decoder.norm_out.weight"], s["decoder.norm_out.bias"] = (cin,), (cin,)
    s["decoder.conv_out.weight"], s["decoder.conv_out.bias"] = (cfg.out_ch, cin, 3, 3), (cfg.out_ch,)
    s["quant_conv.weight"], s["quant_conv.bias"] = (zc, zc, 1, 1), (zc,)
    s["post_quant_conv.weight"], s["post_quant_conv.bias"] = (cfg.z_channels, cfg.z_channels, 1, 1), (cfg.z_channels,)
    return s


class VAERandomWeights(RandomWeights):
    """Random-init VAE parameters (torch default init family; GroupNorm affine = 1 / 0)."""

    def __getitem__(self, name: str) -> Tensor:
        if ".norm" in name:
            return (torch.ones if name.endswith("weight") else torch.zeros)(self.schema[name], device=self.device)
        return super().__getitem__(name)


# ---------------------------------------------------------------------------------------------- graph walk
class VAEB200:
    """AutoencoderKL.encode / decode with A1111's scaling conventions (see oracle/vae_ref.py for the fp32 restatement)."""

    ATTN_QUERY_CHUNK = 8192   # score slab = chunk x tokens fp32 (2 GB at 65 536 tokens)

    def __init__(self, cfg: VAEConfig, weights: Mapping[str, Tensor], device="cuda", scale_factor: float = SCALE_FACTOR,
                 exact: bool = True):
        """exact: every convolution / shortcut / down / upsample contraction takes its activation as fp16 hi + lo halves
        against the weights repeated along K (2x the tensor-core FLOPs): the fp32 operator to fp32 rounding for
        fp16-representable weights, as A1111 runs the SDXL VAE in fp32.  exact=False: single fp16 operand (1.5-2e-3
        relative L2 on the decoded image).  The mid-block attention always runs with fp16 q / k / v."""
        self.cfg, self.dev, self.scale_factor, self.exact = cfg, torch.device(device), scale_factor, bool(exact)
        if cfg.ch % 64 != 0:
            raise ValueError("the tensor-core conv needs channel counts that are multiples of 64")
        schema = vae_param_schema(cfg)
        missing = [n for n in schema if n not in weights]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} parameters, e.g. {missing[:3]}")
        for n, shp in schema.items():
            if tuple(weights[n].shape) != tuple(shp):
                raise ValueError(f"{n}: expected shape {tuple(shp)}, got {tuple(weights[n].shape)}")
        p = self.p = _Packed(weights, self.dev)
        for n, shp in schema.items():          # pack everything once
            if n.endswith("bias") or len(shp) == 1:
                p.f32(n)
            elif n in ("quant_conv.weight", "post_quant_conv.weight"):
                continue
            elif n in ("encoder.conv_in.weight", "decoder.conv_in.weight"):
                p.conv32(n)
            elif shp[2] == 3:
                (self._pack_lin if ".downsample.conv." in n else self._pack_conv)(n)   # the stride-2 conv is an im2col GEMM
            elif n.endswith(".q.weight"):      # q and k projections share one GEMM
                b = n[: -len("q.weight")]
                c = shp[0]
                p.t[b + "qk.weight"] = torch.cat([p._get(b + "q.weight").reshape(c, c), p._get(b + "k.weight").reshape(c, c)],
                                                 0).half().contiguous()
                p.t[b + "qk.bias"] = torch.cat([p._get(b + "q.bias"), p._get(b + "k.bias")]).contiguous()
            elif n.endswith(".k.weight"):
                continue
            elif n.endswith(".nin_shortcut.weight"):
                self._pack_lin(n)
            else:
                p.lin16(n)
        # the two 1x1 quant convs travel as kernel parameters: keep them on the host
        self._quant = (weights["quant_conv.weight"].detach().float().cpu().reshape(schema["quant_conv.weight"][:2]),
                       weights["quant_conv.bias"].detach().float().cpu())
        self._post_quant = (weights["post_quant_conv.weight"].detach().float().cpu().reshape(schema["post_quant_conv.weight"][:2]),
                            weights["post_quant_conv.bias"].detach().float().cpu())
        p.w = None

    # ------------------------------------------------------------------ packing (exact mode repeats the weights along K)
    def _pack_conv(self, name: str) -> Tensor:
        """[Co,Ci,3,3] -> fp16 [Co, 9*Ci] tap-major; exact: [Co, 9*2Ci] with every tap's channel block repeated."""
        key = name + ("#c16x2" if self.exact else "#c16")
        if key not in self.p.t:
            w = self.p._get(name).permute(0, 2, 3, 1)                       # [Co,3,3,Ci]
            if self.exact:
                w = torch.cat([w, w], dim=3)
            self.p.t[key] = w.reshape(w.shape[0], -1).half().contiguous()
        return self.p.t[key]

    def _pack_lin(self, name: str, taps: int = 1) -> Tensor:
        """1x1 conv / im2col GEMM weight [Co, K]; exact: [W | W]."""
        key = name + ("#l16x2" if self.exact else "#l16")
        if key not in self.p.t:
            w = self.p._get(name)
            w = (w.permute(0, 2, 3, 1) if w.dim() == 4 else w).reshape(w.shape[0], -1)
            self.p.t[key] = (torch.cat([w, w], 1) if self.exact else w).half().contiguous()
        return self.p.t[key]

    # ------------------------------------------------------------------ blocks
    def _gn(self, x: Act, name: str, act: bool, split: Optional[bool] = None) -> Tensor:
        return K.group_norm_f16(x.t, x.B, x.H * x.W, 32, 1e-6, self.p.f32(name + ".weight"), self.p.f32(name + ".bias"), act,
                                split=self.exact if split is None else split)

    def resblock(self, pre: str, x: Act) -> Act:
        p = self.p
        cout = p.f32(pre + ".conv1.bias").shape[0]
        a16 = self._gn(x, pre + ".norm1", True)
        h = K.conv3x3_f16(a16.view(x.B, x.H, x.W, -1), self._pack_conv(pre + ".conv1.weight"), p.f32(pre + ".conv1.bias"))
        b16 = self._gn(Act(h, x.B, x.H, x.W), pre + ".norm2", True)
        del a16, h
        if x.C != cout:
            x16 = split_f16(x.t) if self.exact else K.cast_f16(x.t)
            skip = K.gemm_f16(x16, self._pack_lin(pre + ".nin_shortcut.weight"), p.f32(pre + ".nin_shortcut.bias"))
            del x16
            out = K.conv3x3_f16(b16.view(x.B, x.H, x.W, -1), self._pack_conv(pre + ".conv2.weight"), p.f32(pre + ".conv2.bias"),
                                residual=skip, out=skip)
        else:
            out = K.conv3x3_f16(b16.view(x.B, x.H, x.W, -1), self._pack_conv(pre + ".conv2.weight"), p.f32(pre + ".conv2.bias"),
                                residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def attn(self, pre: str, x: Act) -> Act:
        """AttnBlock: x + proj_out(softmax(q k^T / sqrt(C)) v), one head over all H*W tokens."""
        p = self.p
        c, T = x.C, x.H * x.W
        h16 = self._gn(x, pre + ".norm", False, split=False)
        qk = K.gemm_f16(h16, p.t[pre + ".qk.weight"], p.t[pre + ".qk.bias"], out_f16=True)  # [B*T, 2C]
        o16 = torch.empty((x.B * T, c), dtype=torch.float16, device=self.dev)
        if T % 8 != 0:
            raise ValueError("AttnBlock needs H*W % 8 == 0 (row strides of the score / probability matrices)")
        for b in range(x.B):
            rows = slice(b * T, (b + 1) * T)
            # V^T [C, T] directly from the GEMM (W_v . h^T); its bias is added after P.V (rows of P sum to 1)
            vT = K.gemm_f16(p.lin16(pre + ".v.weight"), h16[rows], None, out_f16=True)
            for q0 in range(0, T, self.ATTN_QUERY_CHUNK):
                q1 = min(T, q0 + self.ATTN_QUERY_CHUNK)
                s = K.gemm_f16(qk[b * T + q0:b * T + q1, :c], qk[rows, c:], None)               # fp32 [chunk, T]
                pr = softmax_rows_f16(s, float(c) ** -0.5)
                del s
                K.gemm_f16(pr, vT, p.f32(pre + ".v.bias"), out_f16=True, out=o16[b * T + q0:b * T + q1])
                del pr
        out = K.gemm_f16(o16, p.lin16(pre + ".proj_out.weight"), p.f32(pre + ".proj_out.bias"), residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def mid(self, pre: str, x: Act) -> Act:
        return self.resblock(pre + ".block_2", self.attn(pre + ".attn_1", self.resblock(pre + ".block_1", x)))

    def _conv_in(self, name: str, x_nchw: Tensor) -> Act:
        B, _, H, W = x_nchw.shape
        t = K.conv3x3_direct(x_nchw, True, self.p.conv32(name + ".weight"), self.p.f32(name + ".bias"))
        return Act(t.view(B * H * W, -1), B, H, W)

    def _conv_out(self, side: str, h: Act) -> Tensor:
        a16 = self._gn(h, side + ".norm_out", True)
        return K.conv3x3_f16(a16.view(h.B, h.H, h.W, -1), self._pack_conv(side + ".conv_out.weight"),
                             self.p.f32(side + ".conv_out.bias"), nchw=True)

    # ------------------------------------------------------------------ public
    def moments(self, x: Tensor) -> Tensor:
        """x fp32 [B,3,H,W] in [-1,1] (H, W multiples of 8) -> quant_conv(encoder(x)) [B,2z,H/8,W/8]."""
        _chk(x, torch.float32, "x")
        cfg, p = self.cfg, self.p
        h = self._conv_in("encoder.conv_in", x)
        n = len(cfg.ch_mult)
        for i in range(n):
            for j in range(cfg.num_res_blocks):
                h = self.resblock(f"encoder.down.{i}.block.{j}", h)
            if i != n - 1:
                col, ho, wo = im2col3x3_s2_asym_f16(h.t, h.B, h.H, h.W, split=self.exact)
                pre = f"encoder.down.{i}.downsample.conv"
                h = Act(K.gemm_f16(col, self._pack_lin(pre + ".weight"), p.f32(pre + ".bias")), h.B, ho, wo)
                del col
        h = self.mid("encoder.mid", h)
        return pointwise_nchw(self._conv_out("encoder", h), *self._quant)

    def encode(self, x: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """A1111 init_latent: scale_factor * DiagonalGaussianDistribution(moments).sample() (noise None = the mode)."""
        return vae_sample_latent(self.moments(x), noise, self.scale_factor)

    def decode(self, z: Tensor) -> Tensor:
        """z fp32 [B,z,h,w] (scaled latent) -> decoder(post_quant_conv(z / scale_factor)) fp32 [B,3,8h,8w]."""
        _chk(z, torch.float32, "z")
        cfg, p = self.cfg, self.p
        h = self._conv_in("decoder.conv_in", pointwise_nchw(z, *self._post_quant, in_scale=1.0 / self.scale_factor))
        h = self.mid("decoder.mid", h)
        for i in reversed(range(len(cfg.ch_mult))):
            for j in range(cfg.num_res_blocks + 1):
                h = self.resblock(f"decoder.up.{i}.block.{j}", h)
            if i != 0:
                pre = f"decoder.up.{i}.upsample.conv"
                u16 = upsample2x_split_f16(h.t, h.B, h.H, h.W) if self.exact else K.upsample2x_f16(h.t, h.B, h.H, h.W)
                h = Act(K.conv3x3_f16(u16, self._pack_conv(pre + ".weight"), p.f32(pre + ".bias")), h.B, 2 * h.H, 2 * h.W)
                del u16
        return self._conv_out("decoder", h)
