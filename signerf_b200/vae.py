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
def im2col3x3_s2_asym_f16(x: Tensor, B: int, H: int, W: int):
    """fp32 [B*H*W, C] -> (fp16 [B*Ho*Wo, 9C], Ho, Wo) for ldm Downsample (pad (0,1,0,1), 3x3, stride 2)."""
    _chk(x, torch.float32, "x")
    Cc = x.shape[-1]
    Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    out = torch.empty((B * Ho * Wo, 9 * Cc), dtype=torch.float16, device=x.device)
    _call(x.device, _lib.load().sgn_im2col3x3_s2_asym_f16, _ptr(x), B, H, W, Cc, _ptr(out))
    return out, Ho, Wo


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

    def __init__(self, cfg: VAEConfig, weights: Mapping[str, Tensor], device="cuda", scale_factor: float = SCALE_FACTOR):
        self.cfg, self.dev, self.scale_factor = cfg, torch.device(device), scale_factor
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
                p.conv16(n)
            elif n.endswith(".q.weight"):      # q and k projections share one GEMM
                b = n[: -len("q.weight")]
                c = shp[0]
                p.t[b + "qk.weight"] = torch.cat([p._get(b + "q.weight").reshape(c, c), p._get(b + "k.weight").reshape(c, c)],
                                                 0).half().contiguous()
                p.t[b + "qk.bias"] = torch.cat([p._get(b + "q.bias"), p._get(b + "k.bias")]).contiguous()
            elif n.endswith(".k.weight"):
                continue
            else:
                p.lin16(n)
        # the two 1x1 quant convs travel as kernel parameters: keep them on the host
        self._quant = (weights["quant_conv.weight"].detach().float().cpu().reshape(schema["quant_conv.weight"][:2]),
                       weights["quant_conv.bias"].detach().float().cpu())
        self._post_quant = (weights["post_quant_conv.weight"].detach().float().cpu().reshape(schema["post_quant_conv.weight"][:2]),
                            weights["post_quant_conv.bias"].detach().float().cpu())
        p.w = None

    # ------------------------------------------------------------------ blocks
    def _gn(self, x: Act, name: str, act: bool) -> Tensor:
        return K.group_norm_f16(x.t, x.B, x.H * x.W, 32, 1e-6, self.p.f32(name + ".weight"), self.p.f32(name + ".bias"), act)

    def resblock(self, pre: str, x: Act) -> Act:
        p = self.p
        cout = p.f32(pre + ".conv1.bias").shape[0]
        a16 = self._gn(x, pre + ".norm1", True)
        h = K.conv3x3_f16(a16.view(x.B, x.H, x.W, x.C), p.conv16(pre + ".conv1.weight"), p.f32(pre + ".conv1.bias"))
        b16 = self._gn(Act(h, x.B, x.H, x.W), pre + ".norm2", True)
        del a16, h
        if x.C != cout:
            skip = K.gemm_f16(K.cast_f16(x.t), p.lin16(pre + ".nin_shortcut.weight"), p.f32(pre + ".nin_shortcut.bias"))
            out = K.conv3x3_f16(b16.view(x.B, x.H, x.W, cout), p.conv16(pre + ".conv2.weight"), p.f32(pre + ".conv2.bias"),
                                residual=skip, out=skip)
        else:
            out = K.conv3x3_f16(b16.view(x.B, x.H, x.W, cout), p.conv16(pre + ".conv2.weight"), p.f32(pre + ".conv2.bias"),
                                residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def attn(self, pre: str, x: Act) -> Act:
        """AttnBlock: x + proj_out(softmax(q k^T / sqrt(C)) v), one head over all H*W tokens."""
        p = self.p
        c, T = x.C, x.H * x.W
        h16 = self._gn(x, pre + ".norm", False)
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
        return K.conv3x3_f16(a16.view(h.B, h.H, h.W, h.C), self.p.conv16(side + ".conv_out.weight"),
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
                col, ho, wo = im2col3x3_s2_asym_f16(h.t, h.B, h.H, h.W)
                pre = f"encoder.down.{i}.downsample.conv"
                h = Act(K.gemm_f16(col, p.conv16(pre + ".weight"), p.f32(pre + ".bias")), h.B, ho, wo)
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
                u16 = K.upsample2x_f16(h.t, h.B, h.H, h.W)
                h = Act(K.conv3x3_f16(u16, p.conv16(pre + ".weight"), p.f32(pre + ".bias")), h.B, 2 * h.H, 2 * h.W)
                del u16
        return self._conv_out("decoder", h)
