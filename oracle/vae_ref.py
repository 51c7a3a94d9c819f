"""ORACLE (test infrastructure, NOT product code) — fp32 torch restatement of the SDXL first-stage autoencoder the
A1111 server runs around the denoising loop of the reference's img2img request (signerf/diffuser/diffuser.py:180;
SURVEY §8(f) row 1, Appendix C items 1 and 4): sgm `AutoencodingEngine` / ldm `AutoencoderKL` with the
`Encoder` / `Decoder` of sgm/modules/diffusionmodules/model.py (ResnetBlock, AttnBlock, Downsample with the asymmetric
(0,1,0,1) pad, nearest Upsample + conv), `quant_conv` / `post_quant_conv`, the diagonal-Gaussian posterior
(logvar clamped to [-30, 20], z = mean + std * noise) and SDXL's scale_factor 0.13025.

parity unpinned: A1111 @5ef669de and its generative-models checkout are external and not installable here (no
wheels, no network); module / parameter names follow upstream (`encoder.down.0.block.0.norm1.weight`, ...,
`decoder.up.3.upsample.conv.weight`, `quant_conv.weight`) so that `sdxl_vae.safetensors` / the first_stage_model.* part
of `sd_xl_base_1.0.safetensors` loads key for key; the SDXL-VAE parameter count (83 653 863) is asserted in the tests.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

SCALE_FACTOR = 0.13025   # sd_xl_base.yaml: scale_factor


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


def tiny_vae_config() -> VAEConfig:
    """Same topology (3 downsamples, mid attention), test width."""
    return VAEConfig(ch=64, ch_mult=(1, 2, 2, 2))


def Normalize(c: int) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


class ResnetBlock(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = Normalize(cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2 = Normalize(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1, 1, 0)

    def forward(self, x: Tensor) -> Tensor:
        h = self.conv1(swish(self.norm1(x)))
        h = self.conv2(swish(self.norm2(h)))       # dropout 0.0
        if hasattr(self, "nin_shortcut"):
            x = self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    """Single-head self-attention over all pixels, head dim = channels."""

    def __init__(self, c: int):
        super().__init__()
        self.norm = Normalize(c)
        self.q = nn.Conv2d(c, c, 1)
        self.k = nn.Conv2d(c, c, 1)
        self.v = nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)

    def forward(self, x: Tensor) -> Tensor:
        h = self.norm(x)
        q, k, v = self.q(h), self.k(h), self.v(h)
        b, c, hh, ww = q.shape
        q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
        k = k.reshape(b, c, hh * ww)
        w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
        w_ = F.softmax(w_, dim=2)
        v = v.reshape(b, c, hh * ww)
        h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.proj_out(h)


class Downsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 2, 0)

    def forward(self, x: Tensor) -> Tensor:
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class Upsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x: Tensor) -> Tensor:
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Mid(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.block_1 = ResnetBlock(c, c)
        self.attn_1 = AttnBlock(c)
        self.block_2 = ResnetBlock(c, c)

    def forward(self, x: Tensor) -> Tensor:
        return self.block_2(self.attn_1(self.block_1(x)))


class _Level(nn.Module):
    pass


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.conv_in = nn.Conv2d(cfg.in_channels, cfg.ch, 3, 1, 1)
        in_mult = (1,) + tuple(cfg.ch_mult)
        self.down = nn.ModuleList()
        for i in range(len(cfg.ch_mult)):
            lvl = _Level()
            cin, cout = cfg.ch * in_mult[i], cfg.ch * cfg.ch_mult[i]
            lvl.block = nn.ModuleList()
            for _ in range(cfg.num_res_blocks):
                lvl.block.append(ResnetBlock(cin, cout))
                cin = cout
            lvl.attn = nn.ModuleList()
            if i != len(cfg.ch_mult) - 1:
                lvl.downsample = Downsample(cin)
            self.down.append(lvl)
        self.mid = _Mid(cin)
        self.norm_out = Normalize(cin)
        self.conv_out = nn.Conv2d(cin, 2 * cfg.z_channels if cfg.double_z else cfg.z_channels, 3, 1, 1)

    def forward(self, x: Tensor) -> Tensor:
        h = self.conv_in(x)
        for i, lvl in enumerate(self.down):
            for blk in lvl.block:
                h = blk(h)
            if i != len(self.down) - 1:
                h = lvl.downsample(h)
        h = self.mid(h)
        return self.conv_out(swish(self.norm_out(h)))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        n = len(cfg.ch_mult)
        cin = cfg.ch * cfg.ch_mult[-1]
        self.conv_in = nn.Conv2d(cfg.z_channels, cin, 3, 1, 1)
        self.mid = _Mid(cin)
        self.up = nn.ModuleList()
        for i in reversed(range(n)):
            lvl = _Level()
            cout = cfg.ch * cfg.ch_mult[i]
            lvl.block = nn.ModuleList()
            for _ in range(cfg.num_res_blocks + 1):
                lvl.block.append(ResnetBlock(cin, cout))
                cin = cout
            lvl.attn = nn.ModuleList()
            if i != 0:
                lvl.upsample = Upsample(cin)
            self.up.insert(0, lvl)
        self.norm_out = Normalize(cin)
        self.conv_out = nn.Conv2d(cin, cfg.out_ch, 3, 1, 1)

    def forward(self, z: Tensor) -> Tensor:
        h = self.mid(self.conv_in(z))
        for i in reversed(range(len(self.up))):
            for blk in self.up[i].block:
                h = blk(h)
            if i != 0:
                h = self.up[i].upsample(h)
        return self.conv_out(swish(self.norm_out(h)))


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: Optional[VAEConfig] = None):
        super().__init__()
        self.cfg = cfg or VAEConfig()
        zc = self.cfg.z_channels
        self.encoder = Encoder(self.cfg)
        self.decoder = Decoder(self.cfg)
        self.quant_conv = nn.Conv2d(2 * zc, 2 * zc, 1)
        self.post_quant_conv = nn.Conv2d(zc, zc, 1)

    def moments(self, x: Tensor) -> Tensor:
        """x [B,3,H,W] in [-1,1] -> quant_conv(encoder(x)) [B,2z,H/8,W/8] = (mean | logvar)."""
        return self.quant_conv(self.encoder(x))

    def encode(self, x: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """A1111 `get_first_stage_encoding(encode_first_stage(x))`: DiagonalGaussianDistribution.sample() * scale_factor
        (noise None = the posterior mode)."""
        mean, logvar = torch.chunk(self.moments(x), 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        z = mean if noise is None else mean + torch.exp(0.5 * logvar) * noise
        return SCALE_FACTOR * z

    def decode(self, z: Tensor) -> Tensor:
        """A1111 decode_first_stage: z / scale_factor -> post_quant_conv -> decoder; [B,3,H,W] roughly in [-1,1]."""
        return self.decoder(self.post_quant_conv(z / SCALE_FACTOR))


def make_vae(cfg: Optional[VAEConfig] = None, seed: int = 0, device="cpu", fp16_weights: bool = True) -> AutoencoderKL:
    """Random-init autoencoder (torch default inits, seeded).  fp16_weights: round every parameter to an
    fp16-representable value, as the checkpoints the reference's server loads are (fp16 safetensors up-cast)."""
    torch.manual_seed(seed)
    m = AutoencoderKL(cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:          # exercise the affine part of every GroupNorm
                p.copy_(1.0 + 0.1 * torch.randn_like(p) if n.endswith("weight") else 0.1 * torch.randn_like(p))
            if fp16_weights:
                p.copy_(p.half().float())
    for p in m.parameters():
        p.requires_grad_(False)
    return m.to(device)
