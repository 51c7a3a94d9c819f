"""ORACLE (test infrastructure, not product code): plain-PyTorch fp32 restatement of the denoiser the reference
reaches over HTTP — SDXL `UNetModel` + ControlNet-XL "depth full" + one A1111 CFG / Euler-ancestral sampler step.

PARITY UNPINNED: the arithmetic lives in third-party code that is NOT under /root/reference and cannot be
imported or installed here (no network):
  * Stability-AI/generative-models  `sgm/modules/diffusionmodules/openaimodel.py` (UNetModel, ResBlock,
    Downsample, Upsample, timestep_embedding), `sgm/modules/attention.py` (SpatialTransformer,
    BasicTransformerBlock, CrossAttention, FeedForward/GEGLU) — as vendored by
    AUTOMATIC1111/stable-diffusion-webui @ 5ef669de080814067961f28357256e8fe27544f4 (reference README.md:49);
  * Mikubill/sd-webui-controlnet @ a43e574254d19a362082bbd412f24aeef1beed47 (README.md:60) `scripts/cldm.py`
    (ControlNet), `scripts/hook.py` (residual injection: mid `h += control`, decoder `cat([h, hs.pop()+control.pop()])`);
  * k-diffusion `sample_euler_ancestral`, `CompVisDenoiser` and A1111 `modules/sd_samplers_cfg_denoiser.py`.
The reference's own call site is signerf/diffuser/diffuser.py:132-180 (payload: steps 20, cfg 7, denoising 0.9,
"Euler a", ControlNet weight 0.8 / Balanced / module none).  This file restates the published algorithms with the
upstream module / parameter names so that a real SDXL / ControlNet state_dict loads key-for-key
(`model.diffusion_model.` / `control_model.` prefixes stripped).  The reference ships no tests or golden vectors
for this path, so the restatement is anchored on those call sites only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn


@dataclass
class UNetConfig:
    """sgm UNetModel hyper-parameters; defaults = SDXL base 1.0 (sd_xl_base.yaml `network_config`)."""
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    channel_mult: Tuple[int, ...] = (1, 2, 4)
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (4, 2)
    transformer_depth: Tuple[int, ...] = (1, 2, 10)
    num_head_channels: int = 64
    context_dim: int = 2048
    adm_in_channels: int = 2816
    hint_channels: int = 3   # ControlNet only


def tiny_config(**kw) -> UNetConfig:
    """Same topology at test width: 64/128/256 channels, transformer depth (1, 1, 2)."""
    d = dict(model_channels=64, transformer_depth=(1, 1, 2), context_dim=128, adm_in_channels=96)
    d.update(kw)
    return UNetConfig(**d)


def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class ResBlock(nn.Module):
    def __init__(self, channels: int, emb_channels: int, out_channels: int):
        super().__init__()
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(), nn.Conv2d(channels, out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_channels), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(out_channels, out_channels, 3, padding=1))
        self.skip_connection = nn.Identity() if out_channels == channels else nn.Conv2d(channels, out_channels, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb)[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class Downsample(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.op = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class CrossAttention(nn.Module):
    def __init__(self, query_dim: int, context_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        context = x if context is None else context
        b, n, _ = x.shape
        h = self.heads
        q = self.to_q(x).view(b, n, h, -1).transpose(1, 2)
        k = self.to_k(context).view(b, context.shape[1], h, -1).transpose(1, 2)
        v = self.to_v(context).view(b, context.shape[1], h, -1).transpose(1, 2)
        out = _attention(q, k, v, self.scale)
        return self.to_out(out.transpose(1, 2).reshape(b, n, -1))


def _attention(q, k, v, scale, chunk: int = 2048):
    """softmax(q k^T * scale) v, query-chunked so the 16 384-token sheet fits in memory."""
    outs = []
    for i in range(0, q.shape[2], chunk):
        s = torch.matmul(q[:, :, i:i + chunk], k.transpose(-1, -2)) * scale
        outs.append(torch.matmul(torch.softmax(s, dim=-1), v))
    return torch.cat(outs, dim=2)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, context_dim: int):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads, dim_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class SpatialTransformer(nn.Module):
    """use_linear_in_transformer=True variant (SDXL)."""

    def __init__(self, channels: int, heads: int, dim_head: int, depth: int, context_dim: int):
        super().__init__()
        self.norm = nn.GroupNorm(32, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, dim_head, context_dim)
                                                 for _ in range(depth)])
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, x, context):
        b, c, h, w = x.shape
        x_in = x
        x = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        x = self.proj_out(x)
        return x.reshape(b, h, w, c).permute(0, 3, 1, 2) + x_in


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, x, emb, context):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


def _encoder(cfg: UNetConfig):
    """input_blocks + middle_block shared by UNetModel and ControlNet; returns (input_blocks, chans, middle, ch, ds)."""
    mc, ted = cfg.model_channels, cfg.model_channels * 4
    blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(cfg.in_channels, mc, 3, padding=1))])
    chans, ch, ds = [mc], mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers: List[nn.Module] = [ResBlock(ch, ted, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(SpatialTransformer(ch, ch // cfg.num_head_channels, cfg.num_head_channels,
                                                 cfg.transformer_depth[level], cfg.context_dim))
            blocks.append(TimestepEmbedSequential(*layers))
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            blocks.append(TimestepEmbedSequential(Downsample(ch)))
            chans.append(ch)
            ds *= 2
    middle = TimestepEmbedSequential(
        ResBlock(ch, ted, ch),
        SpatialTransformer(ch, ch // cfg.num_head_channels, cfg.num_head_channels, cfg.transformer_depth[-1], cfg.context_dim),
        ResBlock(ch, ted, ch))
    return blocks, chans, middle, ch, ds


def _embedders(cfg: UNetConfig):
    mc, ted = cfg.model_channels, cfg.model_channels * 4
    time_embed = nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
    label_emb = nn.Sequential(nn.Sequential(nn.Linear(cfg.adm_in_channels, ted), nn.SiLU(), nn.Linear(ted, ted)))
    return time_embed, label_emb


class UNetModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        mc, ted = cfg.model_channels, cfg.model_channels * 4
        self.time_embed, self.label_emb = _embedders(cfg)
        self.input_blocks, chans, self.middle_block, ch, ds = _encoder(cfg)
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
            for i in range(cfg.num_res_blocks + 1):
                ich = chans.pop()
                layers: List[nn.Module] = [ResBlock(ch + ich, ted, mc * mult)]
                ch = mc * mult
                if ds in cfg.attention_resolutions:
                    layers.append(SpatialTransformer(ch, ch // cfg.num_head_channels, cfg.num_head_channels,
                                                     cfg.transformer_depth[level], cfg.context_dim))
                if level and i == cfg.num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(mc, cfg.out_channels, 3, padding=1))

    def forward(self, x, timesteps, context, y, control: Optional[Sequence[Tensor]] = None, control_weight: float = 1.0,
                taps: Optional[dict] = None):
        """`control`: the ControlNet's 9 encoder residuals + 1 middle residual, injected as sd-webui-controlnet's
        hook does.  `taps` (dict) collects per-block activations for parity tests."""
        emb = self.time_embed(timestep_embedding(timesteps, self.cfg.model_channels)) + self.label_emb(y)
        control = list(control) if control is not None else None
        hs, h = [], x
        for i, m in enumerate(self.input_blocks):
            h = m(h, emb, context)
            hs.append(h)
            if taps is not None:
                taps[f"input_blocks.{i}"] = h
        h = self.middle_block(h, emb, context)
        if control is not None:
            h = h + control_weight * control.pop()
        if taps is not None:
            taps["middle_block"] = h
        for i, m in enumerate(self.output_blocks):
            skip = hs.pop()
            if control is not None:
                skip = skip + control_weight * control.pop()
            h = m(torch.cat([h, skip], dim=1), emb, context)
            if taps is not None:
                taps[f"output_blocks.{i}"] = h
        return self.out(h)


class ControlNet(nn.Module):
    """cldm.ControlNet for SDXL ("diffusers_xl_depth_full"): encoder copy + hint stem + zero convs."""

    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        mc = cfg.model_channels
        self.time_embed, self.label_emb = _embedders(cfg)
        self.input_blocks, chans, self.middle_block, ch, _ = _encoder(cfg)
        self.zero_convs = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(c, c, 1)) for c in chans])
        self.middle_block_out = TimestepEmbedSequential(nn.Conv2d(ch, ch, 1))
        hc = cfg.hint_channels
        self.input_hint_block = TimestepEmbedSequential(
            nn.Conv2d(hc, 16, 3, padding=1), nn.SiLU(), nn.Conv2d(16, 16, 3, padding=1), nn.SiLU(),
            nn.Conv2d(16, 32, 3, padding=1, stride=2), nn.SiLU(), nn.Conv2d(32, 32, 3, padding=1), nn.SiLU(),
            nn.Conv2d(32, 96, 3, padding=1, stride=2), nn.SiLU(), nn.Conv2d(96, 96, 3, padding=1), nn.SiLU(),
            nn.Conv2d(96, 256, 3, padding=1, stride=2), nn.SiLU(), nn.Conv2d(256, mc, 3, padding=1))

    def forward(self, x, hint, timesteps, context, y) -> List[Tensor]:
        emb = self.time_embed(timestep_embedding(timesteps, self.cfg.model_channels)) + self.label_emb(y)
        guided_hint = self.input_hint_block(hint, emb, context)
        outs, h = [], x
        for module, zero_conv in zip(self.input_blocks, self.zero_convs):
            h = module(h, emb, context)
            if guided_hint is not None:
                h = h + guided_hint
                guided_hint = None
            outs.append(zero_conv(h, emb, context))
        h = self.middle_block(h, emb, context)
        outs.append(self.middle_block_out(h, emb, context))
        return outs


def make_models(cfg: UNetConfig, seed: int = 0, device="cpu", fp16_weights: bool = True,
                fast_init: bool = False) -> Tuple[UNetModel, ControlNet]:
    """Random-init denoiser.  sgm / cldm zero-initialise (`zero_module`) the ResBlock output convs, proj_out, the UNet
    output conv and every ControlNet zero conv, which would make the random-weight network trivial; like every other
    layer they keep torch's default init here so that each branch is exercised.

    fp16_weights: round every parameter to the nearest fp16 value (kept as fp32 tensors).  The checkpoints the reference
    names (`sd_xl_base_1.0.safetensors`, `diffusers_xl_depth_full`, diffuser.py:47-49) are stored in fp16 and A1111's
    `--no-half` only up-casts them, so the reference's fp32 arithmetic runs on fp16-representable weights."""
    torch.manual_seed(seed)
    if fast_init:
        # full-width models for CPU timing only: skip torch's per-parameter init (minutes for 3.8 B parameters) and tile
        # one block of random values over every tensor, scaled like the default init (values do not affect timing)
        with torch.device("meta"):
            unet, ctrl = UNetModel(cfg), ControlNet(cfg)
        block = torch.rand(1 << 20) * 2 - 1
        for m in (unet, ctrl):
            m.to_empty(device="cpu")
            with torch.no_grad():
                for name, prm in m.named_parameters():
                    if prm.dim() == 1:
                        prm.fill_(1.0 if name.endswith("weight") and ("norm" in name or "layers.0" in name or name.startswith("out.0")) else 0.0)
                        continue
                    flat = prm.view(-1)
                    bound = 1.0 / math.sqrt(prm[0].numel())
                    for i in range(0, flat.numel(), block.numel()):
                        n = min(block.numel(), flat.numel() - i)
                        torch.mul(block[:n], bound, out=flat[i:i + n])
        return unet.eval(), ctrl.eval()
    unet, ctrl = UNetModel(cfg), ControlNet(cfg)
    if fp16_weights:
        with torch.no_grad():
            for m in (unet, ctrl):
                for prm in m.parameters():
                    prm.copy_(prm.half().float())
    return unet.to(device).eval(), ctrl.to(device).eval()


# ---------------------------------------------------------------------------------------------- sampler (A13)
def sdxl_sigmas(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> Tensor:
    """k-diffusion DiscreteSchedule sigmas of SDXL's scaled-linear betas: sigma_t = sqrt((1-acp)/acp)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    acp = torch.cumprod(1.0 - betas, dim=0)
    return ((1 - acp) / acp).sqrt().float()


def get_sigmas(all_sigmas: Tensor, n: int) -> Tensor:
    """DiscreteSchedule.get_sigmas(n): n sigmas from t_max to 0 (log-linear interpolation of the table) + final 0."""
    t = torch.linspace(len(all_sigmas) - 1, 0, n)
    log_s = all_sigmas.log()
    lo, hi, w = t.floor().long(), t.ceil().long(), t.frac()
    return torch.cat([((1 - w) * log_s[lo] + w * log_s[hi]).exp(), torch.zeros(1)])


def sigma_to_t(all_sigmas: Tensor, sigma: Tensor) -> Tensor:
    """DiscreteSchedule.sigma_to_t (quantize=False): fractional timestep by log-sigma interpolation."""
    log_s = all_sigmas.log()
    dists = sigma.log()[None] - log_s[:, None]
    low = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=len(all_sigmas) - 2)
    high = low + 1
    w = ((log_s[low] - sigma.log()) / (log_s[low] - log_s[high])).clamp(0, 1)
    return ((1 - w) * low + w * high).view(sigma.shape)


def img2img_schedule(steps: int = 20, denoising_strength: float = 0.9) -> Tensor:
    """A1111 img2img: t_enc = int(min(strength, 0.999) * steps); sigmas[steps - t_enc - 1:]  (t_enc UNet calls)."""
    t_enc = int(min(denoising_strength, 0.999) * steps)
    return get_sigmas(sdxl_sigmas(), steps)[steps - t_enc - 1:]


def ancestral_step(sigma_from: float, sigma_to: float, eta: float = 1.0) -> Tuple[float, float]:
    """k-diffusion get_ancestral_step -> (sigma_down, sigma_up)."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    return (sigma_to ** 2 - sigma_up ** 2) ** 0.5, sigma_up


@torch.no_grad()
def denoise_step(unet: UNetModel, ctrl: Optional[ControlNet], x: Tensor, sigma: float, sigma_next: float, context: Tensor,
                 y: Tensor, hint: Optional[Tensor], noise: Optional[Tensor], init_latent: Optional[Tensor] = None,
                 mask: Optional[Tensor] = None, cfg_scale: float = 7.0, control_weight: float = 0.8):
    """One A1111 "Euler a" step with CFG and ControlNet.  x [B,4,h,w]; context [2B,77,D] / y [2B,adm] are stacked
    (cond, uncond); hint [B,3,8h,8w] in [0,1]; mask [B,1,h,w] = 1 where the ORIGINAL latent is kept.
    Returns (x_next, denoised, eps)."""
    all_sigmas = sdxl_sigmas().to(x.device)
    s = torch.tensor([sigma], device=x.device)
    t = sigma_to_t(all_sigmas, s).repeat(2 * x.shape[0])
    c_in = 1.0 / (sigma ** 2 + 1.0) ** 0.5
    xin = torch.cat([x, x]) * c_in
    control = ctrl(xin, torch.cat([hint, hint]), t, context, y) if ctrl is not None else None
    eps = unet(xin, t, context, y, control=control, control_weight=control_weight)
    eps_c, eps_u = eps.chunk(2)
    e = eps_u + cfg_scale * (eps_c - eps_u)
    denoised = x - sigma * e
    if mask is not None:
        denoised = init_latent * mask + (1.0 - mask) * denoised
    sigma_down, sigma_up = ancestral_step(sigma, sigma_next)
    d = (x - denoised) / sigma
    x_next = x + d * (sigma_down - sigma)
    if noise is not None and sigma_next > 0:
        x_next = x_next + noise * sigma_up
    return x_next, denoised, eps
