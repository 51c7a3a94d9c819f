"""In-process SDXL + ControlNet-depth denoiser on B200: the arithmetic behind reference `Diffuser.diffuse`
(signerf/diffuser/diffuser.py:92-195), which the reference reaches over HTTP on an A1111 SD-WebUI server
(README.md:38-60).  This module is the host side of SURVEY §8a rows A12/A13: it walks the sgm `UNetModel` /
cldm `ControlNet` graphs and issues one C-ABI call per operator (include/signerf_b200.h K5-K9):

    tcgen05 GEMM / implicit-GEMM conv  (sgn_gemm_f16, sgn_conv3x3_f16)     every Linear / Conv2d with Cin % 64 == 0
    tcgen05 flash attention            (sgn_attention_f16)                 attn1 (self) / attn2 (77-token context)
    GroupNorm+SiLU, LayerNorm, casts, nearest-upsample, concat(+ControlNet residual), stride-2 im2col,
    direct small-channel convs, embedding linears, fused CFG / inpaint-blend / Euler-ancestral update.

Weights arrive as a state_dict with the upstream parameter names (`input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight`
...; a real `sd_xl_base_1.0` / `diffusers_xl_depth_full` checkpoint loads after stripping its `model.diffusion_model.` /
`control_model.` prefix) and are packed once into the fp16 operand layouts the kernels want.  Activations are NHWC:
the residual stream is fp32 [B*H*W, C]; every GEMM operand is fp16 with fp32 accumulation in TMEM.
There is no torch math on the compute path and no CPU fallback; torch provides memory, streams and CUDA graphs.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from . import nn_ops as K


@dataclass
class UNetConfig:
    """sgm UNetModel hyper-parameters; defaults = SDXL base 1.0."""
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
    hint_channels: int = 3


SMALL_TC_SHAPES = ((16, 16), (16, 32), (32, 32))   # (Cin, Cout) of sgn_conv3x3_small_tc
HINT_STACK = ((16, 1), (16, 1), (32, 2), (32, 1), (96, 2), (96, 1), (256, 2))  # (out channels, stride) before the 256->mc conv


# ---------------------------------------------------------------------------------------------- parameter schema
def _res_schema(s: Dict[str, tuple], p: str, cin: int, cout: int, ted: int) -> None:
    s[f"{p}.in_layers.0.weight"] = (cin,)
    s[f"{p}.in_layers.0.bias"] = (cin,)
    s[f"{p}.in_layers.2.weight"] = (cout, cin, 3, 3)
    s[f"{p}.in_layers.2.bias"] = (cout,)
    s[f"{p}.emb_layers.1.weight"] = (cout, ted)
    s[f"{p}.emb_layers.1.bias"] = (cout,)
    s[f"{p}.out_layers.0.weight"] = (cout,)
    s[f"{p}.out_layers.0.bias"] = (cout,)
    s[f"{p}.out_layers.3.weight"] = (cout, cout, 3, 3)
    s[f"{p}.out_layers.3.bias"] = (cout,)
    if cin != cout:
        s[f"{p}.skip_connection.weight"] = (cout, cin, 1, 1)
        s[f"{p}.skip_connection.bias"] = (cout,)


def _st_schema(s: Dict[str, tuple], p: str, c: int, depth: int, ctx: int) -> None:
    s[f"{p}.norm.weight"] = (c,)
    s[f"{p}.norm.bias"] = (c,)
    s[f"{p}.proj_in.weight"] = (c, c)
    s[f"{p}.proj_in.bias"] = (c,)
    for d in range(depth):
        b = f"{p}.transformer_blocks.{d}"
        for a, kd in (("attn1", c), ("attn2", ctx)):
            s[f"{b}.{a}.to_q.weight"] = (c, c)
            s[f"{b}.{a}.to_k.weight"] = (c, kd)
            s[f"{b}.{a}.to_v.weight"] = (c, kd)
            s[f"{b}.{a}.to_out.0.weight"] = (c, c)
            s[f"{b}.{a}.to_out.0.bias"] = (c,)
        s[f"{b}.ff.net.0.proj.weight"] = (8 * c, c)
        s[f"{b}.ff.net.0.proj.bias"] = (8 * c,)
        s[f"{b}.ff.net.2.weight"] = (c, 4 * c)
        s[f"{b}.ff.net.2.bias"] = (c,)
        for n in ("norm1", "norm2", "norm3"):
            s[f"{b}.{n}.weight"] = (c,)
            s[f"{b}.{n}.bias"] = (c,)
    s[f"{p}.proj_out.weight"] = (c, c)
    s[f"{p}.proj_out.bias"] = (c,)


def _encoder_layout(cfg: UNetConfig):
    """[(kind, ...)] per input block + channel list, shared by the schema and the graph walk.
    kinds: ("conv_in",), ("res", cin, cout, depth-or-0), ("down", c)."""
    mc = cfg.model_channels
    blocks, chans, ch, ds = [("conv_in",)], [mc], mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            depth = cfg.transformer_depth[level] if ds in cfg.attention_resolutions else 0
            blocks.append(("res", ch, mult * mc, depth))
            ch = mult * mc
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            blocks.append(("down", ch))
            chans.append(ch)
            ds *= 2
    return blocks, chans, ch, ds


def _decoder_layout(cfg: UNetConfig):
    """[(cin, cout, depth-or-0, upsample)] per output block."""
    mc = cfg.model_channels
    _, chans, ch, ds = _encoder_layout(cfg)
    chans = list(chans)
    out = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            depth = cfg.transformer_depth[level] if ds in cfg.attention_resolutions else 0
            up = bool(level and i == cfg.num_res_blocks)
            out.append((ch + ich, mc * mult, depth, up))
            ch = mc * mult
            if up:
                ds //= 2
    return out


def param_schema(cfg: UNetConfig, controlnet: bool = False) -> "OrderedDict[str, tuple]":
    """name -> shape of every parameter of sgm UNetModel (or cldm ControlNet) for `cfg`, upstream naming."""
    s: Dict[str, tuple] = OrderedDict()
    mc, ted = cfg.model_channels, 4 * cfg.model_channels
    s["time_embed.0.weight"], s["time_embed.0.bias"] = (ted, mc), (ted,)
    s["time_embed.2.weight"], s["time_embed.2.bias"] = (ted, ted), (ted,)
    s["label_emb.0.0.weight"], s["label_emb.0.0.bias"] = (ted, cfg.adm_in_channels), (ted,)
    s["label_emb.0.2.weight"], s["label_emb.0.2.bias"] = (ted, ted), (ted,)
    blocks, chans, ch, _ = _encoder_layout(cfg)
    for i, blk in enumerate(blocks):
        if blk[0] == "conv_in":
            s[f"input_blocks.{i}.0.weight"] = (mc, cfg.in_channels, 3, 3)
            s[f"input_blocks.{i}.0.bias"] = (mc,)
        elif blk[0] == "res":
            _res_schema(s, f"input_blocks.{i}.0", blk[1], blk[2], ted)
            if blk[3]:
                _st_schema(s, f"input_blocks.{i}.1", blk[2], blk[3], cfg.context_dim)
        else:
            s[f"input_blocks.{i}.0.op.weight"] = (blk[1], blk[1], 3, 3)
            s[f"input_blocks.{i}.0.op.bias"] = (blk[1],)
    _res_schema(s, "middle_block.0", ch, ch, ted)
    _st_schema(s, "middle_block.1", ch, cfg.transformer_depth[-1], cfg.context_dim)
    _res_schema(s, "middle_block.2", ch, ch, ted)
    if controlnet:
        for i, c in enumerate(chans):
            s[f"zero_convs.{i}.0.weight"], s[f"zero_convs.{i}.0.bias"] = (c, c, 1, 1), (c,)
        s["middle_block_out.0.weight"], s["middle_block_out.0.bias"] = (ch, ch, 1, 1), (ch,)
        cin = cfg.hint_channels
        for j, (cout, _) in enumerate(HINT_STACK):
            s[f"input_hint_block.{2 * j}.weight"], s[f"input_hint_block.{2 * j}.bias"] = (cout, cin, 3, 3), (cout,)
            cin = cout
        j = len(HINT_STACK)
        s[f"input_hint_block.{2 * j}.weight"], s[f"input_hint_block.{2 * j}.bias"] = (mc, cin, 3, 3), (mc,)
    else:
        for i, (cin, cout, depth, up) in enumerate(_decoder_layout(cfg)):
            _res_schema(s, f"output_blocks.{i}.0", cin, cout, ted)
            k = 1
            if depth:
                _st_schema(s, f"output_blocks.{i}.1", cout, depth, cfg.context_dim)
                k = 2
            if up:
                s[f"output_blocks.{i}.{k}.conv.weight"] = (cout, cout, 3, 3)
                s[f"output_blocks.{i}.{k}.conv.bias"] = (cout,)
        s["out.0.weight"], s["out.0.bias"] = (mc,), (mc,)
        s["out.2.weight"], s["out.2.bias"] = (cfg.out_channels, mc, 3, 3), (cfg.out_channels,)
    return s


class RandomWeights(Mapping):
    """Random-init parameters generated on demand on the device (no checkpoint can be downloaded here): torch's
    default init family — U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for Linear / Conv weights and biases, ones / zeros for the
    norms.  The layers sgm zero-initialises get the same uniform init so that every branch carries signal."""

    def __init__(self, schema: Mapping[str, tuple], seed: int, device):
        self.schema, self.seed, self.device = schema, seed, torch.device(device)
        self._index = {n: i for i, n in enumerate(schema)}

    def __getitem__(self, name: str) -> Tensor:
        shape = self.schema[name]
        g = torch.Generator(device=self.device).manual_seed(self.seed * 1_000_003 + self._index[name])
        is_norm = ".norm" in name or "in_layers.0." in name or "out_layers.0." in name or name.startswith("out.0.")
        if is_norm:
            return (torch.ones if name.endswith("weight") else torch.zeros)(shape, device=self.device)
        if name.endswith("bias"):
            wshape = self.schema[name[:-4] + "weight"]
            fan_in = math.prod(wshape[1:])
        else:
            fan_in = math.prod(shape[1:])
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, device=self.device) * 2 - 1) * bound

    def __iter__(self):
        return iter(self.schema)

    def __len__(self):
        return len(self.schema)


# ---------------------------------------------------------------------------------------------- packed parameters
class _Packed:
    """Weights of one network in kernel layouts (built once; fp16 GEMM operands, fp32 biases / norms / small convs)."""

    def __init__(self, w: Mapping[str, Tensor], device):
        self.w, self.dev = w, device
        self.t: Dict[str, Tensor] = {}

    def _get(self, name: str) -> Tensor:
        return self.w[name].detach().to(self.dev, torch.float32)

    def f32(self, name: str) -> Tensor:
        if name not in self.t:
            self.t[name] = self._get(name).contiguous()
        return self.t[name]

    def lin16(self, name: str) -> Tensor:
        key = name + "#16"
        if key not in self.t:
            x = self._get(name)
            self.t[key] = x.reshape(x.shape[0], -1).half().contiguous()
        return self.t[key]

    def conv16(self, name: str) -> Tensor:
        """[Co,Ci,3,3] -> [Co, 9*Ci], k = (ky*3+kx)*Ci + ci."""
        key = name + "#c16"
        if key not in self.t:
            x = self._get(name)
            self.t[key] = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1).half().contiguous()
        return self.t[key]

    def conv_split16(self, name: str) -> Tensor:
        """[Co,Ci,3,3] -> fp16 [Co, 2*Kp] = [W | W], Kp = round_up(9*Ci, 8): weight of the hi/lo-split im2col GEMM."""
        key = name + "#s16"
        if key not in self.t:
            x = self._get(name)
            w = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
            kp = (w.shape[1] + 7) // 8 * 8
            wp = torch.zeros((w.shape[0], kp), dtype=torch.float32, device=w.device)
            wp[:, :w.shape[1]] = w
            self.t[key] = torch.cat([wp, wp], 1).half().contiguous()
        return self.t[key]

    def conv32(self, name: str) -> Tensor:
        """[Co,Ci,3,3] -> fp32 [Co,3,3,Ci] for the direct conv."""
        key = name + "#c32"
        if key not in self.t:
            self.t[key] = self._get(name).permute(0, 2, 3, 1).contiguous()
        return self.t[key]

    def cat16(self, names: Sequence[str]) -> Tensor:
        key = "|".join(names) + "#16"
        if key not in self.t:
            self.t[key] = torch.cat([self._get(n) for n in names], 0).half().contiguous()
        return self.t[key]

    def geglu16(self, wname: str, bname: str) -> Tuple[Tensor, Tensor]:
        """GEGLU proj rows interleaved (value_j, gate_j) so one epilogue thread holds both halves."""
        key = wname + "#geglu"
        if key not in self.t:
            w, b = self._get(wname), self._get(bname)
            half = w.shape[0] // 2
            self.t[key] = torch.stack([w[:half], w[half:]], 1).reshape(w.shape).half().contiguous()
            self.t[key + "b"] = torch.stack([b[:half], b[half:]], 1).reshape(-1).contiguous()
        return self.t[key], self.t[key + "b"]


@dataclass
class Act:
    """NHWC activation: fp32 [B*H*W, C]."""
    t: Tensor
    B: int
    H: int
    W: int
    t16: Optional[Tensor] = None    # fp16 copy of `t` when the producer made one (skip concat)

    @property
    def C(self) -> int:
        return self.t.shape[1]

    def nchw(self) -> Tensor:
        return self.t.view(self.B, self.H, self.W, self.C).permute(0, 3, 1, 2).contiguous()


class _Net:
    """Operator-level graph walk shared by the UNet and the ControlNet."""

    def __init__(self, cfg: UNetConfig, weights: Mapping[str, Tensor], device, controlnet: bool):
        self.cfg, self.dev = cfg, torch.device(device)
        schema = param_schema(cfg, controlnet)
        missing = [n for n in schema if n not in weights]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} parameters, e.g. {missing[:3]}")
        for n, shp in schema.items():
            if tuple(weights[n].shape) != tuple(shp):
                raise ValueError(f"{n}: expected shape {tuple(shp)}, got {tuple(weights[n].shape)}")
        self.p = _Packed(weights, self.dev)
        self.heads_dim = cfg.num_head_channels
        if self.heads_dim != 64:
            raise ValueError("sgn_attention_f16 is specialised for head_dim 64 (SDXL)")
        self._prepack(schema)
        self._pack_context_kv(schema)
        self._pack_emb_layers(schema)
        self.p.w = None  # drop the reference to the source state_dict: only packed copies stay resident
        self._kv_ctx: Optional[Tensor] = None            # context the batched K/V projections were computed for
        self._kv_out: Dict[int, Tensor] = {}

    # every parameter is packed exactly once, up front, so forward() never allocates weights
    def _prepack(self, schema) -> None:
        p = self.p
        for n, shp in schema.items():
            if n.endswith("bias") or len(shp) == 1:
                if "ff.net.0.proj" not in n:
                    p.f32(n)
                continue
            if n.startswith(("time_embed", "label_emb")) or ".emb_layers." in n:
                p.f32(n)
            elif "input_hint_block" in n:
                if shp[1] % 64 == 0:
                    p.conv16(n)
                else:   # the fp32-exact routes stay available (hint_embedding picks per layer)
                    p.conv_split16(n)
                    p.conv32(n)
                    if (shp[1], shp[0]) in SMALL_TC_SHAPES:
                        p.conv16(n)
            elif len(shp) == 4 and shp[2] == 3:
                (p.conv16 if shp[1] % 64 == 0 else p.conv_split16)(n)
            elif len(shp) == 4:
                p.lin16(n)
            elif ".attn1.to_q" in n:
                b = n[: -len("to_q.weight")]
                p.cat16([b + "to_q.weight", b + "to_k.weight", b + "to_v.weight"])
            elif ".attn1.to_k" in n or ".attn1.to_v" in n or ".attn2.to_k" in n or ".attn2.to_v" in n:
                continue   # attn2 K/V: packed per channel width by _pack_context_kv
            elif "ff.net.0.proj.weight" in n:
                p.geglu16(n, n[:-6] + "bias")
            else:
                p.lin16(n)

    def _pack_context_kv(self, schema) -> None:
        """The cross-attention K/V projections of ALL transformer blocks read the same prompt context, so they are one
        GEMM per channel width: weights [sum over blocks of (to_k | to_v), ctx_dim], block b at row offset _kv_off[b]."""
        groups: Dict[int, List[str]] = {}
        for n, shp in schema.items():
            if n.endswith(".attn2.to_k.weight"):
                groups.setdefault(shp[0], []).append(n[: -len(".attn2.to_k.weight")])
        self._kv_w: Dict[int, Tensor] = {}
        self._kv_off: Dict[str, int] = {}
        for c, blocks in groups.items():
            rows = []
            for i, b in enumerate(blocks):
                self._kv_off[b] = 2 * c * i
                rows += [self.p._get(b + ".attn2.to_k.weight"), self.p._get(b + ".attn2.to_v.weight")]
            self._kv_w[c] = torch.cat(rows, 0).half().contiguous()

    def _pack_emb_layers(self, schema) -> None:
        """Every ResBlock projects the same emb (`emb_layers`: SiLU -> Linear): one launch for the whole network, weights
        concatenated along the rows, block `pre` at rows _emb_off[pre] .. + cout."""
        pres = [n[: -len(".emb_layers.1.weight")] for n in schema if n.endswith(".emb_layers.1.weight")]
        self._emb_off: Dict[str, Tuple[int, int]] = {}
        off = 0
        for pre in pres:
            cout = schema[pre + ".emb_layers.1.weight"][0]
            self._emb_off[pre] = (off, cout)
            off += cout
        self._emb_w = torch.cat([self.p.f32(pre + ".emb_layers.1.weight") for pre in pres], 0).contiguous()
        self._emb_b = torch.cat([self.p.f32(pre + ".emb_layers.1.bias") for pre in pres], 0).contiguous()
        self._emb_seg = torch.tensor([self._emb_off[pre][0] for pre in pres] + [off], dtype=torch.int32, device=self.dev)
        self._emb_for: Optional[Tensor] = None           # the emb tensor the batched projections were computed for
        self._emb_out: Optional[Tensor] = None

    def emb_projection(self, pre: str, emb: Tensor) -> Tensor:
        """emb_layers(emb) of ResBlock `pre` [Bt, cout]: a view of the batched projection when it was made for this emb
        (embed() does that), otherwise the block's own small linear (block-level callers, tests)."""
        if getattr(self, "_emb_for", None) is emb:
            off, cout = self._emb_off[pre]
            B = emb.shape[0]
            return self._emb_out[B * off:B * (off + cout)].view(B, cout)
        p = self.p
        return K.linear_small(emb, p.f32(pre + ".emb_layers.1.weight"), p.f32(pre + ".emb_layers.1.bias"), silu_in=True)

    def project_context(self, ctx16: Tensor) -> None:
        """One K/V GEMM per channel width for the whole network (forward() calls this once per step)."""
        self._kv_ctx = ctx16
        self._kv_out = {c: K.gemm_f16(ctx16, w, None, out_f16=True) for c, w in self._kv_w.items()}

    def context16(self, context: Tensor) -> Tensor:
        """fp16 prompt context with the batched K/V projections made for it.  The prompt does not change between the
        denoising steps of one request, so both are kept for as long as the caller passes the same (unmodified) tensor."""
        capturing = torch.cuda.is_current_stream_capturing()
        hit = (not capturing and getattr(self, "_ctx_ref", None) is context and self._ctx_version == context._version)
        if not hit:
            ctx16 = K.cast_f16(context.reshape(-1, context.shape[-1]))
            self.project_context(ctx16)
            # the cache holds the tensor itself (its storage cannot be recycled under us) and its in-place version
            self._ctx_ref, self._ctx_version = (None, -1) if capturing else (context, context._version)
        return self._kv_ctx

    def context_kv(self, block: str, c: int, ctx16: Tensor) -> Tuple[Tensor, Tensor]:
        """(k, v) [Bt*n_ctx, C] of transformer block `block`: slices of the batched projection when it was made for this
        context, otherwise this block's own GEMM (block-level callers, tests)."""
        off = self._kv_off[block]
        if self._kv_ctx is ctx16:
            kv = self._kv_out[c]
            return kv[:, off:off + c], kv[:, off + c:off + 2 * c]
        kv = K.gemm_f16(ctx16, self._kv_w[c][off:off + 2 * c], None, out_f16=True)
        return kv[:, :c], kv[:, c:]

    # ------------------------------------------------------------------ building blocks
    def gn16(self, x: Act, name: str, eps: float, act_silu: bool) -> Tensor:
        return K.group_norm_f16(x.t, x.B, x.H * x.W, 32, eps, self.p.f32(name + ".weight"), self.p.f32(name + ".bias"),
                                act_silu)

    def embed(self, t: Tensor, y: Tensor) -> Tensor:
        """emb = time_embed(timestep_embedding(t)) + label_emb(y)   [Bt, 4*mc]"""
        p = self.p
        te = K.timestep_embedding(t, self.cfg.model_channels)
        a = K.linear_small(te, p.f32("time_embed.0.weight"), p.f32("time_embed.0.bias"), silu_out=True)
        a = K.linear_small(a, p.f32("time_embed.2.weight"), p.f32("time_embed.2.bias"))
        b = K.linear_small(y, p.f32("label_emb.0.0.weight"), p.f32("label_emb.0.0.bias"), silu_out=True)
        emb = K.linear_small(b, p.f32("label_emb.0.2.weight"), p.f32("label_emb.0.2.bias"), residual=a)
        self._emb_out = K.linear_small_segments(emb, self._emb_w, self._emb_b, self._emb_seg, silu_in=True)
        self._emb_for = emb
        return emb

    def resblock(self, pre: str, x: Act, emb: Tensor) -> Act:
        p = self.p
        cout = p.f32(pre + ".in_layers.2.bias").shape[0]
        a16 = self.gn16(x, pre + ".in_layers.0", 1e-5, True)
        e = self.emb_projection(pre, emb)
        h = K.conv3x3_f16(a16.view(x.B, x.H, x.W, x.C), p.conv16(pre + ".in_layers.2.weight"),
                          p.f32(pre + ".in_layers.2.bias"), rowbias=e)
        b16 = self.gn16(Act(h, x.B, x.H, x.W), pre + ".out_layers.0", 1e-5, True)
        if x.C != cout:
            x16 = x.t16 if x.t16 is not None else K.cast_f16(x.t)
            skip = K.gemm_f16(x16, p.lin16(pre + ".skip_connection.weight"), p.f32(pre + ".skip_connection.bias"))
            out = K.conv3x3_f16(b16.view(x.B, x.H, x.W, cout), p.conv16(pre + ".out_layers.3.weight"),
                                p.f32(pre + ".out_layers.3.bias"), residual=skip, out=skip)
        else:
            out = K.conv3x3_f16(b16.view(x.B, x.H, x.W, cout), p.conv16(pre + ".out_layers.3.weight"),
                                p.f32(pre + ".out_layers.3.bias"), residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def transformer(self, pre: str, x: Act, depth: int, ctx16: Tensor, n_ctx: int) -> Act:
        p = self.p
        c, heads = x.C, x.C // 64
        n16 = self.gn16(x, pre + ".norm", 1e-6, False)
        t = K.gemm_f16(n16, p.lin16(pre + ".proj_in.weight"), p.f32(pre + ".proj_in.bias"))
        for d in range(depth):
            b = f"{pre}.transformer_blocks.{d}"
            a16 = K.layer_norm_f16(t, p.f32(b + ".norm1.weight"), p.f32(b + ".norm1.bias"))
            qkv = K.gemm_f16(a16, p.cat16([b + ".attn1.to_q.weight", b + ".attn1.to_k.weight", b + ".attn1.to_v.weight"]),
                             None, out_f16=True)
            o16 = K.attention_f16(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], x.B, heads)
            K.gemm_f16(o16, p.lin16(b + ".attn1.to_out.0.weight"), p.f32(b + ".attn1.to_out.0.bias"), residual=t, out=t)
            a16 = K.layer_norm_f16(t, p.f32(b + ".norm2.weight"), p.f32(b + ".norm2.bias"))
            q = K.gemm_f16(a16, p.lin16(b + ".attn2.to_q.weight"), None, out_f16=True)
            k2, v2 = self.context_kv(b, c, ctx16)
            o16 = K.attention_f16(q, k2, v2, x.B, heads)
            K.gemm_f16(o16, p.lin16(b + ".attn2.to_out.0.weight"), p.f32(b + ".attn2.to_out.0.bias"), residual=t, out=t)
            a16 = K.layer_norm_f16(t, p.f32(b + ".norm3.weight"), p.f32(b + ".norm3.bias"))
            wg, bg = p.geglu16(b + ".ff.net.0.proj.weight", b + ".ff.net.0.proj.bias")
            g16 = K.gemm_f16(a16, wg, bg, geglu=True)
            K.gemm_f16(g16, p.lin16(b + ".ff.net.2.weight"), p.f32(b + ".ff.net.2.bias"), residual=t, out=t)
        out = K.gemm_f16(K.cast_f16(t), p.lin16(pre + ".proj_out.weight"), p.f32(pre + ".proj_out.bias"), residual=x.t)
        return Act(out, x.B, x.H, x.W)

    def downsample(self, pre: str, x: Act) -> Act:
        col, ho, wo = K.im2col3x3_s2_f16(x.t, x.B, x.H, x.W)
        out = K.gemm_f16(col, self.p.conv16(pre + ".op.weight"), self.p.f32(pre + ".op.bias"))
        return Act(out, x.B, ho, wo)

    def upsample(self, pre: str, x: Act) -> Act:
        u16 = K.upsample2x_f16(x.t, x.B, x.H, x.W)
        out = K.conv3x3_f16(u16, self.p.conv16(pre + ".conv.weight"), self.p.f32(pre + ".conv.bias"))
        return Act(out, x.B, 2 * x.H, 2 * x.W)

    def input_block(self, i: int, h: Optional[Act], x_nchw: Optional[Tensor], emb: Tensor, ctx16: Tensor, n_ctx: int,
                    first_residual: Optional[Tensor] = None) -> Act:
        """input_blocks[i]: conv_in (from the NCHW latent), ResBlock (+ SpatialTransformer) or Downsample."""
        blk = _encoder_layout(self.cfg)[0][i]
        pre = f"input_blocks.{i}"
        if blk[0] == "conv_in":
            # 4 -> 320 channels: the latent's patches as fp16 [hi | lo] pairs against [W | W] on the tensor cores (the
            # fp32 conv to fp32 rounding, like the hint stack's stride-2 layers); the direct CUDA-core conv took 0.4-0.6 ms
            B, _, H, W = x_nchw.shape
            col, _, _, _ = K.im2col3x3_split_f16(x_nchw, True, 1)
            w16, bias = self.p.conv_split16(pre + ".0.weight"), self.p.f32(pre + ".0.bias")
            if first_residual is None:
                t = K.gemm_f16(col, w16, bias)
            else:                                      # ControlNet: + guided hint, one hint per CFG pair (image b uses hint b % Bh)
                res = first_residual.reshape(first_residual.shape[0], H * W, -1)
                t = torch.empty((B * H * W, w16.shape[0]), dtype=torch.float32, device=col.device)
                for b in range(B):
                    K.gemm_f16(col[b * H * W:(b + 1) * H * W], w16, bias, residual=res[b % res.shape[0]], out=t[b * H * W:(b + 1) * H * W])
            return Act(t, B, H, W)
        if blk[0] == "res":
            h = self.resblock(pre + ".0", h, emb)
            return self.transformer(pre + ".1", h, blk[3], ctx16, n_ctx) if blk[3] else h
        return self.downsample(pre + ".0", h)

    def middle(self, h: Act, emb: Tensor, ctx16: Tensor, n_ctx: int) -> Act:
        h = self.resblock("middle_block.0", h, emb)
        h = self.transformer("middle_block.1", h, self.cfg.transformer_depth[-1], ctx16, n_ctx)
        return self.resblock("middle_block.2", h, emb)

    def encoder(self, x_nchw: Tensor, emb: Tensor, ctx16: Tensor, n_ctx: int, first_residual: Optional[Tensor] = None,
                on_block=None) -> Tuple[List[Act], Act]:
        """input_blocks + middle_block.  `first_residual` (ControlNet): guided hint added to conv_in's output."""
        hs: List[Act] = []
        h: Optional[Act] = None
        for i in range(len(_encoder_layout(self.cfg)[0])):
            h = self.input_block(i, h, x_nchw, emb, ctx16, n_ctx, first_residual)
            hs.append(h)
            if on_block is not None:
                on_block(i, h)
        return hs, self.middle(h, emb, ctx16, n_ctx)


class SDXLUNetB200(_Net):
    """sgm UNetModel.forward(x, timesteps, context, y) with optional ControlNet residual injection."""

    def __init__(self, cfg: UNetConfig, weights: Mapping[str, Tensor], device="cuda"):
        super().__init__(cfg, weights, device, controlnet=False)

    def forward(self, x: Tensor, timesteps: Tensor, context: Tensor, y: Tensor, control: Optional[List[Tensor]] = None,
                control_weight: float = 1.0, taps: Optional[dict] = None) -> Tensor:
        """x [Bt,4,h,w] fp32 NCHW, timesteps [Bt], context [Bt,n_ctx,D], y [Bt,adm] -> eps [Bt,4,h,w] fp32 NCHW.
        control: ControlNetB200.forward output (NHWC fp32 residuals, encoder order + middle last)."""
        x, timesteps, context, y = (_f32(x, "x"), _f32(timesteps, "timesteps"), _f32(context, "context"), _f32(y, "y"))
        emb = self.embed(timesteps, y)
        n_ctx = context.shape[1]
        ctx16 = self.context16(context)
        tap = (lambda n, a: taps.__setitem__(n, a.nchw())) if taps is not None else (lambda n, a: None)
        hs, h = self.encoder(x, emb, ctx16, n_ctx, on_block=lambda i, a: tap(f"input_blocks.{i}", a))
        if callable(control):      # ControlNet running on a side stream: join it only now, where its outputs are first used
            control = control()
        control = list(control) if control is not None else None
        if control is not None:
            K.axpy_f32(control.pop(), control_weight, h.t)
        tap("middle_block", h)
        for i in range(len(_decoder_layout(self.cfg))):
            c = control.pop() if control is not None else None
            h = self.output_block(i, h, hs.pop(), c, control_weight, emb, ctx16, n_ctx)
            tap(f"output_blocks.{i}", h)
        return self.head(h)

    def output_block(self, i: int, h: Act, skip: Act, control: Optional[Tensor], control_weight: float, emb: Tensor,
                     ctx16: Tensor, n_ctx: int) -> Act:
        """output_blocks[i] on cat([h, skip + w * control]): ResBlock (+ SpatialTransformer) (+ Upsample)."""
        _, _, depth, up = _decoder_layout(self.cfg)[i]
        cat, cat16 = K.concat_f32(h.t, skip.t, control, control_weight, with_f16=True)   # every decoder ResBlock has a shortcut
        h = self.resblock(f"output_blocks.{i}.0", Act(cat, h.B, h.H, h.W, cat16), emb)
        k = 1
        if depth:
            h = self.transformer(f"output_blocks.{i}.1", h, depth, ctx16, n_ctx)
            k = 2
        if up:
            h = self.upsample(f"output_blocks.{i}.{k}", h)
        return h

    def head(self, h: Act) -> Tensor:
        """out: GroupNorm + SiLU + 3x3 conv -> eps, fp32 NCHW."""
        a16 = self.gn16(h, "out.0", 1e-5, True)
        return K.conv3x3_f16(a16.view(h.B, h.H, h.W, h.C), self.p.conv16("out.2.weight"), self.p.f32("out.2.bias"), nchw=True)


class ControlNetB200(_Net):
    """cldm ControlNet.forward(x, hint, timesteps, context, y) -> 9 encoder residuals + 1 middle residual (NHWC fp32)."""

    def __init__(self, cfg: UNetConfig, weights: Mapping[str, Tensor], device="cuda"):
        super().__init__(cfg, weights, device, controlnet=True)

    def hint_embedding(self, hint: Tensor) -> Tensor:
        """input_hint_block: hint [Bh,3,8h,8w] NCHW in [0,1] -> [Bh*h*w, mc] fp32."""
        p = self.p
        hint = _f32(hint, "hint")
        Bh, _, H, W = hint.shape
        # Two fp32-exact routes per layer, picked by measured cost on B200 (scratch/bench_hint.py): the direct CUDA-core
        # conv for the wide, shallow layers (3->16, 16->16 @2048^2, 32->32 @1024^2: their im2col would be 1-2.4 GB), and
        # the tensor cores for stride-2 / >= 64-channel layers: im2col with the fp32 activation split into fp16 hi + lo
        # against [W | W] reproduces the fp32 conv to fp32 rounding.
        x, nchw, cin = hint, True, self.cfg.hint_channels
        for j, (cout, stride) in enumerate(HINT_STACK):
            last = j == len(HINT_STACK) - 1
            wn, bn = f"input_hint_block.{2 * j}.weight", f"input_hint_block.{2 * j}.bias"
            if not nchw and (cin, cout) in SMALL_TC_SHAPES and (stride == 1 or (cin, cout) == (16, 32)):
                # 16 / 32 channels at the sheet resolution: hi / lo split in shared memory + mma.sync (fp32-exact, no im2col)
                x = K.conv3x3_small_tc(x, p.conv16(wn), p.f32(bn), stride=stride, act_silu=True, out_f16=last)
            elif stride == 2 or cin >= 64:
                col, ho, wo, _ = K.im2col3x3_split_f16(x, nchw, stride)
                y = K.gemm_f16(col, p.conv_split16(wn), p.f32(bn), act_silu=True, out_f16=last)
                x = y.view(Bh, ho, wo, cout)
            else:
                x = K.conv3x3_direct(x, nchw, p.conv32(wn), p.f32(bn), stride=stride, act_silu=True, out_f16=last)
            nchw, cin = False, cout
        j = len(HINT_STACK)
        return K.conv3x3_f16(x, p.conv16(f"input_hint_block.{2 * j}.weight"), p.f32(f"input_hint_block.{2 * j}.bias"))

    def forward(self, x: Tensor, hint: Tensor, timesteps: Tensor, context: Tensor, y: Tensor,
                guided: Optional[Tensor] = None) -> List[Tensor]:
        """`hint` may hold fewer images than `x` (the CFG pair shares one hint): image b uses hint b % Bh.
        `guided`: hint_embedding(hint) computed earlier (it depends on the hint only: once per request, not per step)."""
        x, timesteps, context, y = (_f32(x, "x"), _f32(timesteps, "timesteps"), _f32(context, "context"), _f32(y, "y"))
        emb = self.embed(timesteps, y)
        ctx16 = self.context16(context)
        if guided is None:
            guided = self.hint_embedding(hint)
        outs: List[Tensor] = []
        p = self.p

        def zero_conv(i: int, a: Act) -> None:
            outs.append(K.gemm_f16(K.cast_f16(a.t), p.lin16(f"zero_convs.{i}.0.weight"), p.f32(f"zero_convs.{i}.0.bias")))

        _, h = self.encoder(x, emb, ctx16, context.shape[1], first_residual=guided.view(hint.shape[0], x.shape[2], x.shape[3], -1),
                            on_block=zero_conv)
        outs.append(K.gemm_f16(K.cast_f16(h.t), p.lin16("middle_block_out.0.weight"), p.f32("middle_block_out.0.bias")))
        return outs


def _f32(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (signerf_b200 has no CPU path)")
    return t.to(torch.float32).contiguous()


# ---------------------------------------------------------------------------------------------- sampler (host scalars)
def sdxl_sigmas() -> Tensor:
    """k-diffusion DiscreteSchedule table for SDXL's scaled-linear betas (0.00085 .. 0.012, 1000 steps)."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    acp = torch.cumprod(1.0 - betas, dim=0)
    return ((1 - acp) / acp).sqrt().float()


def img2img_sigmas(steps: int = 20, denoising_strength: float = 0.9) -> List[float]:
    """A1111 img2img with a k-diffusion sampler: get_sigmas(steps)[steps - t_enc - 1:], t_enc = int(min(s, .999)*steps)
    (DiffuserConfig: num_inference_steps 20, denoising_strength 0.9 -> t_enc = 18, sigmas[1:] = 20 values = 19 UNet
    evaluations: k-diffusion get_sigmas(20) returns 21 values and the slice keeps t_enc + 2 of them; diffuser.py:33-39)."""
    table = sdxl_sigmas()
    t = torch.linspace(len(table) - 1, 0, steps)
    log_s = table.log()
    lo, hi, w = t.floor().long(), t.ceil().long(), t.frac()
    sig = torch.cat([((1 - w) * log_s[lo] + w * log_s[hi]).exp(), torch.zeros(1)])
    t_enc = int(min(denoising_strength, 0.999) * steps)
    return sig[steps - t_enc - 1:].tolist()


def sigma_to_t(sigma: float) -> float:
    """DiscreteSchedule.sigma_to_t (fractional timestep, log-sigma interpolation)."""
    log_s = sdxl_sigmas().log()
    ls = math.log(sigma)
    low = int((ls - log_s >= 0).cumsum(0).argmax().clamp(max=len(log_s) - 2))
    w = float(((log_s[low] - ls) / (log_s[low] - log_s[low + 1])).clamp(0, 1))
    return (1 - w) * low + w * (low + 1)


def ancestral_step(sigma_from: float, sigma_to: float, eta: float = 1.0) -> Tuple[float, float]:
    if not eta:
        return sigma_to, 0.0
    up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    return (sigma_to ** 2 - up ** 2) ** 0.5, up


class SDXLDenoiserB200:
    """UNet + ControlNet + sampler update = one "Euler a" step of the A1111 img2img request the reference sends
    (diffuser.py:132-169: cfg 7, ControlNet weight 0.8, Balanced, preprocessor none, guidance 0..1)."""

    def __init__(self, cfg: UNetConfig, unet_weights: Mapping[str, Tensor], ctrl_weights: Optional[Mapping[str, Tensor]],
                 device="cuda"):
        self.cfg, self.dev = cfg, torch.device(device)
        self.unet = SDXLUNetB200(cfg, unet_weights, device)
        self.ctrl = ControlNetB200(cfg, ctrl_weights, device) if ctrl_weights is not None else None
        # The ControlNet and the UNet encoder + middle block are independent until the first residual is added: they run
        # on two streams (fork / join, also inside a CUDA-graph capture) so that the CTAs of one network fill the partial
        # last waves of the other (640-CTA attention launches and 160-tile GEMMs on 148 SMs leave 14-28 % of a wave idle).
        self.two_streams = True
        self._side = torch.cuda.Stream(self.dev) if self.dev.type == "cuda" else None

    def eps(self, x: Tensor, sigma: float, context: Tensor, y: Tensor, hint: Optional[Tensor], control_weight: float = 0.8,
            guided: Optional[Tensor] = None) -> Tensor:
        """eps for the stacked (cond, uncond) batch: x [B,4,h,w] -> [2B,4,h,w]."""
        B = x.shape[0]
        c_in = 1.0 / math.sqrt(sigma * sigma + 1.0)
        xin = K.scale_cat2(x, c_in)
        t = torch.full((2 * B,), sigma_to_t(sigma), dtype=torch.float32, device=self.dev)
        control = None
        if self.ctrl is not None and hint is not None:
            if self.two_streams and self._side is not None:
                main, side = torch.cuda.current_stream(self.dev), self._side
                side.wait_stream(main)                       # fork: xin / t / guided are ready on the main stream
                with torch.cuda.stream(side):
                    ctl = self.ctrl.forward(xin, hint, t, context, y, guided=guided)

                def control():                               # join, called by the UNet after its own encoder + middle
                    main.wait_stream(side)
                    return ctl
            else:
                control = self.ctrl.forward(xin, hint, t, context, y, guided=guided)
        return self.unet.forward(xin, t, context, y, control=control, control_weight=control_weight)

    def step(self, x: Tensor, sigma: float, sigma_next: float, context: Tensor, y: Tensor, hint: Optional[Tensor],
             noise: Optional[Tensor] = None, init_latent: Optional[Tensor] = None, mask: Optional[Tensor] = None,
             cfg_scale: float = 7.0, control_weight: float = 0.8, guided: Optional[Tensor] = None):
        """-> (x_next, denoised, eps).  mask [B,1,h,w] = 1 where the original latent is kept."""
        eps = self.eps(x, sigma, context, y, hint, control_weight, guided)
        down, up = ancestral_step(sigma, sigma_next)
        x_next, den = K.cfg_euler_step(x, eps, init_latent, mask, noise if sigma_next > 0 else None, cfg_scale, sigma, down, up)
        return x_next, den, eps


class BenchUNet:
    """bench.py's diffusion half: one ControlNet + UNet + sampler step on the sheet latent, replayed as a CUDA graph.
    Random-init SDXL-base / ControlNet-XL weights (seed), latent N(0,1) seed 1, context / y N(0,1)."""

    def __init__(self, device, sheet_hw: Tuple[int, int], seed: int = 0, cfg: Optional[UNetConfig] = None, use_graph: bool = True):
        self.dev = torch.device(device)
        self.cfg = cfg or UNetConfig()
        self.net = SDXLDenoiserB200(self.cfg, RandomWeights(param_schema(self.cfg), seed, self.dev),
                                    RandomWeights(param_schema(self.cfg, True), seed + 1, self.dev), self.dev)
        hs, ws = sheet_hw
        self.h, self.w = hs // 8, ws // 8
        g = torch.Generator(device=self.dev).manual_seed(1)
        self.x = torch.randn(1, 4, self.h, self.w, generator=g, device=self.dev)
        self.init = torch.randn(1, 4, self.h, self.w, generator=g, device=self.dev)
        self.noise = torch.randn(1, 4, self.h, self.w, generator=g, device=self.dev)
        self.context = torch.randn(2, 77, self.cfg.context_dim, generator=g, device=self.dev)
        self.y = torch.randn(2, self.cfg.adm_in_channels, generator=g, device=self.dev)
        self.hint = torch.zeros(1, 3, hs, ws, device=self.dev)
        self.lat_mask = torch.zeros(1, 1, self.h, self.w, device=self.dev)
        sig = img2img_sigmas()
        self.sigma, self.sigma_next = sig[0], sig[1]
        self.use_graph, self.graph, self.out = use_graph, None, None
        self.launches_per_step = 0   # kernel launches one step issues (counted while capturing; replays issue the same)

    def _run(self):
        return self.net.step(self.x, self.sigma, self.sigma_next, self.context, self.y, self.hint, self.noise, self.init,
                             self.lat_mask)[0]

    def set_inputs(self, mask_sheet: Tensor, cond_sheet: Tensor) -> None:
        """Sheet-side conditioning -> static graph inputs: hint = condition replicated to 3 channels (ControlNet
        preprocessor "none"); latent mask = 1 - round(mask at latent resolution) (A1111: nmask = latmask)."""
        K.make_hint_and_latent_mask(cond_sheet, mask_sheet, self.hint, self.lat_mask)

    def step(self, image_sheet: Tensor, mask_sheet: Tensor, cond_sheet: Tensor) -> Tensor:
        self.set_inputs(mask_sheet, cond_sheet)
        if not self.use_graph:
            return self._run()
        if self.graph is None:
            self._run()  # warm-up outside capture (function attributes, lazy module load)
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
            self.launches_per_step = _lib.launch_count() - n0
        self.graph.replay()
        return self.out

    def full_loop_ms(self) -> Tuple[int, float]:
        """BASELINE config 3 at latent level: the whole img2img trajectory the reference requests (20 configured steps,
        denoising strength 0.9 -> 19 UNet+ControlNet evaluations, Euler-ancestral, CFG 7), eager launches.
        -> (UNet evaluations, device milliseconds)."""
        sig = img2img_sigmas()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x = self.init + self.noise * sig[0]
        a.record()
        guided = self.net.ctrl.hint_embedding(self.hint)       # once per request; the prompt K/V likewise (context16)
        for i in range(len(sig) - 1):
            x, _, _ = self.net.step(x, sig[i], sig[i + 1], self.context, self.y, self.hint, self.noise, self.init,
                                    self.lat_mask, guided=guided)
        b.record()
        torch.cuda.synchronize()
        if not bool(torch.isfinite(x).all()):
            raise RuntimeError("non-finite latent after the denoising loop")
        return len(sig) - 1, a.elapsed_time(b)

    def profile_eager(self) -> dict:
        """One eager step with every C-ABI call bracketed by CUDA events -> per-kernel-family time / FLOPs."""
        two, self.net.two_streams = self.net.two_streams, False    # one stream: per-call events must not span the other network
        try:
            K.start_profile()
            self._run()
            return K.stop_profile()
        finally:
            self.net.two_streams = two
