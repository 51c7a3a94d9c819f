"""SDXL prompt conditioning ("also" clause): the two CLIP text transformers the reference's A1111 server
runs once per request for `prompt` / `negative_prompt` (diffuser.py:132-141), on the sm_100a kernels.

  * CLIP ViT-L/14 text model  (sgm `FrozenCLIPEmbedder`, layer="hidden", layer_idx=11): 12 pre-LN layers, width 768, 12 heads,
    MLP 3072 quick_gelu, causal; SDXL takes the PENULTIMATE layer's hidden states (no final LayerNorm) -> [B,77,768]
  * OpenCLIP ViT-bigG/14 text model (sgm `FrozenOpenCLIPEmbedder2`, layer="penultimate"): 32 layers, width 1280, 20 heads,
    MLP 5120 gelu, causal; penultimate hidden states [B,77,1280] and pooled = text_projection(ln_final(last)[eot]) [B,1280]
  context = concat(L, bigG) -> [B,77,2048]; y = conditioning.sdxl_vector(pooled, H, W) -> [B,2816].
Parameter names are HuggingFace `CLIPTextModel` / `CLIPTextModelWithProjection`'s (`text_model.encoder.layers.N...`), the
checkpoints diffusers / A1111 load.  Tokenisation (BPE vocabulary, A1111's 75-token chunking and emphasis weights) stays
with the caller: the entry point takes `input_ids` [B,77].

Per layer: sgn_layer_norm_f16 -> fused q|k|v GEMM -> sgn_attention_causal_f16 (77 tokens, head_dim 64) -> out-projection
GEMM with the fp32 residual in place -> sgn_layer_norm_f16 -> fc1 GEMM -> sgn_act_f16 -> fc2 GEMM with residual."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Mapping, Optional

import torch
from torch import Tensor

from . import _lib
from . import nn_ops as K
from .ops import _ptr, _stream


@dataclass
class CLIPTextConfig:
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    vocab_size: int = 49408
    hidden_act: str = "quick_gelu"          # "quick_gelu" (CLIP-L) | "gelu" (OpenCLIP bigG)
    layer_norm_eps: float = 1e-5
    projection_dim: Optional[int] = None    # bigG: 1280 (text_projection, no bias)
    eos_token_id: int = 2                   # HF's legacy configs: the EOT position is argmax(input_ids)

    @staticmethod
    def clip_l() -> "CLIPTextConfig":
        return CLIPTextConfig()

    @staticmethod
    def open_clip_bigg() -> "CLIPTextConfig":
        return CLIPTextConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                              hidden_act="gelu", projection_dim=1280)


class CLIPTextEncoderB200:
    def __init__(self, cfg: CLIPTextConfig, state_dict: Mapping[str, Tensor], device="cuda"):
        if cfg.hidden_size != 64 * cfg.num_attention_heads:
            raise ValueError("the attention kernel is built for head_dim 64 (both SDXL text encoders have it)")
        if cfg.max_position_embeddings > 80:
            raise ValueError("causal attention kernel handles up to 80 tokens (CLIP uses 77)")
        self.cfg, self.dev = cfg, torch.device(device)
        sd = {k[len("text_model."):] if k.startswith("text_model.") else k: v for k, v in state_dict.items()}
        f32 = lambda k: sd[k].detach().to(self.dev, torch.float32).contiguous()  # noqa: E731
        f16 = lambda t: t.detach().to(self.dev, torch.float16).contiguous()      # noqa: E731
        self.tok, self.pos = f32("embeddings.token_embedding.weight"), f32("embeddings.position_embedding.weight")
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"encoder.layers.{i}."
            qkv_w = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0)
            qkv_b = torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], dim=0)
            self.layers.append({
                "ln1": (f32(p + "layer_norm1.weight"), f32(p + "layer_norm1.bias")),
                "ln2": (f32(p + "layer_norm2.weight"), f32(p + "layer_norm2.bias")),
                "qkv": (f16(qkv_w), qkv_b.detach().to(self.dev, torch.float32).contiguous()),
                "out": (f16(sd[p + "self_attn.out_proj.weight"]), f32(p + "self_attn.out_proj.bias")),
                "fc1": (f16(sd[p + "mlp.fc1.weight"]), f32(p + "mlp.fc1.bias")),
                "fc2": (f16(sd[p + "mlp.fc2.weight"]), f32(p + "mlp.fc2.bias")),
            })
        self.final_ln = (f32("final_layer_norm.weight"), f32("final_layer_norm.bias"))
        self.text_projection = (state_dict["text_projection.weight"].detach().to(self.dev, torch.float32).contiguous()
                                if cfg.projection_dim is not None else None)

    # ------------------------------------------------------------------ kernels
    def _embed(self, ids: Tensor) -> Tensor:
        B, T = ids.shape
        ids32 = ids.to(self.dev, torch.int32).contiguous()
        out = torch.empty((B * T, self.cfg.hidden_size), dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().sgn_embed_tokens(_ptr(ids32), _ptr(self.tok), _ptr(self.pos), B * T, T, self.cfg.hidden_size,
                                                    self.tok.shape[0], _ptr(out), _stream(self.dev)))
        return out

    def _attention(self, qkv: Tensor, B: int, T: int) -> Tensor:
        Cw, heads = self.cfg.hidden_size, self.cfg.num_attention_heads
        q, k, v = qkv[:, :Cw], qkv[:, Cw:2 * Cw], qkv[:, 2 * Cw:]
        out = torch.empty((B * T, Cw), dtype=torch.float16, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().sgn_attention_causal_f16(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), B,
                                                            heads, T, 0.125, _ptr(out), out.stride(0), _stream(self.dev)))
        return out

    def _act(self, x: Tensor) -> Tensor:
        out = torch.empty(x.shape, dtype=torch.float16, device=self.dev)
        mode = {"quick_gelu": 0, "gelu": 1}[self.cfg.hidden_act]
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().sgn_act_f16(_ptr(x), x.numel(), mode, _ptr(out), _stream(self.dev)))
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids: Tensor) -> Dict[str, Tensor]:
        """input_ids [B,77] -> {"penultimate" [B,77,W] (hidden_states[-2], what SDXL conditions on), "last" [B,77,W] after
        the final LayerNorm, "pooled" [B,W] (EOT token of "last") or [B,projection_dim] with the text projection}."""
        B, T = input_ids.shape
        eps = self.cfg.layer_norm_eps
        x = self._embed(input_ids)                                   # fp32 residual stream [B*T, W]
        penultimate = None
        for i, L in enumerate(self.layers):
            if i == len(self.layers) - 1:
                penultimate = x.clone()
            h = K.layer_norm_f16(x, *L["ln1"], eps)
            qkv = K.gemm_f16(h, *L["qkv"], out_f16=True)
            a = self._attention(qkv, B, T)
            K.gemm_f16(a, *L["out"], residual=x, out=x)
            h = K.layer_norm_f16(x, *L["ln2"], eps)
            u = self._act(K.gemm_f16(h, *L["fc1"]))
            K.gemm_f16(u, *L["fc2"], residual=x, out=x)
        last = K.layer_norm_f16(x, *self.final_ln, eps).to(torch.float32).view(B, T, -1)
        ids = input_ids.to(self.dev)
        eot = ids.argmax(dim=-1) if self.cfg.eos_token_id == 2 else (ids == self.cfg.eos_token_id).int().argmax(dim=-1)
        pooled = last[torch.arange(B, device=self.dev), eot].contiguous()
        if self.text_projection is not None:
            pooled = K.linear_small(pooled, self.text_projection, None)
        return {"penultimate": penultimate.view(B, T, -1), "last": last, "pooled": pooled}


def sdxl_prompt_conditioning(clip_l: CLIPTextEncoderB200, clip_g: CLIPTextEncoderB200, ids_l: Tensor, ids_g: Tensor,
                             height: int, width: int):
    """(context [B,77,2048], y [B,2816]) for `InProcessSDXL(context=..., y=...)`: rows = (prompt, negative prompt)."""
    from .conditioning import sdxl_vector
    a, b = clip_l.forward(ids_l), clip_g.forward(ids_g)
    context = torch.cat([a["penultimate"], b["penultimate"]], dim=-1).contiguous()
    return context, sdxl_vector(b["pooled"], height, width)
