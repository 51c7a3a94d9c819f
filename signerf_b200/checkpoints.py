"""Weight files either side of the diffusion path (the checkpoints the reference's A1111 server loads,
`diffuser.py:47-49`: `sd_xl_base_1.0.safetensors [31e35c80fc]` and `diffusers_xl_depth_full [2f51180b]`) -> the state dicts
the in-process backend takes:

  * `split_sdxl_checkpoint`: a single-file SDXL checkpoint as A1111 / sgm store it - UNet under `model.diffusion_model.`, VAE
    under `first_stage_model.`, CLIP ViT-L under `conditioner.embedders.0.transformer.` (HuggingFace names) and OpenCLIP
    ViT-bigG under `conditioner.embedders.1.model.` (open_clip names) - into the four dicts of `SDXLDenoiserB200` / `VAEB200` /
    `CLIPTextEncoderB200`;
  * `open_clip_text_to_hf`: open_clip's text tower (`transformer.resblocks.N.attn.in_proj_weight`, `ln_1`, `mlp.c_fc`, ...)
    under HuggingFace `CLIPTextModelWithProjection` names (fused q | k | v split, `text_projection` transposed);
  * `controlnet_from_diffusers`: `diffusers_xl_depth_full` is a diffusers `ControlNetModel` file; sd-webui-controlnet renames
    it to cldm's module tree when it loads it ([EXT] `convert_from_diffuser_state_dict`).  Same renaming here: `down_blocks.i.
    resnets.j` -> `input_blocks.{3i+j+1}.0`, `attentions.j` -> `.1`, `downsamplers.0.conv` -> `input_blocks.{3(i+1)}.0.op`,
    `mid_block` -> `middle_block`, `controlnet_cond_embedding` -> `input_hint_block`, `controlnet_down_blocks.n` ->
    `zero_convs.n.0`, `controlnet_mid_block` -> `middle_block_out.0`, and the ResnetBlock2D / embedding member names.

Nothing here touches a GPU or does arithmetic beyond a transpose and a split; there are no weights offline, so the tests
build the source-format key sets from the architectures (tests/test_checkpoints.py) and, for the text tower, compare against
`transformers`' own module."""
from __future__ import annotations

import re
from typing import Dict, Mapping

import torch
from torch import Tensor

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."
CLIP_L_PREFIX = "conditioner.embedders.0.transformer."
CLIP_G_PREFIX = "conditioner.embedders.1.model."
CONTROLNET_PREFIX = "control_model."


def load_safetensors(path, device: str = "cpu") -> Dict[str, Tensor]:
    from safetensors.torch import load_file
    return load_file(str(path), device=device)


def _strip(sd: Mapping[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


# ---------------------------------------------------------------------------------------------- OpenCLIP text tower -> HF
def open_clip_text_to_hf(sd: Mapping[str, Tensor]) -> Dict[str, Tensor]:
    """open_clip text-tower names -> `CLIPTextModelWithProjection` names.  `in_proj_weight` [3D, D] is (q | k | v) stacked
    along the rows; open_clip applies `x @ text_projection` ([D, P]) where HF has `nn.Linear(D, P)`, i.e. the transpose."""
    out: Dict[str, Tensor] = {}
    for k, v in sd.items():
        if k == "token_embedding.weight":
            out["text_model.embeddings.token_embedding.weight"] = v
        elif k == "positional_embedding":
            out["text_model.embeddings.position_embedding.weight"] = v
        elif k in ("ln_final.weight", "ln_final.bias"):
            out["text_model.final_layer_norm." + k.split(".")[1]] = v
        elif k == "text_projection":
            out["text_projection.weight"] = v.t().contiguous()
        elif k.startswith("text_projection."):                      # already an nn.Linear (some exports)
            out[k] = v
        elif k in ("logit_scale", "attn_mask"):
            continue
        else:
            m = re.fullmatch(r"transformer\.resblocks\.(\d+)\.(.+)", k)
            if m is None:
                raise KeyError(f"unexpected open_clip text-tower key {k!r}")
            pre, rest = f"text_model.encoder.layers.{m.group(1)}.", m.group(2)
            if rest in ("attn.in_proj_weight", "attn.in_proj_bias"):
                kind = rest.rsplit("_", 1)[1]
                if v.shape[0] % 3:
                    raise ValueError(f"{k}: fused q | k | v tensor with {v.shape[0]} rows")
                for name, part in zip("qkv", v.chunk(3, dim=0)):
                    out[f"{pre}self_attn.{name}_proj.{kind}"] = part.contiguous()
                continue
            table = {"ln_1.": "layer_norm1.", "ln_2.": "layer_norm2.", "attn.out_proj.": "self_attn.out_proj.",
                     "mlp.c_fc.": "mlp.fc1.", "mlp.c_proj.": "mlp.fc2."}
            for a, b in table.items():
                if rest.startswith(a):
                    out[pre + b + rest[len(a):]] = v
                    break
            else:
                raise KeyError(f"unexpected open_clip text-tower key {k!r}")
    return out


# ---------------------------------------------------------------------------------------------- single-file SDXL checkpoint
def split_sdxl_checkpoint(sd: Mapping[str, Tensor]) -> Dict[str, Dict[str, Tensor]]:
    """{"unet", "vae", "clip_l", "clip_g"}: the parts of an A1111 / sgm SDXL checkpoint under the names the B200 modules read
    (`SDXLDenoiserB200(cfg, unet, controlnet)`, `VAEB200(cfg, vae)`, `CLIPTextEncoderB200(cfg, clip_l | clip_g)`)."""
    parts = {"unet": _strip(sd, UNET_PREFIX), "vae": _strip(sd, VAE_PREFIX), "clip_l": _strip(sd, CLIP_L_PREFIX)}
    g = _strip(sd, CLIP_G_PREFIX)
    parts["clip_g"] = open_clip_text_to_hf(g) if g else {}
    parts["clip_l"].pop("text_model.embeddings.position_ids", None)      # a buffer of older transformers versions
    for name in ("unet", "vae"):
        if not parts[name]:
            raise KeyError(f"no {name} tensors found (expected the prefix {UNET_PREFIX if name == 'unet' else VAE_PREFIX!r})")
    return parts


# ---------------------------------------------------------------------------------------------- diffusers ControlNet -> cldm
_RESNET = {"norm1.": "in_layers.0.", "conv1.": "in_layers.2.", "time_emb_proj.": "emb_layers.1.", "norm2.": "out_layers.0.",
           "conv2.": "out_layers.3.", "conv_shortcut.": "skip_connection."}
_TOP = {"time_embedding.linear_1.": "time_embed.0.", "time_embedding.linear_2.": "time_embed.2.",
        "add_embedding.linear_1.": "label_emb.0.0.", "add_embedding.linear_2.": "label_emb.0.2.",
        "conv_in.": "input_blocks.0.0.", "controlnet_mid_block.": "middle_block_out.0.",
        "controlnet_cond_embedding.conv_in.": "input_hint_block.0."}


def is_diffusers_controlnet(sd: Mapping[str, Tensor]) -> bool:
    return any(k.startswith(("controlnet_cond_embedding.", "controlnet_down_blocks.")) for k in sd)


def _resnet(rest: str) -> str:
    for a, b in _RESNET.items():
        if rest.startswith(a):
            return b + rest[len(a):]
    raise KeyError(f"unexpected ResnetBlock2D member {rest!r}")


def controlnet_from_diffusers(sd: Mapping[str, Tensor], layers_per_block: int = 2, hint_blocks: int = 6) -> Dict[str, Tensor]:
    """diffusers `ControlNetModel` names -> cldm `ControlNet` names (the module tree of oracle/sdxl_ref.py and
    `param_schema(cfg, controlnet=True)`).  Tensors are passed through untouched (SDXL's transformer projections are
    linear layers in both)."""
    stride = layers_per_block + 1                                  # input blocks per resolution level: resnets + downsampler
    out: Dict[str, Tensor] = {}
    for k, v in sd.items():
        new = None
        for a, b in _TOP.items():
            if k.startswith(a):
                new = b + k[len(a):]
                break
        if new is None:
            m = re.fullmatch(r"controlnet_cond_embedding\.blocks\.(\d+)\.(.+)", k)
            if m:
                new = f"input_hint_block.{2 * (int(m.group(1)) + 1)}.{m.group(2)}"
        if new is None and k.startswith("controlnet_cond_embedding.conv_out."):
            new = f"input_hint_block.{2 * (hint_blocks + 1)}." + k[len("controlnet_cond_embedding.conv_out."):]
        if new is None:
            m = re.fullmatch(r"controlnet_down_blocks\.(\d+)\.(.+)", k)
            if m:
                new = f"zero_convs.{m.group(1)}.0.{m.group(2)}"
        if new is None:
            m = re.fullmatch(r"down_blocks\.(\d+)\.(resnets|attentions)\.(\d+)\.(.+)", k)
            if m:
                idx = stride * int(m.group(1)) + int(m.group(3)) + 1
                new = (f"input_blocks.{idx}.0." + _resnet(m.group(4))) if m.group(2) == "resnets" else f"input_blocks.{idx}.1.{m.group(4)}"
        if new is None:
            m = re.fullmatch(r"down_blocks\.(\d+)\.downsamplers\.0\.conv\.(.+)", k)
            if m:
                new = f"input_blocks.{stride * (int(m.group(1)) + 1)}.0.op.{m.group(2)}"
        if new is None:
            m = re.fullmatch(r"mid_block\.resnets\.(\d+)\.(.+)", k)
            if m:
                new = f"middle_block.{2 * int(m.group(1))}." + _resnet(m.group(2))
        if new is None:
            m = re.fullmatch(r"mid_block\.attentions\.(\d+)\.(.+)", k)
            if m:
                new = f"middle_block.{2 * int(m.group(1)) + 1}.{m.group(2)}"
        if new is None:
            raise KeyError(f"unexpected diffusers ControlNetModel key {k!r}")
        if new in out:
            raise KeyError(f"{k!r} and another key both map to {new!r}")
        out[new] = v
    return out


def load_controlnet(path, device: str = "cpu") -> Dict[str, Tensor]:
    """A ControlNet file in either format -> cldm names (`control_model.` prefix of A1111-style files stripped)."""
    sd = load_safetensors(path, device)
    if any(k.startswith(CONTROLNET_PREFIX) for k in sd):
        sd = _strip(sd, CONTROLNET_PREFIX)
    return controlnet_from_diffusers(sd) if is_diffusers_controlnet(sd) else dict(sd)


def load_sdxl(path, device: str = "cpu") -> Dict[str, Dict[str, Tensor]]:
    return split_sdxl_checkpoint(load_safetensors(path, device))
