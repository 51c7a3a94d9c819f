"""SDXL vector conditioning (`y`, adm_in_channels 2816) for the request reference `Diffuser` sends:
y = [pooled OpenCLIP-bigG text embedding (1280) | Fourier features of (original_h, original_w, crop_top, crop_left,
target_h, target_w), 256 each], the layout of sgm `GeneralConditioner` with `ConcatTimestepEmbedderND(outdim=256)`
embedders (sd_xl_base.yaml) as A1111 fills it (`sd_models_xl.get_learned_conditioning`: original = target = (height,
width) of the request, crop (0, 0)).  The two CLIP text encoders run once per prompt and stay with the caller
(`transformers` has both); this is the deterministic remainder.  The sinusoids come from `sgn_timestep_embedding`."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import Tensor

from . import nn_ops as K


def size_embedding(values: Sequence[float], device, dim: int = 256) -> Tensor:
    """ConcatTimestepEmbedderND: timestep_embedding(v, 256) = [cos | sin] for every scalar, concatenated -> [1, len * dim]."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float32, device=device)
    return K.timestep_embedding(t, dim).reshape(1, -1)


def sdxl_vector(pooled: Tensor, height: int, width: int, crop_top_left: Tuple[int, int] = (0, 0)) -> Tensor:
    """pooled [B, 1280] fp32 on the device -> y [B, 2816]; (height, width) of the image the request denoises (the sheet)."""
    if not pooled.is_cuda:
        raise RuntimeError("pooled must be a CUDA tensor (signerf_b200 has no CPU path)")
    emb = size_embedding([height, width, crop_top_left[0], crop_top_left[1], height, width], pooled.device)
    return torch.cat([pooled.to(torch.float32), emb.expand(pooled.shape[0], -1)], dim=1).contiguous()
