"""A1111 img2img inpaint pre / post-processing around the denoising loop for
the request reference `Diffuser._diffuse_remote_sdwebui_controlnet` sends (signerf/diffuser/diffuser.py:121-169:
init image = the sheet, mask_blur 4, inpainting_fill 1 (original), inpaint_full_res 0, ControlNet preprocessor "none"):

    tensor_to_image truncation of sheet / mask / condition                         sgn_quantize_u8
    mask -> cv2.GaussianBlur (21,1) then (1,21), sigma 4                            sgn_gaussian_blur_u8 (bit-exact)
    mask_for_overlay = clip(2 * blurred)                                            sgn_inpaint_overlay_mask_u8
    latent mask = round(PIL bicubic resize of the blurred mask to W/8 x H/8 / 255)  sgn_pil_resize_bicubic_u8 (bit-exact)
    init_latent = scale_factor * VAE posterior sample of 2 * sheet - 1              vae.VAEB200.encode
    ... denoising loop (unet.SDXLDenoiserB200) ...
    decode -> uint8(255 * clamp((x+1)/2)) -> PIL paste + alpha_composite of the original outside the blurred mask
    -> image_to_tensor (v / 255)                                                    sgn_overlay_composite_u8 (bit-exact)

Torch tensors carry device pointers only; there is no CPU path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from . import nn_ops as K
from . import ops
from . import vae as V
from .nn_ops import _call, _chk
from .ops import _ptr


def gaussian_blur_u8(img: Tensor, ksize: int, sigma: float, horizontal: bool) -> Tensor:
    """cv2.GaussianBlur(img, (ksize,1) if horizontal else (1,ksize), sigma) for a uint8 [H,W] device image."""
    _chk(img, torch.uint8, "img")
    H, W = img.shape
    out = torch.empty_like(img)
    _call(img.device, _lib.load().sgn_gaussian_blur_u8, _ptr(img), H, W, int(ksize), float(sigma), int(horizontal), _ptr(out))
    return out


def pil_resize_bicubic_u8(img: Tensor, out_hw: Tuple[int, int]) -> Tensor:
    """PIL.Image.fromarray(img).resize((w, h), BICUBIC) for a uint8 [H,W] device image."""
    _chk(img, torch.uint8, "img")
    H, W = img.shape
    h, w = out_hw
    lib = _lib.load()
    ws = torch.empty(int(lib.sgn_pil_resize_ws_bytes(H, W, h, w)) // 4 + 1, dtype=torch.int32, device=img.device)
    out = torch.empty((h, w), dtype=torch.uint8, device=img.device)
    _call(img.device, lib.sgn_pil_resize_bicubic_u8, _ptr(img), H, W, h, w, _ptr(ws), _ptr(out))
    return out


def overlay_mask_u8(blurred: Tensor) -> Tensor:
    _chk(blurred, torch.uint8, "blurred")
    out = torch.empty_like(blurred)
    _call(blurred.device, _lib.load().sgn_inpaint_overlay_mask_u8, _ptr(blurred), blurred.numel(), _ptr(out))
    return out


def latent_keep_mask(lat_u8: Tensor) -> Tensor:
    """uint8 [h,w] -> fp32 [1,1,h,w]: 1 - round(v / 255) (1 = keep the original latent)."""
    _chk(lat_u8, torch.uint8, "lat_u8")
    out = torch.empty((1, 1) + tuple(lat_u8.shape), dtype=torch.float32, device=lat_u8.device)
    _call(lat_u8.device, _lib.load().sgn_latent_keep_mask, _ptr(lat_u8), lat_u8.numel(), _ptr(out))
    return out


def overlay_composite(generated: Tensor, original: Tensor, overlay_mask: Tensor) -> Tuple[Tensor, Tensor]:
    """A1111 apply_overlay on uint8 [H,W,3] images -> (uint8 [H,W,3], fp32 [H,W,3] = / 255)."""
    for t, n in ((generated, "generated"), (original, "original"), (overlay_mask, "overlay_mask")):
        _chk(t, torch.uint8, n)
    H, W, _ = generated.shape
    out_u8 = torch.empty_like(generated)
    out_f = torch.empty((H, W, 3), dtype=torch.float32, device=generated.device)
    _call(generated.device, _lib.load().sgn_overlay_composite_u8, _ptr(generated), _ptr(original), _ptr(overlay_mask), H, W,
          _ptr(out_u8), _ptr(out_f))
    return out_u8, out_f


def a1111_mask_blur(mask_u8: Tensor, mask_blur: int = 4) -> Tensor:
    """x pass then y pass, kernel_size = 2 * int(2.5 * blur + 0.5) + 1, each through uint8 (processing.py)."""
    if mask_blur <= 0:
        return mask_u8
    ks = 2 * int(2.5 * mask_blur + 0.5) + 1
    return gaussian_blur_u8(gaussian_blur_u8(mask_u8, ks, float(mask_blur), True), ks, float(mask_blur), False)


@dataclass
class InpaintState:
    """What `prepare` hands to the denoising loop and `finish` needs back."""
    init_latent: Tensor            # [1,4,H/8,W/8] scaled VAE latent of the quantised sheet
    hint: Tensor                   # [1,3,H,W] ControlNet hint in [0,1]
    keep_mask: Optional[Tensor]    # [1,1,H/8,W/8] 1 = keep the original latent (None: plain img2img)
    original_u8: Tensor            # [H,W,3]
    overlay_mask: Optional[Tensor]  # uint8 [H,W]


class A1111InpaintCodec:
    """Sheet-space <-> latent-space ends of the img2img request, around a VAEB200."""

    def __init__(self, vae: V.VAEB200, mask_blur: int = 4):
        self.vae, self.mask_blur = vae, mask_blur

    def prepare(self, original_image: Tensor, mask_image: Optional[Tensor], condition_image: Optional[Tensor],
                posterior_noise: Optional[Tensor] = None) -> InpaintState:
        """original [H,W,3], mask [H,W,1] (1 = repaint) or None, condition [H,W,1] or None: fp32 in [0,1] on the device;
        H, W multiples of 8 (the sheet is, datasetgenerator.py:498-508)."""
        dev = self.vae.dev
        H, W = int(original_image.shape[0]), int(original_image.shape[1])
        if H % 8 or W % 8:
            raise ValueError("image sides must be multiples of 8")
        orig_u8 = ops.quantize_u8(original_image.to(dev).contiguous())
        cond = (condition_image.to(dev, torch.float32) if condition_image is not None
                else torch.zeros((H, W, 1), device=dev)).contiguous()
        hint = torch.empty((1, 3, H, W), dtype=torch.float32, device=dev)
        scratch = torch.empty((1, 1, H // 8, W // 8), dtype=torch.float32, device=dev)
        K.make_hint_and_latent_mask(cond, cond, hint, scratch)      # hint = uint8(cond * 255) / 255 on 3 channels
        keep = overlay = None
        if mask_image is not None:
            mask_u8 = ops.quantize_u8(mask_image.to(dev).contiguous()).view(H, W)
            blurred = a1111_mask_blur(mask_u8, self.mask_blur)
            overlay = overlay_mask_u8(blurred)
            keep = latent_keep_mask(pil_resize_bicubic_u8(blurred, (H // 8, W // 8)))
        init = self.vae.encode(V.u8_to_vae_input(orig_u8), posterior_noise)
        return InpaintState(init, hint, keep, orig_u8, overlay)

    def finish(self, latent: Tensor, st: InpaintState) -> Tensor:
        """latent [1,4,h,w] -> edited image fp32 [H,W,3] in [0,1] (what reference Diffuser.diffuse returns)."""
        gen_u8 = V.vae_output_to_u8(self.vae.decode(latent))
        if st.overlay_mask is None:    # plain img2img: nothing to paste back (overlay alpha 0 everywhere)
            opaque = torch.full(gen_u8.shape[:2], 255, dtype=torch.uint8, device=gen_u8.device)
            return overlay_composite(gen_u8, gen_u8, opaque)[1]
        return overlay_composite(gen_u8, st.original_u8, st.overlay_mask)[1]
