"""Mirror of reference signerf/diffuser/diffuser.py.  `DiffuserConfig` keeps every field (incl. the unused ones);
`Diffuser.diffuse` keeps the signature and the `_custom(original, condition, rendered, mask)` argument order
(diffuser.py:92-113).  mode="custom" — the hook the reference leaves unimplemented — runs the in-process
SDXL + ControlNet-depth denoiser (signerf_b200/unet.py) instead of POSTing PNGs to an A1111 server."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Literal, Optional, Type

import torch
from torch import Tensor

from ..unet import SDXLDenoiserB200, img2img_sigmas
from .base import InstantiateConfig


@dataclass
class DiffuserConfig(InstantiateConfig):
    """diffuser.py:19-60"""
    _target: Type = field(default_factory=lambda: Diffuser)
    mode: Literal["custom", "remoteSDWebUIControlNet"] = "remoteSDWebUIControlNet"
    url: str = "http://127.0.0.1"
    port: int = 5000
    prompt: str = "don't change the image"
    guidance_scale: float = 7
    image_guidance_scale: float = 1.5
    denoising_strength: float = 0.9
    num_inference_steps: int = 20
    lower_bound: float = 0.02
    upper_bound: float = 0.98
    seed: int = 1
    stable_diffusion_model: str = "sd_xl_base_1.0.safetensors [31e35c80fc]"
    controlnet_model: str = "diffusers_xl_depth_full [2f51180b]"
    controlnet_lowvram: bool = False
    controlnet_conditioning_scale: float = 0.8
    controlnet_conditioning_scale_start: float = 0.0
    controlnet_conditioning_scale_end: float = 1.0
    controlnet_control_mode: Literal["Balanced", "My prompt is more important", "ControlNet is more important"] = "Balanced"


class InProcessSDXL:
    """The A1111 img2img request of diffuser.py:132-169 executed in-process on latents:
    t_enc = int(min(strength, 0.999) * steps) Euler-ancestral steps with CFG, ControlNet (weight, Balanced,
    guidance 0..1) and the inpaint blend.  Prompt conditioning (context [2,77,2048], y [2,2816]: cond then uncond) comes
    from the caller (the CLIP text encoders run once per prompt and are not on the path); `codec` is the
    signerf_b200.inpaint.A1111InpaintCodec (VAE + mask blur + latent mask + overlay) around the latent loop."""

    def __init__(self, denoiser: SDXLDenoiserB200, context: Tensor, y: Tensor,
                 codec: Optional[object] = None, use_graph: bool = True):
        self.net, self.context, self.y, self.codec = denoiser, context, y, codec
        self.use_graph = use_graph          # replay the whole trajectory as ONE CUDA graph (captured per request shape)
        self._graphs: dict = {}

    def _trajectory(self, init_latent: Tensor, hint: Tensor, latent_mask: Tensor, noises: Tensor, sig, cfg_scale: float,
                    control_weight: float, context: Tensor, y: Tensor) -> Tensor:
        """The t_enc + 1 sampler steps on latents; `noises` [len(sig), 1,4,h,w]: row 0 seeds x, row i+1 is step i's."""
        x = init_latent + noises[0] * sig[0]
        guided = self.net.ctrl.hint_embedding(hint) if self.net.ctrl is not None else None   # constant over the steps
        for i in range(len(sig) - 1):
            noise = noises[i + 1] if sig[i + 1] > 0 else None
            x, _, _ = self.net.step(x, sig[i], sig[i + 1], context, y, hint, noise, init_latent, latent_mask,
                                    cfg_scale, control_weight, guided=guided)
        return x

    def denoise_latents(self, init_latent: Tensor, hint: Tensor, latent_mask: Tensor, steps: int, strength: float,
                        cfg_scale: float, control_weight: float, seed: int) -> Tensor:
        """init_latent [1,4,h,w]; hint [1,3,8h,8w] in [0,1]; latent_mask [1,1,h,w] = 1 where the original is kept."""
        sig = img2img_sigmas(steps, strength)
        dev = init_latent.device
        g = torch.Generator(device=dev).manual_seed(int(seed))
        # the same Philox stream as drawing step by step: one randn per sampler step, in order
        noises = torch.stack([torch.randn(init_latent.shape, generator=g, device=dev) for _ in range(len(sig))])
        if not self.use_graph:
            return self._trajectory(init_latent, hint, latent_mask, noises, sig, cfg_scale, control_weight, self.context, self.y)
        # Every sigma is a host scalar known up front, so the 19 UNet + ControlNet evaluations of a request (about 30 000
        # kernel launches) replay as one CUDA graph; the graph is keyed by everything that is baked into it.
        key = (tuple(init_latent.shape), tuple(hint.shape), int(steps), float(strength), float(cfg_scale), float(control_weight))
        slot = self._graphs.get(key)
        if slot is None:
            bufs = {"init": torch.empty_like(init_latent), "hint": torch.empty_like(hint), "mask": torch.empty_like(latent_mask),
                    "noise": torch.empty_like(noises), "ctx": torch.empty_like(self.context), "y": torch.empty_like(self.y)}
            for k, src in (("init", init_latent), ("hint", hint), ("mask", latent_mask), ("noise", noises), ("ctx", self.context),
                           ("y", self.y)):
                bufs[k].copy_(src)
            run = lambda: self._trajectory(bufs["init"], bufs["hint"], bufs["mask"], bufs["noise"], sig,  # noqa: E731
                                           cfg_scale, control_weight, bufs["ctx"], bufs["y"])
            self.net.step(bufs["init"], sig[0], sig[1], bufs["ctx"], bufs["y"], bufs["hint"], bufs["noise"][1], bufs["init"],
                          bufs["mask"], cfg_scale, control_weight)     # one eager step: lazy initialisation outside capture
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = run()
            slot = self._graphs[key] = (graph, bufs, out)
        graph, bufs, out = slot
        for k, src in (("init", init_latent), ("hint", hint), ("mask", latent_mask), ("noise", noises), ("ctx", self.context),
                       ("y", self.y)):
            bufs[k].copy_(src)
        graph.replay()
        return out.clone()


class Diffuser:
    """diffuser.py:62-195"""

    def __init__(self, config: DiffuserConfig, device: str) -> None:
        self.config, self.device = config, device
        self.prompt = config.prompt
        self.guidance_scale = config.guidance_scale
        self.image_guidance_scale = config.image_guidance_scale
        self.denoising_strength = config.denoising_strength
        self.num_inference_steps = config.num_inference_steps
        self.seed = config.seed
        self.stable_diffusion_model = config.stable_diffusion_model
        self.controlnet_conditioning_scale = config.controlnet_conditioning_scale
        self.controlnet_model = config.controlnet_model
        self.controlnet_lowvram = config.controlnet_lowvram
        self.controlnet_conditioning_scale_start = config.controlnet_conditioning_scale_start
        self.controlnet_conditioning_scale_end = config.controlnet_conditioning_scale_end
        self.controlnet_control_mode = config.controlnet_control_mode
        self.url = f"{self.config.url}:{self.config.port}"
        self.backend: Optional[InProcessSDXL] = None   # attached by the host application (weights are not ours to ship)

    def attach(self, backend: InProcessSDXL) -> None:
        self.backend = backend

    def diffuse(self, original_image: Tensor, rendered_image: Tensor, mask_image: Tensor = None,
                condition_image: Tensor = None) -> Tensor:
        if self.config.mode == "custom":
            return self._custom(original_image, condition_image, rendered_image, mask_image)
        if self.config.mode == "remoteSDWebUIControlNet":
            raise NotImplementedError("the HTTP client of diffuser.py:116-195 is out of scope (SURVEY §2 row 2): keep the "
                                      "reference's own Diffuser for mode='remoteSDWebUIControlNet', or use mode='custom'")
        raise ValueError(f"unknown diffuser mode {self.config.mode}")

    def _custom(self, original_image: Tensor, condition_image: Tensor, rendered_image: Tensor, mask_image: Tensor) -> Tensor:
        """Sheet RGB [H,W,3] + mask [H,W,1] + condition [H,W,1] -> edited sheet [H,W,3] in [0,1]."""
        if self.backend is None:
            raise RuntimeError("Diffuser(mode='custom') needs an InProcessSDXL backend: call diffuser.attach(...)")
        codec = self.backend.codec
        if codec is None:
            raise RuntimeError("no latent codec attached: build signerf_b200.inpaint.A1111InpaintCodec(VAEB200(...)) and pass "
                               "it to InProcessSDXL(codec=...)")
        dev = self.backend.net.dev
        st = codec.prepare(original_image.to(dev, torch.float32), None if mask_image is None else mask_image.to(dev, torch.float32),
                           None if condition_image is None else condition_image.to(dev, torch.float32))
        x = self.backend.denoise_latents(st.init_latent, st.hint, st.keep_mask, self.num_inference_steps,
                                         self.denoising_strength, self.guidance_scale, self.controlnet_conditioning_scale,
                                         self.seed)
        return codec.finish(x, st)
