"""SURVEY §8(f) row 4 — the NeRF fine-tune step between two dataset generations (BASELINE config 5's "refinement rounds"):
host side of `sgn_train_*` (csrc/sgn_train.cu).

Mirrors, on the reference side,
  * `PatchPixelSampler.sample_method` (signerf/data/signerf_patch_pixel_sampler.py:44-80): random 32 x 32 patches, so
    that the batch reshapes into patches for the perceptual loss;
  * `SIGNeRFModel.get_loss_dict`'s image terms (signerf/signerf.py:41-60): `rgb_loss` = L1Loss or MSELoss on the rendered
    rgb, plus an LPIPS term on the patches.  LPIPS is a pretrained VGG the host owns (torchmetrics; no weights offline):
    it plugs in as any differentiable function of the rendered rgb through `render_rays_train` (a torch.autograd.Function
    whose backward is the CUDA backward of the field);
  * nerfstudio's Adam on the `fields` parameter group (signerf_config.py:43-50).
What this slice does NOT train yet: the proposal networks (interlevel loss), the distortion / normal regularisers and
per-image appearance embeddings (the mean embedding stays folded into the head's bias, as in the eval renderer); the rays'
bins come from the caller (the eval cascade's, or the flat piecewise bins)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .field import NerfactoFieldB200
from .ops import _ptr, _req, _stream


# ---------------------------------------------------------------------------------------------- patch sampler
@dataclass
class PatchPixelSamplerConfig:
    """signerf_patch_pixel_sampler.py:16-24 (+ the PixelSamplerConfig fields the method reads)."""
    patch_size: int = 32
    num_rays_per_batch: int = 4096
    ignore_mask: bool = False


class PatchPixelSampler:
    """signerf_patch_pixel_sampler.py:27-80.  `sample_method` -> int64 [batch, 3] = (image, y, x) indices: `batch //
    patch_size^2` patches with a uniformly random top-left corner, every pixel of each patch, patch-major / row-major."""

    def __init__(self, config: PatchPixelSamplerConfig):
        self.config = config
        self.set_num_rays_per_batch(config.num_rays_per_batch)

    def set_num_rays_per_batch(self, num_rays_per_batch: int) -> None:
        ps2 = self.config.patch_size ** 2
        self.num_rays_per_batch = (num_rays_per_batch // ps2) * ps2

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int,
                      mask: Optional[Tensor] = None, device="cpu", generator: Optional[torch.Generator] = None) -> Tensor:
        ps = self.config.patch_size
        if isinstance(mask, Tensor) and not self.config.ignore_mask:
            # with a mask the reference falls back to nerfstudio's per-pixel sampling of the mask's nonzero entries
            nonzero = torch.nonzero(mask[..., 0], as_tuple=False)
            pick = torch.randint(0, nonzero.shape[0], (batch_size,), generator=generator, device=nonzero.device)
            return nonzero[pick].to(device)
        sub_bs = batch_size // (ps ** 2)
        corner = torch.rand((sub_bs, 3), device=device, generator=generator) * torch.tensor(
            [num_images, image_height - ps, image_width - ps], device=device)
        idx = corner.view(sub_bs, 1, 1, 3).broadcast_to(sub_bs, ps, ps, 3).clone()
        yys, xxs = torch.meshgrid(torch.arange(ps, device=device), torch.arange(ps, device=device), indexing="ij")
        idx[..., 1] += yys
        idx[..., 2] += xxs
        return torch.floor(idx).long().flatten(0, 2)


# ---------------------------------------------------------------------------------------------- parameter views
class _DevicePtr:
    """CUDA array interface over memory the C library owns (the field's fp32 parameter block)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


MLP_LAYOUT = (("w_base0", (64, 32)), ("w_base1", (16, 64)), ("w_head0", (64, 32)), ("w_head1", (64, 64)), ("w_head2", (3, 64)),
              ("b_base0", (64,)), ("b_base1", (16,)), ("b_head0", (64,)), ("b_head1", (64,)), ("b_head2", (4,)), ("tail", (4,)))


def mlp_block_views(block: Tensor) -> Dict[str, Tensor]:
    """Named views into a parameter / gradient block of sgn_mlp_param_count() floats (include/signerf_b200.h layout)."""
    out, off = {}, 0
    for name, shape in MLP_LAYOUT:
        n = 1
        for d in shape:
            n *= d
        out[name] = block[off:off + n].view(*shape)
        off += n
    assert off == block.numel()
    return out


def nerfstudio_gradients(grad_table: Tensor, grad_block: Tensor, appearance_mean: Tensor) -> Dict[str, Tensor]:
    """The accumulated gradients under nerfstudio's parameter names / shapes (torch-fallback nerfacto field): what
    `loss.backward()` leaves in `.grad` of the reference model.  The head's first layer is [64, 63] = SH 16 | geo 15 |
    appearance 32 there; the appearance columns see the constant mean embedding, so their gradient is db (x) mean."""
    g = mlp_block_views(grad_block)
    wh0 = torch.cat([g["w_head0"][:, :16], g["w_head0"][:, 17:32],
                     g["b_head0"][:, None] * appearance_mean.to(grad_block.device)[None, :]], dim=1)
    return {
        "field.mlp_base.encoding.hash_table": grad_table,
        "field.mlp_base.mlp.layers.0.weight": g["w_base0"], "field.mlp_base.mlp.layers.0.bias": g["b_base0"],
        "field.mlp_base.mlp.layers.1.weight": g["w_base1"], "field.mlp_base.mlp.layers.1.bias": g["b_base1"],
        "field.mlp_head.layers.0.weight": wh0, "field.mlp_head.layers.0.bias": g["b_head0"],
        "field.mlp_head.layers.1.weight": g["w_head1"], "field.mlp_head.layers.1.bias": g["b_head1"],
        "field.mlp_head.layers.2.weight": g["w_head2"], "field.mlp_head.layers.2.bias": g["b_head2"][:3],
    }


# ---------------------------------------------------------------------------------------------- forward / backward
def _bins_args(bins: Tensor, n_rays: int):
    bins = _req(bins, torch.float32, "bins")
    if bins.dim() == 1:
        return bins, None, bins.shape[0] - 1
    if bins.dim() == 2 and bins.shape[0] == n_rays:
        return None, bins, bins.shape[1] - 1
    raise ValueError(f"bins must be [S+1] or [N,S+1], got {tuple(bins.shape)}")


def train_forward(fld: NerfactoFieldB200, origins: Tensor, directions: Tensor, bins: Tensor):
    """-> (rgb [N,3], acc [N], saved (sigma [N,S], color [N,S,3])); rgb is not clamped (training-mode RGBRenderer)."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    n = o.shape[0]
    shared, per_ray, S = _bins_args(bins, n)
    dev = fld.device
    sigma = torch.empty((n, S), dtype=torch.float32, device=dev)
    color = torch.empty((n, S, 3), dtype=torch.float32, device=dev)
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    acc = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_train_forward(fld.handle, _ptr(o), _ptr(d), n, _ptr(shared), _ptr(per_ray), S, _ptr(sigma),
                                                 _ptr(color), _ptr(rgb), _ptr(acc), _stream(dev)))
    return rgb, acc, (sigma, color)


def train_backward(fld: NerfactoFieldB200, origins: Tensor, directions: Tensor, bins: Tensor, saved, grad_rgb: Tensor,
                   grad_table: Tensor, grad_mlp: Tensor) -> None:
    """Accumulates dL/d(hash table) into grad_table [L*T,2] and dL/d(MLP block) into grad_mlp [sgn_mlp_param_count()]."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    n = o.shape[0]
    shared, per_ray, S = _bins_args(bins, n)
    sigma, color = saved
    g = _req(grad_rgb.reshape(-1, 3), torch.float32, "grad_rgb")
    for t, name in ((grad_table, "grad_table"), (grad_mlp, "grad_mlp")):
        _req(t, torch.float32, name)
    lib = _lib.load()
    dev = fld.device
    need = int(lib.sgn_train_ws_bytes(n, S))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sgn_train_backward(fld.handle, _ptr(o), _ptr(d), n, _ptr(shared), _ptr(per_ray), S, _ptr(sigma), _ptr(color),
                                          _ptr(g), _ptr(grad_table), _ptr(grad_mlp), _ptr(ws), need, _stream(dev)))


def rgb_loss(pred: Tensor, target: Tensor, use_l1: bool = True, want_grad: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """signerf.py:36-47: nerfstudio L1Loss / MSELoss (mean reduction) -> (loss [1], d loss / d pred)."""
    p = _req(pred, torch.float32, "pred")
    t = _req(target, torch.float32, "target")
    if p.shape != t.shape:
        raise ValueError("pred / target shapes differ")
    loss = torch.empty(1, dtype=torch.float32, device=p.device)
    grad = torch.empty_like(p) if want_grad else None
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().sgn_rgb_loss(_ptr(p), _ptr(t), p.numel(), int(use_l1), _ptr(loss), _ptr(grad), _stream(p.device)))
    return loss, grad


class FieldTrainer:
    """One optimizer over the main field's parameters: the caller's hash-table tensor (updated in place) and the
    field's fp32 MLP block.  `step` = forward + image loss + backward + Adam, all on the launching stream."""

    def __init__(self, fld: NerfactoFieldB200, lr: float = 1e-2, eps: float = 1e-15, betas: Tuple[float, float] = (0.9, 0.999),
                 use_l1: bool = True):
        self.field, self.lr, self.eps, self.betas, self.use_l1 = fld, lr, eps, betas, use_l1
        lib = _lib.load()
        dev = fld.device
        self.n_mlp = int(lib.sgn_mlp_param_count())
        ptr = C.c_void_p()
        _lib.check(lib.sgn_field_mlp_params(fld.handle, C.byref(ptr)))
        with torch.cuda.device(dev):
            self.mlp = torch.as_tensor(_DevicePtr(int(ptr.value), self.n_mlp), device=dev)      # aliases the field's block
        self.table = fld.grid.table                                                                 # [L*T, 2], by reference
        self.trainable = self.n_mlp - 4                                                             # not avg_density / padding
        self.grad_table = torch.zeros_like(self.table)
        self.grad_mlp = torch.zeros(self.n_mlp, dtype=torch.float32, device=dev)
        # The block's b_head0 is FOLDED: b + W_app . mean(appearance embedding).  nerfstudio optimises b and the appearance
        # columns W_app [64,32] of the head's first layer separately (the mean embedding is a constant of the step), so
        # both are kept here with their own Adam state and the folded bias is rewritten after every update.
        self.app_mean = fld.appearance_mean.detach().to(dev, torch.float32).contiguous()
        self.w_app = fld.head[0].weight.detach()[:, 31:63].to(dev, torch.float32).contiguous()
        self.b_head0 = fld.head[0].bias.detach().to(dev, torch.float32).contiguous()
        self.grad_w_app, self.grad_b_head0 = torch.zeros_like(self.w_app), torch.zeros_like(self.b_head0)
        self.state = {k: (torch.zeros_like(t), torch.zeros_like(t)) for k, t in
                      (("table", self.table), ("mlp", self.grad_mlp), ("w_app", self.w_app), ("b_head0", self.b_head0))}
        self.steps = 0

    def zero_grad(self) -> None:
        self.grad_table.zero_()
        self.grad_mlp.zero_()

    def backward(self, origins: Tensor, directions: Tensor, bins: Tensor, saved, grad_rgb: Tensor) -> None:
        train_backward(self.field, origins, directions, bins, saved, grad_rgb, self.grad_table, self.grad_mlp)

    def optimizer_step(self) -> None:
        self.steps += 1
        lib = _lib.load()
        dev = self.field.device
        g = mlp_block_views(self.grad_mlp)
        self.grad_b_head0.copy_(g["b_head0"])
        torch.outer(g["b_head0"], self.app_mean, out=self.grad_w_app)          # d/dW_app = d/db' (x) mean embedding
        g["b_head0"].zero_()                                                   # the folded slot is rewritten below, not stepped
        with torch.cuda.device(dev):
            for key, param, grad, n in (("table", self.table, self.grad_table, self.table.numel()),
                                        ("mlp", self.mlp, self.grad_mlp, self.trainable),
                                        ("w_app", self.w_app, self.grad_w_app, self.w_app.numel()),
                                        ("b_head0", self.b_head0, self.grad_b_head0, self.b_head0.numel())):
                m, v = self.state[key]
                _lib.check(lib.sgn_adam_step(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), n, self.lr, self.betas[0], self.betas[1],
                                             self.eps, self.steps, _stream(dev)))
        mlp_block_views(self.mlp)["b_head0"].copy_(self.b_head0 + self.w_app @ self.app_mean)

    def step(self, origins: Tensor, directions: Tensor, bins: Tensor, target_rgb: Tensor,
             extra_loss: Optional[Callable[[Tensor], Tensor]] = None) -> Tensor:
        """One fine-tune step on a batch of rays; `extra_loss(rgb)` is an optional differentiable torch term (the host's
        LPIPS on the 32 x 32 patches, signerf.py:49-60) whose gradient is added to the image loss's."""
        self.zero_grad()
        rgb, _, saved = train_forward(self.field, origins, directions, bins)
        loss, grad = rgb_loss(rgb, target_rgb.reshape(-1, 3).to(rgb.device), self.use_l1)
        if extra_loss is not None:
            leaf = rgb.detach().requires_grad_(True)
            extra = extra_loss(leaf)
            (g_extra,) = torch.autograd.grad(extra, leaf)
            grad = grad + g_extra
            loss = loss + extra.detach().reshape(1)
        self.backward(origins, directions, bins, saved, grad)
        self.optimizer_step()
        return loss

    def refresh_renderer(self) -> None:
        """Re-derive the tensor-core fragments from the updated parameters before the next sgn_render_* call."""
        with torch.cuda.device(self.field.device):
            _lib.check(_lib.load().sgn_field_refresh(self.field.handle, _stream(self.field.device)))


class _RenderRaysTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer: FieldTrainer, origins: Tensor, directions: Tensor, bins: Tensor, anchor: Tensor):
        rgb, _, saved = train_forward(trainer.field, origins, directions, bins)
        ctx.trainer, ctx.rays, ctx.saved = trainer, (origins, directions, bins), saved
        return rgb

    @staticmethod
    def backward(ctx, grad_rgb: Tensor):
        o, d, b = ctx.rays
        ctx.trainer.backward(o, d, b, ctx.saved, grad_rgb.contiguous())
        return None, None, None, None, None


def render_rays_train(trainer: FieldTrainer, origins: Tensor, directions: Tensor, bins: Tensor) -> Tensor:
    """Differentiable rgb [N,3] of a ray batch: any torch loss on it back-propagates into `trainer.grad_table /
    grad_mlp` through the CUDA backward (call `trainer.zero_grad()` before, `trainer.optimizer_step()` after)."""
    anchor = torch.zeros(1, device=trainer.field.device, requires_grad=True)     # makes the output require grad
    return _RenderRaysTrain.apply(trainer, origins, directions, bins, anchor)
