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
`FieldTrainer` is the main-field slice (caller's bins, mean appearance embedding folded into the head's bias);
`NerfactoTrainer` is the whole nerfacto training step as `SIGNeRFModel` inherits it (signerf.py:62-68): the proposal
sampler in training mode (stratified bins, jittered PDF re-sampling; the draws are inputs), the two proposal networks
trained by `interlevel_loss`, `distortion_loss` on the final level, per-image appearance embeddings, and Adam on the
`fields` and `proposal_networks` groups, and - with the prediction branch's tensors - the normal regularisers of
`predict_normals=True` (signerf.py:69-80).  NOT built: the camera optimizer and the lr schedulers (host-side scalars); LPIPS's
pretrained network stays a host-side torch term (`extra_loss`)."""
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


def train_forward(fld: NerfactoFieldB200, origins: Tensor, directions: Tensor, bins: Tensor, head_bias: Optional[Tensor] = None):
    """-> (rgb [N,3], acc [N], saved (sigma [N,S], color [N,S,3])); rgb is not clamped (training-mode RGBRenderer).
    head_bias [N,64]: per-ray bias of the head's first layer (`appearance_bias`), None = folded mean embedding."""
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
        hb = None if head_bias is None else _req(head_bias.reshape(n, 64), torch.float32, "head_bias")
        _lib.check(_lib.load().sgn_train_forward(fld.handle, _ptr(o), _ptr(d), n, _ptr(shared), _ptr(per_ray), S, _ptr(hb),
                                                 _ptr(sigma), _ptr(color), _ptr(rgb), _ptr(acc), _stream(dev)))
    return rgb, acc, (sigma, color)


def train_backward(fld: NerfactoFieldB200, origins: Tensor, directions: Tensor, bins: Tensor, saved, grad_rgb: Tensor,
                   grad_table: Tensor, grad_mlp: Tensor, grad_weights: Optional[Tensor] = None,
                   head_bias: Optional[Tensor] = None, grad_head_bias: Optional[Tensor] = None,
                   grad_geo: Optional[Tensor] = None) -> None:
    """Accumulates dL/d(hash table) into grad_table [L*T,2] and dL/d(MLP block) into grad_mlp [sgn_mlp_param_count()];
    grad_weights [N,S]: gradient of the terms reading the weights directly (distortion loss); head_bias / grad_head_bias
    [N,64]: the per-ray appearance bias of the forward and where its gradient accumulates."""
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
        gw = None if grad_weights is None else _req(grad_weights.reshape(n, S), torch.float32, "grad_weights")
        hb = None if head_bias is None else _req(head_bias.reshape(n, 64), torch.float32, "head_bias")
        ghb = None if grad_head_bias is None else _req(grad_head_bias.reshape(n, 64), torch.float32, "grad_head_bias")
        gg = None if grad_geo is None else _req(grad_geo.reshape(n, S, 15), torch.float32, "grad_geo")
        _lib.check(lib.sgn_train_backward(fld.handle, _ptr(o), _ptr(d), n, _ptr(shared), _ptr(per_ray), S, _ptr(hb), _ptr(sigma),
                                          _ptr(color), _ptr(g), _ptr(gw), _ptr(gg), _ptr(grad_table), _ptr(grad_mlp), _ptr(ghb),
                                          _ptr(ws), need, _stream(dev)))


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

    def state_dict(self) -> Dict[str, Tensor]:
        """The trained parameters under nerfstudio's torch-fallback names (the `_model.` part of a pipeline checkpoint,
        signerf_trainer.py:278-306): what `load_into` copies back into the reference's model and what
        `plugin.FusedNerfactoGraph.from_state_dict` reads.  Detached copies."""
        v = mlp_block_views(self.mlp)
        wh0 = torch.cat([v["w_head0"][:, :16], v["w_head0"][:, 17:32], self.w_app], dim=1)      # SH 16 | geo 15 | appearance 32
        sd = {"field.mlp_base.encoding.hash_table": self.table,
              "field.mlp_base.mlp.layers.0.weight": v["w_base0"], "field.mlp_base.mlp.layers.0.bias": v["b_base0"],
              "field.mlp_base.mlp.layers.1.weight": v["w_base1"], "field.mlp_base.mlp.layers.1.bias": v["b_base1"],
              "field.mlp_head.layers.0.weight": wh0, "field.mlp_head.layers.0.bias": self.b_head0,
              "field.mlp_head.layers.1.weight": v["w_head1"], "field.mlp_head.layers.1.bias": v["b_head1"],
              "field.mlp_head.layers.2.weight": v["w_head2"], "field.mlp_head.layers.2.bias": v["b_head2"][:3]}
        return {k: t.detach().clone() for k, t in sd.items()}

    # the other parameter layout nerfstudio's torch-fallback modules have used (plugin/model.py from_state_dict)
    _ALT_NAMES = (("field.mlp_base.encoding.", "field.mlp_base_grid."), ("field.mlp_base.mlp.", "field.mlp_base_mlp."),
                  (".mlp_base.encoding.", ".encoding."), (".mlp_base.mlp.", ".mlp_base."),
                  ("field_head_pred_normals.net.", "field_head_pred_normals."))

    def load_into(self, model: torch.nn.Module) -> list:
        """Copy the trained parameters into the matching tensors of a nerfstudio model (`pipeline.model`), in place, so that
        the reference's own checkpointing / viewer / eval see the fine-tuned field.  Returns the names written."""
        target = dict(model.state_dict(keep_vars=True))
        written = []
        with torch.no_grad():
            for name, value in self.state_dict().items():
                cands = [name] + [name.replace(a, b) for a, b in self._ALT_NAMES if a in name]
                hit = next((c for c in cands if c in target), None)
                if hit is None:
                    if name.startswith("field.embedding_appearance"):
                        continue                       # SIGNeRF re-creates the table per dataset (signerf_pipeline.py:110-111)
                    raise KeyError(f"the model has no parameter for {name} (tried {cands})")
                if tuple(target[hit].shape) != tuple(value.shape):
                    raise ValueError(f"{hit}: model has shape {tuple(target[hit].shape)}, trained tensor {tuple(value.shape)}")
                target[hit].copy_(value.to(target[hit].device, target[hit].dtype))
                written.append(hit)
        return written

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


# ---------------------------------------------------------------------------------------------- proposal half of the step
@dataclass
class TrainSamples:
    """One batch's sampling cascade (levels 0..2 = initial / after proposal 0 / after proposal 1)."""
    spacing: Tuple[Tensor, Tensor, Tensor]      # [N, S_l + 1] bin edges, spacing domain
    euclid: Tuple[Tensor, Tensor, Tensor]       # [N, S_l + 1] bin edges, metres
    sigma: Tuple[Tensor, Tensor]                # [N, S_l] proposal densities on level l
    weights: Tuple[Tensor, Tensor]              # [N, S_l] proposal weights on level l


def train_sample(fld: NerfactoFieldB200, origins: Tensor, directions: Tensor, counts: Tuple[int, int, int] = (256, 96, 48),
                 near: float = 0.05, far: float = 1000.0, jitter: Optional[Tensor] = None, anneal: float = 1.0) -> TrainSamples:
    """ProposalNetworkSampler.generate_ray_samples while training; jitter [3,N] uniform draws, None = eval bins; anneal:
    exponent on the proposal weights before each re-sampling (`proposal_anneal`)."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    n, dev = o.shape[0], fld.device
    S = tuple(int(c) for c in counts)
    f32 = dict(dtype=torch.float32, device=dev)
    sp = tuple(torch.empty((n, c + 1), **f32) for c in S)
    eu = tuple(torch.empty((n, c + 1), **f32) for c in S)
    sg = tuple(torch.empty((n, c), **f32) for c in S[:2])
    w = tuple(torch.empty((n, c), **f32) for c in S[:2])
    out = _lib.SgnTrainSamples()
    for l in range(3):
        out.d_spacing[l], out.d_euclid[l] = sp[l].data_ptr(), eu[l].data_ptr()
    for l in range(2):
        out.d_sigma[l], out.d_weights[l] = sg[l].data_ptr(), w[l].data_ptr()
    j = None if jitter is None else _req(jitter.reshape(3, n), torch.float32, "jitter")
    lib = _lib.load()
    need = int(lib.sgn_train_sample_ws_bytes(n, S[0], S[1]))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sgn_train_sample(fld.handle, _ptr(o), _ptr(d), n, S[0], S[1], S[2], float(near), float(far), _ptr(j),
                                        float(anneal), C.byref(out), _ptr(ws), need, _stream(dev)))
    return TrainSamples(sp, eu, sg, w)


def proposal_anneal(step: int, max_num_iters: int = 1000, slope: float = 10.0) -> float:
    """[EXT] ProposalNetworkSampler.set_anneal as NerfactoModel's training callback drives it: bias(train_frac, slope) with
    train_frac = clip(step / proposal_weights_anneal_max_num_iters, 0, 1) - 0 at step 0, 1 from max_num_iters on."""
    x = min(max(step / max_num_iters, 0.0), 1.0)
    return slope * x / ((slope - 1.0) * x + 1.0)


class ProposalUpdateSchedule:
    """[EXT] ProposalNetworkSampler.step_cb / generate_ray_samples: the proposal networks only get gradients on "update"
    steps - `steps_since_update > update_sched(step) or step < 10`, update_sched = clamp(lerp(step / proposal_warmup, 1,
    proposal_update_every), 1, proposal_update_every) (nerfacto: warm-up 5000, every 5)."""

    def __init__(self, update_every: int = 5, warmup: int = 5000):
        self.update_every, self.warmup, self.since = update_every, warmup, 0

    def __call__(self, step: int) -> bool:
        self.since += 1                                              # step_cb, before the forward of this step
        sched = min(max(1.0 + (self.update_every - 1.0) * step / max(self.warmup, 1), 1.0), float(self.update_every))
        updated = self.since > sched or step < 10
        if updated:
            self.since = 0
        return updated


def weights_from_density(euclid: Tensor, sigma: Tensor) -> Tensor:
    """RaySamples.get_weights: [N,S+1] bin edges, [N,S] densities -> [N,S]."""
    e, s = _req(euclid, torch.float32, "euclid"), _req(sigma, torch.float32, "sigma")
    w = torch.empty_like(s)
    with torch.cuda.device(s.device):
        _lib.check(_lib.load().sgn_weights_from_density(_ptr(e), _ptr(s), s.shape[0], s.shape[1], _ptr(w), _stream(s.device)))
    return w


def interlevel_loss(spacing_final: Tensor, weights_final: Tensor, spacing_prop: Tensor, weights_prop: Tensor, loss: Tensor,
                    mult: float = 1.0) -> Tensor:
    """One proposal level's term of losses.py interlevel_loss: loss [1] += term, returns d term / d weights_prop [N,Sp]."""
    n, sf, spn = weights_final.shape[0], weights_final.shape[1], weights_prop.shape[1]
    g = torch.empty_like(weights_prop)
    ws = torch.empty(n * (spn + 1), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().sgn_interlevel_loss(_ptr(_req(spacing_final, torch.float32, "spacing_final")),
                                                   _ptr(_req(weights_final, torch.float32, "weights_final")), sf,
                                                   _ptr(_req(spacing_prop, torch.float32, "spacing_prop")),
                                                   _ptr(_req(weights_prop, torch.float32, "weights_prop")), spn, n, float(mult),
                                                   _ptr(loss), _ptr(g), _ptr(ws), ws.numel() * 4, _stream(g.device)))
    return g


def distortion_loss(spacing: Tensor, weights: Tensor, loss: Tensor, mult: float = 0.002, want_grad: bool = True) -> Optional[Tensor]:
    """losses.py distortion_loss on the final level: loss [1] += mult * mean, returns its gradient w.r.t. the weights."""
    w = _req(weights, torch.float32, "weights")
    g = torch.empty_like(w) if want_grad else None
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().sgn_distortion_loss(_ptr(_req(spacing, torch.float32, "spacing")), _ptr(w), w.shape[0], w.shape[1],
                                                   float(mult), _ptr(loss), _ptr(g), _stream(w.device)))
    return g


def prop_backward(fld: NerfactoFieldB200, level: int, origins: Tensor, directions: Tensor, euclid: Tensor, sigma: Tensor,
                  grad_weights: Tensor, grad_table: Tensor, grad_mlp: Tensor) -> None:
    """dL/d(weights of proposal level) -> that network's hash-table and MLP gradients (+=)."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    n, S = sigma.shape
    ws = torch.empty(n * S, dtype=torch.float32, device=fld.device)
    with torch.cuda.device(fld.device):
        _lib.check(_lib.load().sgn_prop_backward(fld.handle, level, _ptr(o), _ptr(d), n, S, _ptr(_req(euclid, torch.float32, "euclid")),
                                                 _ptr(_req(sigma, torch.float32, "sigma")),
                                                 _ptr(_req(grad_weights, torch.float32, "grad_weights")),
                                                 _ptr(_req(grad_table, torch.float32, "grad_table")),
                                                 _ptr(_req(grad_mlp, torch.float32, "grad_mlp")), _ptr(ws), ws.numel() * 4,
                                                 _stream(fld.device)))


def appearance_bias(w_app: Tensor, bias: Tensor, embedding: Tensor, camera_indices: Tensor) -> Tensor:
    """head_bias [N,64] = bias + W_app . embedding[camera]  (NerfactoField while training)."""
    cam = _req(camera_indices, torch.int32, "camera_indices")
    out = torch.empty((cam.shape[0], 64), dtype=torch.float32, device=cam.device)
    with torch.cuda.device(cam.device):
        _lib.check(_lib.load().sgn_appearance_bias(_ptr(_req(w_app, torch.float32, "w_app")), _ptr(_req(bias, torch.float32, "bias")),
                                                   _ptr(_req(embedding, torch.float32, "embedding")), embedding.shape[0], _ptr(cam),
                                                   cam.shape[0], _ptr(out), _stream(cam.device)))
    return out


def appearance_bias_backward(w_app: Tensor, embedding: Tensor, camera_indices: Tensor, grad_head_bias: Tensor, grad_w_app: Tensor,
                             grad_bias: Tensor, grad_embedding: Tensor) -> None:
    cam = _req(camera_indices, torch.int32, "camera_indices")
    with torch.cuda.device(cam.device):
        _lib.check(_lib.load().sgn_appearance_bias_backward(
            _ptr(_req(w_app, torch.float32, "w_app")), _ptr(_req(embedding, torch.float32, "embedding")), embedding.shape[0],
            _ptr(cam), _ptr(_req(grad_head_bias, torch.float32, "grad_head_bias")), cam.shape[0],
            _ptr(_req(grad_w_app, torch.float32, "grad_w_app")), _ptr(_req(grad_bias, torch.float32, "grad_bias")),
            _ptr(_req(grad_embedding, torch.float32, "grad_embedding")), _stream(cam.device)))


# ---------------------------------------------------------------------------------------------- normal regularisers
PN_LAYOUT = (("w1", (64, 64)), ("w2", (64, 64)), ("wh", (3, 64)), ("w0", (64, 27)), ("b0", (64,)), ("b1", (64,)), ("b2", (64,)),
             ("bh", (4,)))
PN_NAMES = {"w0": "field.mlp_pred_normals.layers.0.weight", "b0": "field.mlp_pred_normals.layers.0.bias",
            "w1": "field.mlp_pred_normals.layers.1.weight", "b1": "field.mlp_pred_normals.layers.1.bias",
            "w2": "field.mlp_pred_normals.layers.2.weight", "b2": "field.mlp_pred_normals.layers.2.bias",
            "wh": "field.field_head_pred_normals.net.weight", "bh": "field.field_head_pred_normals.net.bias"}


def pn_block_views(block: Tensor) -> Dict[str, Tensor]:
    """Named views into a parameter / gradient block of sgn_pred_normals_param_count() floats (include/signerf_b200.h)."""
    out, off = {}, 0
    for name, shape in PN_LAYOUT:
        n = 1
        for d in shape:
            n *= d
        out[name] = block[off:off + n].view(*shape)
        off += n
    assert off == block.numel()
    return out


def pack_pred_normals(params: Dict[str, Tensor], device) -> Tensor:
    """nerfstudio's `field.mlp_pred_normals.layers.*` / `field.field_head_pred_normals.net.*` tensors -> the C block."""
    block = torch.zeros(int(_lib.load().sgn_pred_normals_param_count()), dtype=torch.float32, device=device)
    v = pn_block_views(block)
    for key, name in PN_NAMES.items():
        t = params[name] if name in params else params[name.replace(".net.", ".")]
        (v[key][:3] if key == "bh" else v[key]).copy_(t.detach().to(device, torch.float32))
    return block


def normals_forward(fld: NerfactoFieldB200, pn_block: Tensor, origins: Tensor, directions: Tensor, ray_bins: Tensor):
    """-> (analytic normals [N,S,3] (no graph: Field.get_normals), predicted normals [N,S,3]) on the final level's bins."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    b = _req(ray_bins, torch.float32, "ray_bins")
    n, S = b.shape[0], b.shape[1] - 1
    normals = torch.empty((n, S, 3), dtype=torch.float32, device=fld.device)
    pred = torch.empty_like(normals)
    with torch.cuda.device(fld.device):
        _lib.check(_lib.load().sgn_train_normals_forward(fld.handle, _ptr(_req(pn_block, torch.float32, "pn_block")), _ptr(o), _ptr(d),
                                                         n, _ptr(b), S, _ptr(normals), _ptr(pred), _stream(fld.device)))
    return normals, pred


def normal_losses(weights: Tensor, normals: Tensor, pred: Tensor, directions: Tensor, loss_orientation: Tensor,
                  loss_pred_normal: Tensor, orientation_mult: float = 1e-4, pred_normal_mult: float = 1e-3) -> Tensor:
    """losses.py orientation_loss / pred_normal_loss (means over rays x multipliers): loss_* [1] +=; returns d / d pred."""
    w = _req(weights, torch.float32, "weights")
    g = torch.empty_like(pred)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().sgn_normal_losses(_ptr(w), _ptr(_req(normals, torch.float32, "normals")),
                                                 _ptr(_req(pred, torch.float32, "pred")),
                                                 _ptr(_req(directions.reshape(-1, 3), torch.float32, "directions")), w.shape[0],
                                                 w.shape[1], float(orientation_mult), float(pred_normal_mult),
                                                 _ptr(loss_orientation), _ptr(loss_pred_normal), _ptr(g), _stream(w.device)))
    return g


def normals_backward(fld: NerfactoFieldB200, pn_block: Tensor, origins: Tensor, directions: Tensor, ray_bins: Tensor,
                     grad_pred: Tensor, grad_pn_block: Tensor) -> Tensor:
    """d loss / d pred -> grad_pn_block (+=); returns d loss / d geo features [N,S,15] for `train_backward(grad_geo=...)`."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    b = _req(ray_bins, torch.float32, "ray_bins")
    n, S = b.shape[0], b.shape[1] - 1
    lib = _lib.load()
    need = int(lib.sgn_train_normals_ws_bytes(n, S))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=fld.device)
    ggeo = torch.empty((n, S, 15), dtype=torch.float32, device=fld.device)
    with torch.cuda.device(fld.device):
        _lib.check(lib.sgn_train_normals_backward(fld.handle, _ptr(_req(pn_block, torch.float32, "pn_block")), _ptr(o), _ptr(d), n,
                                                  _ptr(b), S, _ptr(_req(grad_pred, torch.float32, "grad_pred")),
                                                  _ptr(_req(grad_pn_block, torch.float32, "grad_pn_block")), _ptr(ggeo), _ptr(ws),
                                                  need, _stream(fld.device)))
    return ggeo


def normals_step(fld: NerfactoFieldB200, pn_block: Tensor, origins: Tensor, directions: Tensor, ray_bins: Tensor, weights: Tensor,
                 loss_orientation: Tensor, loss_pred_normal: Tensor, grad_pn_block: Tensor, orientation_mult: float = 1e-4,
                 pred_normal_mult: float = 1e-3) -> Tensor:
    """normals_forward + normal_losses + normals_backward in one pass over the samples (sgn_train_normals_step): loss_* [1]
    +=, grad_pn_block +=, returns d loss / d geo features [N,S,15]."""
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    b = _req(ray_bins, torch.float32, "ray_bins")
    n, S = b.shape[0], b.shape[1] - 1
    lib = _lib.load()
    need = int(lib.sgn_train_normals_ws_bytes(n, S))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device=fld.device)
    ggeo = torch.empty((n, S, 15), dtype=torch.float32, device=fld.device)
    with torch.cuda.device(fld.device):
        _lib.check(lib.sgn_train_normals_step(fld.handle, _ptr(_req(pn_block, torch.float32, "pn_block")), _ptr(o), _ptr(d), n, _ptr(b), S,
                                              _ptr(_req(weights, torch.float32, "weights")), float(orientation_mult),
                                              float(pred_normal_mult), _ptr(loss_orientation), _ptr(loss_pred_normal),
                                              _ptr(_req(grad_pn_block, torch.float32, "grad_pn_block")), _ptr(ggeo), _ptr(ws), need,
                                              _stream(fld.device)))
    return ggeo


class NerfactoTrainer(FieldTrainer):
    """The whole training step of `SIGNeRFModel` short of LPIPS and the normal regularisers: `get_outputs` while training
    + `get_loss_dict` (signerf/signerf.py:41-68) + Adam on `fields` and `proposal_networks` (signerf_config.py:43-50).
    `embedding` [num_images, 32]: the per-image appearance table (SIGNeRF re-initialises it when it loads the pretrained
    scene, signerf_pipeline.py:110-111); None trains with the folded mean embedding like `FieldTrainer`."""

    def __init__(self, fld: NerfactoFieldB200, embedding: Optional[Tensor] = None, counts: Tuple[int, int, int] = (256, 96, 48),
                 near: float = 0.05, far: float = 1000.0, interlevel_loss_mult: float = 1.0, distortion_loss_mult: float = 0.002,
                 pred_normals: Optional[Dict[str, Tensor]] = None, orientation_loss_mult: float = 1e-4,
                 pred_normal_loss_mult: float = 1e-3, **kw):
        """pred_normals: the `field.mlp_pred_normals.*` / `field.field_head_pred_normals.*` tensors of a model trained with
        `predict_normals=True` (signerf_config.py:33) - enables the orientation / pred-normal losses of signerf.py:69-80."""
        super().__init__(fld, **kw)
        if len(fld.prop_grids) != 2:
            raise ValueError("NerfactoTrainer needs a field with its two proposal networks")
        lib, dev = _lib.load(), fld.device
        self.counts, self.near, self.far = tuple(counts), near, far
        self.interlevel_mult, self.distortion_mult = interlevel_loss_mult, distortion_loss_mult
        n_prop = int(lib.sgn_prop_param_count())
        self.prop_tables = [g.table for g in fld.prop_grids]
        self.prop_mlps = []
        for l in range(2):
            ptr = C.c_void_p()
            _lib.check(lib.sgn_field_prop_params(fld.handle, l, C.byref(ptr)))
            with torch.cuda.device(dev):
                self.prop_mlps.append(torch.as_tensor(_DevicePtr(int(ptr.value), n_prop), device=dev))
        self.grad_prop_tables = [torch.zeros_like(t) for t in self.prop_tables]
        self.grad_prop_mlps = [torch.zeros(n_prop, dtype=torch.float32, device=dev) for _ in range(2)]
        for l in range(2):
            self.state[f"prop_table{l}"] = (torch.zeros_like(self.prop_tables[l]), torch.zeros_like(self.prop_tables[l]))
            self.state[f"prop_mlp{l}"] = (torch.zeros(n_prop, device=dev), torch.zeros(n_prop, device=dev))
        self.embedding = None if embedding is None else embedding.detach().to(dev, torch.float32).contiguous()
        if self.embedding is not None:
            self.grad_embedding = torch.zeros_like(self.embedding)
            self.state["embedding"] = (torch.zeros_like(self.embedding), torch.zeros_like(self.embedding))
        self.pn = None
        if pred_normals is not None:
            self.pn = pack_pred_normals(pred_normals, dev)
            self.grad_pn = torch.zeros_like(self.pn)
            self.state["pred_normals"] = (torch.zeros_like(self.pn), torch.zeros_like(self.pn))
        self.orientation_mult, self.pred_normal_mult = orientation_loss_mult, pred_normal_loss_mult
        self.loss_dict: Dict[str, Tensor] = {}

    def zero_grad(self) -> None:
        super().zero_grad()
        if self.pn is not None:
            self.grad_pn.zero_()
        for g in self.grad_prop_tables + self.grad_prop_mlps:
            g.zero_()
        self.grad_w_app.zero_()
        self.grad_b_head0.zero_()
        if self.embedding is not None:
            self.grad_embedding.zero_()

    def forward_backward(self, origins: Tensor, directions: Tensor, target_rgb: Tensor, jitter: Optional[Tensor] = None,
                         camera_indices: Optional[Tensor] = None,
                         extra_loss: Optional[Callable[[Tensor], Tensor]] = None, anneal: float = 1.0,
                         update_proposals: bool = True) -> Dict[str, Tensor]:
        """Gradients of rgb_loss + interlevel_loss + distortion_loss (+ extra_loss(rgb), the LPIPS hook) into every grad_*
        buffer; returns the loss dict (device scalars).  anneal: `proposal_anneal(step)`; update_proposals False = a step
        between two proposal updates (`ProposalUpdateSchedule`): the proposal densities carry no graph, the interlevel
        term is reported but moves nothing."""
        fld, dev = self.field, self.field.device
        self.zero_grad()
        smp = train_sample(fld, origins, directions, self.counts, self.near, self.far, jitter, anneal)
        per_image = self.embedding is not None and camera_indices is not None
        hb = appearance_bias(self.w_app, self.b_head0, self.embedding, camera_indices.to(dev, torch.int32)) if per_image else None
        rgb, _, saved = train_forward(fld, origins, directions, smp.euclid[2], hb)
        w_final = weights_from_density(smp.euclid[2], saved[0])
        loss_rgb, grad_rgb = rgb_loss(rgb, target_rgb.reshape(-1, 3).to(dev), self.use_l1)
        inter = torch.zeros(1, dtype=torch.float32, device=dev)
        dist_l = torch.zeros(1, dtype=torch.float32, device=dev)
        g_prop = [interlevel_loss(smp.spacing[2], w_final, smp.spacing[l], smp.weights[l], inter, self.interlevel_mult) for l in range(2)]
        g_wf = distortion_loss(smp.spacing[2], w_final, dist_l, self.distortion_mult)
        out = {"rgb_loss": loss_rgb, "interlevel_loss": inter, "distortion_loss": dist_l}
        if extra_loss is not None:
            leaf = rgb.detach().requires_grad_(True)
            extra = extra_loss(leaf)
            (g_extra,) = torch.autograd.grad(extra, leaf)
            grad_rgb = grad_rgb + g_extra
            out["lpips_loss"] = extra.detach().reshape(1)
        ghb = torch.zeros_like(hb) if per_image else None
        g_geo = None
        if self.pn is not None:      # predict_normals: signerf.py:69-80 on the detached final weights
            l_or = torch.zeros(1, dtype=torch.float32, device=dev)
            l_pn = torch.zeros(1, dtype=torch.float32, device=dev)
            g_geo = normals_step(fld, self.pn, origins, directions, smp.euclid[2], w_final, l_or, l_pn, self.grad_pn,
                                 self.orientation_mult, self.pred_normal_mult)
            out["orientation_loss"], out["pred_normal_loss"] = l_or, l_pn
        train_backward(fld, origins, directions, smp.euclid[2], saved, grad_rgb, self.grad_table, self.grad_mlp, g_wf, hb, ghb,
                       g_geo)
        if per_image:
            appearance_bias_backward(self.w_app, self.embedding, camera_indices.to(dev, torch.int32), ghb, self.grad_w_app,
                                     self.grad_b_head0, self.grad_embedding)
        for l in range(2 if update_proposals else 0):
            prop_backward(fld, l, origins, directions, smp.euclid[l], smp.sigma[l], g_prop[l], self.grad_prop_tables[l],
                          self.grad_prop_mlps[l])
        self._per_image = per_image
        self.loss_dict = out
        return out

    def state_dict(self) -> Dict[str, Tensor]:
        sd = super().state_dict()
        if self.embedding is not None:
            sd["field.embedding_appearance.embedding.weight"] = self.embedding.detach().clone()
        if self.pn is not None:
            v = pn_block_views(self.pn)
            for key, name in PN_NAMES.items():
                sd[name] = (v[key][:3] if key == "bh" else v[key]).detach().clone()
        for l in range(2):
            pre, m = f"proposal_networks.{l}.mlp_base", self.prop_mlps[l]
            sd[pre + ".encoding.hash_table"] = self.prop_tables[l].detach().clone()
            sd[pre + ".mlp.layers.0.weight"], sd[pre + ".mlp.layers.0.bias"] = m[:160].view(16, 10).clone(), m[160:176].clone()
            sd[pre + ".mlp.layers.1.weight"], sd[pre + ".mlp.layers.1.bias"] = m[176:192].view(1, 16).clone(), m[192:193].clone()
        return sd

    def all_gradients(self):
        """Every gradient buffer of the step (for a data-parallel all-reduce)."""
        g = [self.grad_table, self.grad_mlp, self.grad_w_app, self.grad_b_head0] + self.grad_prop_tables + self.grad_prop_mlps
        g += [self.grad_pn] if self.pn is not None else []
        return g + ([self.grad_embedding] if self.embedding is not None else [])

    def optimizer_step(self) -> None:
        lib, dev = _lib.load(), self.field.device
        if getattr(self, "_per_image", False):
            # per-image embeddings: W_app / b / E got their own gradients; the block's (folded) b_head0 slot is not a
            # parameter of this mode - it is rewritten from the mean embedding for the eval renderer below
            self.steps += 1
            mlp_block_views(self.grad_mlp)["b_head0"].zero_()
            todo = [("table", self.table, self.grad_table, self.table.numel()), ("mlp", self.mlp, self.grad_mlp, self.trainable),
                    ("w_app", self.w_app, self.grad_w_app, self.w_app.numel()),
                    ("b_head0", self.b_head0, self.grad_b_head0, self.b_head0.numel()),
                    ("embedding", self.embedding, self.grad_embedding, self.embedding.numel())]
            with torch.cuda.device(dev):
                for key, param, grad, n in todo:
                    m, v = self.state[key]
                    _lib.check(lib.sgn_adam_step(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), n, self.lr, self.betas[0],
                                                 self.betas[1], self.eps, self.steps, _stream(dev)))
            self.app_mean = self.embedding.mean(dim=0)
            mlp_block_views(self.mlp)["b_head0"].copy_(self.b_head0 + self.w_app @ self.app_mean)
        else:
            super().optimizer_step()
        with torch.cuda.device(dev):
            if self.pn is not None:
                m, v = self.state["pred_normals"]
                _lib.check(lib.sgn_adam_step(_ptr(self.pn), _ptr(self.grad_pn), _ptr(m), _ptr(v), self.pn.numel() - 1, self.lr,
                                             self.betas[0], self.betas[1], self.eps, self.steps, _stream(dev)))
            for l in range(2):
                for key, param, grad in ((f"prop_table{l}", self.prop_tables[l], self.grad_prop_tables[l]),
                                         (f"prop_mlp{l}", self.prop_mlps[l], self.grad_prop_mlps[l])):
                    m, v = self.state[key]
                    _lib.check(lib.sgn_adam_step(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(), self.lr, self.betas[0],
                                                 self.betas[1], self.eps, self.steps, _stream(dev)))

    def train_step(self, origins: Tensor, directions: Tensor, target_rgb: Tensor, jitter: Optional[Tensor] = None,
                   camera_indices: Optional[Tensor] = None, extra_loss: Optional[Callable[[Tensor], Tensor]] = None) -> Dict[str, Tensor]:
        """forward + losses + backward + Adam; jitter None draws the three per-ray uniforms on the device."""
        if jitter is None:
            jitter = torch.rand((3, origins.reshape(-1, 3).shape[0]), device=self.field.device)
        out = self.forward_backward(origins, directions, target_rgb, jitter, camera_indices, extra_loss)
        self.optimizer_step()
        return out


# ---------------------------------------------------------------------------------------------- behind nerfstudio's own optimizers
class _AttachGradients(torch.autograd.Function):
    """value = the summed loss of the step; the gradients were already computed by the CUDA backward - hand them to
    autograd as the gradients of the model's own parameters (scaled by the upstream gradient: GradScaler, loss weights)."""

    @staticmethod
    def forward(ctx, value: Tensor, grads, *params):
        ctx.grads = grads
        return value.clone()

    @staticmethod
    def backward(ctx, g: Tensor):
        return (None, None) + tuple(g * gr for gr in ctx.grads)


class FusedTrainingStep:
    """`SIGNeRFModel`'s training step - `get_outputs` while training + `get_loss_dict` (signerf.py:41-68) - on the CUDA
    kernels, BEHIND the reference's own parameters and optimizers: `loss_dict(...)` returns the reference's loss dict whose
    sum back-propagates (one autograd node) into the `.grad` of the model's own tensors, so nerfstudio's `Optimizers`
    (Adam + ExponentialDecay per group, signerf_config.py:43-58), GradScaler and checkpointing stay exactly as they are.

    `model`: a torch-fallback nerfacto / SIGNeRF model (`pipeline.model`) on the GPU - anything whose `state_dict` carries
    nerfstudio's parameter names.  The hash tables and the appearance table are used IN PLACE (no copies: the optimizer's
    updates are seen at once); the small MLP tensors are re-uploaded into the field's parameter block at every call."""

    def __init__(self, model: torch.nn.Module, counts: Tuple[int, int, int] = (256, 96, 48), near: float = 0.05, far: float = 1000.0,
                 average_init_density: float = 0.01, use_l1: bool = True, interlevel_loss_mult: float = 1.0,
                 distortion_loss_mult: float = 0.002, predict_normals: Optional[bool] = None,
                 orientation_loss_mult: float = 1e-4, pred_normal_loss_mult: float = 1e-3):
        from .plugin.model import FusedNerfactoGraph
        self.model = model
        sd = dict(model.state_dict(keep_vars=True))

        def find(name: str) -> Tensor:
            for cand in [name] + [name.replace(a, b) for a, b in FieldTrainer._ALT_NAMES if a in name]:
                if cand in sd:
                    return sd[cand]
            raise KeyError(f"the model has no parameter {name}")

        names = ["field.mlp_base.encoding.hash_table"]
        names += [f"field.mlp_base.mlp.layers.{i}.{p}" for i in range(2) for p in ("weight", "bias")]
        names += [f"field.mlp_head.layers.{i}.{p}" for i in range(3) for p in ("weight", "bias")]
        names += ["field.embedding_appearance.embedding.weight"]
        for l in range(2):
            names += [f"proposal_networks.{l}.mlp_base.encoding.hash_table"]
            names += [f"proposal_networks.{l}.mlp_base.mlp.layers.{i}.{p}" for i in range(2) for p in ("weight", "bias")]
        self.params: Dict[str, Tensor] = {n: find(n) for n in names}
        # predict_normals=True (signerf_config.py:33): the prediction branch's tensors, when the model has them
        self.pn_names = {}
        if predict_normals is None:
            predict_normals = any("mlp_pred_normals" in k for k in sd)
        if predict_normals:
            for key, n in PN_NAMES.items():
                self.params[n] = find(n)
                self.pn_names[key] = n
        for n, t in self.params.items():
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{n} must be a contiguous fp32 CUDA tensor (the kernels read the model's tensors in place)")
        detached = {n: t.detach() for n, t in self.params.items()}
        emb = detached["field.embedding_appearance.embedding.weight"]
        graph = FusedNerfactoGraph.from_state_dict(detached, device=emb.device, average_init_density=average_init_density,
                                                   near_plane=near, far_plane=far)
        self.field = graph.field
        if self.field.grid.table.data_ptr() != detached["field.mlp_base.encoding.hash_table"].data_ptr():
            raise RuntimeError("the hash table was copied: it must be usable in place")
        pn = {n: detached[n] for n in self.pn_names.values()} if self.pn_names else None
        self.trainer = NerfactoTrainer(self.field, embedding=emb, counts=counts, near=near, far=far, use_l1=use_l1,
                                       interlevel_loss_mult=interlevel_loss_mult, distortion_loss_mult=distortion_loss_mult,
                                       pred_normals=pn, orientation_loss_mult=orientation_loss_mult,
                                       pred_normal_loss_mult=pred_normal_loss_mult)
        self.trainer.embedding = emb                                   # the model's table itself, not a copy
        self.update_schedule = ProposalUpdateSchedule()

    def _sync_from_model(self) -> None:
        """The model's current MLP tensors -> the field's parameter blocks (tables / embedding are shared storage)."""
        p, tr = self.params, self.trainer
        v = mlp_block_views(tr.mlp)
        with torch.no_grad():
            for i, (w, b) in enumerate((("w_base0", "b_base0"), ("w_base1", "b_base1"))):
                v[w].copy_(p[f"field.mlp_base.mlp.layers.{i}.weight"])
                v[b].copy_(p[f"field.mlp_base.mlp.layers.{i}.bias"])
            h0 = p["field.mlp_head.layers.0.weight"]
            v["w_head0"][:, :16].copy_(h0[:, :16])
            v["w_head0"][:, 16].zero_()
            v["w_head0"][:, 17:32].copy_(h0[:, 16:31])
            tr.w_app.copy_(h0[:, 31:63])
            tr.b_head0.copy_(p["field.mlp_head.layers.0.bias"])
            v["w_head1"].copy_(p["field.mlp_head.layers.1.weight"]), v["b_head1"].copy_(p["field.mlp_head.layers.1.bias"])
            v["w_head2"].copy_(p["field.mlp_head.layers.2.weight"]), v["b_head2"][:3].copy_(p["field.mlp_head.layers.2.bias"])
            tr.app_mean = tr.embedding.mean(dim=0)
            v["b_head0"].copy_(tr.b_head0 + tr.w_app @ tr.app_mean)    # folded bias: what the eval renderer uses
            if self.pn_names:
                pv = pn_block_views(tr.pn)
                for key, n in self.pn_names.items():
                    (pv[key][:3] if key == "bh" else pv[key]).copy_(p[n])
            for l in range(2):
                pre, m = f"proposal_networks.{l}.mlp_base.mlp.layers", tr.prop_mlps[l]
                m[:160].copy_(p[pre + ".0.weight"].reshape(-1)), m[160:176].copy_(p[pre + ".0.bias"])
                m[176:192].copy_(p[pre + ".1.weight"].reshape(-1)), m[192:193].copy_(p[pre + ".1.bias"])

    def _gradients(self) -> Dict[str, Tensor]:
        tr = self.trainer
        g = nerfstudio_gradients(tr.grad_table, tr.grad_mlp, tr.app_mean)
        g["field.mlp_head.layers.0.weight"] = torch.cat([g["field.mlp_head.layers.0.weight"][:, :31], tr.grad_w_app], dim=1)
        g["field.mlp_head.layers.0.bias"] = tr.grad_b_head0
        g["field.embedding_appearance.embedding.weight"] = tr.grad_embedding
        if self.pn_names:
            gv = pn_block_views(tr.grad_pn)
            for key, n in self.pn_names.items():
                g[n] = gv[key][:3] if key == "bh" else gv[key]
        for l in range(2):
            pre, m = f"proposal_networks.{l}.mlp_base", tr.grad_prop_mlps[l]
            g[pre + ".encoding.hash_table"] = tr.grad_prop_tables[l]
            g[pre + ".mlp.layers.0.weight"], g[pre + ".mlp.layers.0.bias"] = m[:160].view(16, 10), m[160:176]
            g[pre + ".mlp.layers.1.weight"], g[pre + ".mlp.layers.1.bias"] = m[176:192].view(1, 16), m[192:193]
        return g

    def loss_dict(self, origins: Tensor, directions: Tensor, target_rgb: Tensor, camera_indices: Tensor,
                  jitter: Optional[Tensor] = None, extra_loss: Optional[Callable[[Tensor], Tensor]] = None,
                  step: Optional[int] = None) -> Dict[str, Tensor]:
        """One batch -> {"rgb_loss", "interlevel_loss", "distortion_loss" (, "orientation_loss", "pred_normal_loss", "lpips_loss")}.  `sum(values).backward()` leaves
        the step's gradients in the model's parameters; the individual entries carry the reference's values (only
        "rgb_loss" holds the autograd node, for the sum of all terms)."""
        self._sync_from_model()
        if jitter is None:
            jitter = torch.rand((3, origins.reshape(-1, 3).shape[0]), device=self.field.device)
        # step given: nerfacto's training callbacks - proposal-weight annealing and the proposal update schedule
        anneal = proposal_anneal(step) if step is not None else 1.0
        update = self.update_schedule(step) if step is not None else True
        out = self.trainer.forward_backward(origins, directions, target_rgb, jitter, camera_indices, extra_loss, anneal, update)
        grads = self._gradients()
        order = list(self.params)
        others = sum(v for k, v in out.items() if k != "rgb_loss")
        total = _AttachGradients.apply((out["rgb_loss"] + others).reshape(()), [grads[n].clone() for n in order],
                                       *[self.params[n] for n in order])
        res = {k: v.reshape(()).detach() for k, v in out.items()}
        res["rgb_loss"] = total - others.reshape(()).detach()          # value = rgb loss, gradient = that of the whole sum
        return res

    def refresh_renderer(self) -> None:
        """After optimizer steps: the model's tensors -> parameter block -> the eval renderer's tensor-core fragments."""
        self._sync_from_model()
        self.trainer.refresh_renderer()
