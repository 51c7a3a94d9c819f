"""ORACLE (test infrastructure, NOT product code) — CPU/PyTorch fp32 restatement of the
nerfstudio==1.0.2 `implementation="torch"` nerfacto eval path that SIGNeRF's
`DatasetGenerator.render_camera` drives (reference call sites:
signerf/datasetgenerator/datasetgenerator.py:691 `camera.generate_rays(...)` and :694
`graph.get_outputs_for_camera_ray_bundle(...)`; only outputs["rgb"] / outputs["depth"] are
consumed, :700-701).

PARITY UNPINNED for this file: nerfstudio 1.0.2 (pyproject.toml:6 of the reference) is a
third-party dependency whose source is absent from /root/reference and is not installable
here (no wheel, no network), and the reference ships no tests / golden vectors for this
path.  The algorithm below restates the published nerfstudio 1.0.x
code (cameras/cameras.py, cameras/rays.py, model_components/{scene_colliders,ray_samplers,
renderers}.py, field_components/{spatial_distortions,encodings,mlp,activations}.py,
fields/{nerfacto_field,density_fields}.py, models/nerfacto.py) op-for-op in fp32 so that
torch's own rounding order is preserved.  Pieces of the path that DO live in
/root/reference (AABB slab test, pose generator, uint8 quantisation, cv2 dilation) are
pinned separately in oracle/sheet_ref.py against fixtures generated from the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

HASH_PRIMES = (1, 2654435761, 805459861)  # nerfstudio encodings.py HashEncoding.hash_fn


# --------------------------------------------------------------------------- field params
def hash_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """HashEncoding.__init__: `torch.floor(min_res * growth_factor ** levels)`.

    growth_factor is a numpy float64 scalar and `levels` an int64 tensor, so the power is
    evaluated in torch's default dtype (float32) — 16 levels 16→2048 end at 2047, not 2048
    (SURVEY.md §7 "fp32 scalings quirk").  Never recompute this elsewhere: pass this table.
    """
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth**levels)


class HashEncodingRef(nn.Module):
    """nerfstudio HashEncoding, torch fallback (`pytorch_fwd`)."""

    def __init__(self, num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=19,
                 features_per_level=2, hash_init_scale=1e-3):
        super().__init__()
        self.num_levels = num_levels
        self.features_per_level = features_per_level
        self.hash_table_size = 2**log2_hashmap_size
        self.register_buffer("scalings", hash_scalings(num_levels, min_res, max_res))
        self.hash_offset = torch.arange(num_levels) * self.hash_table_size
        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1
        table *= hash_init_scale
        self.hash_table = nn.Parameter(table)

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def hash_fn(self, in_tensor: Tensor) -> Tensor:
        # int32 * int64 tensor -> int64 (type promotion); xor; python-style modulo; level offset
        in_tensor = in_tensor * torch.tensor(HASH_PRIMES).to(in_tensor.device)
        x = torch.bitwise_xor(in_tensor[..., 0], in_tensor[..., 1])
        x = torch.bitwise_xor(x, in_tensor[..., 2])
        x %= self.hash_table_size
        x += self.hash_offset.to(x.device)
        return x

    def corner_indices(self, in_tensor: Tensor) -> Tuple[Tensor, Tensor]:
        """Return ([...,L,8] int64 table rows, [...,L,3] trilinear offsets) for positions in [0,1]."""
        in_tensor = in_tensor[..., None, :]
        scaled = in_tensor * self.scalings.view(-1, 1).to(in_tensor.device)
        scaled_c = torch.ceil(scaled).type(torch.int32)
        scaled_f = torch.floor(scaled).type(torch.int32)
        offset = scaled - scaled_f
        c, f = scaled_c, scaled_f

        def pick(a, b, d):
            return torch.cat([a[..., 0:1], b[..., 1:2], d[..., 2:3]], dim=-1)

        hashed = [
            self.hash_fn(c),               # 0: c c c
            self.hash_fn(pick(c, f, c)),   # 1: c f c
            self.hash_fn(pick(f, f, c)),   # 2: f f c
            self.hash_fn(pick(f, c, c)),   # 3: f c c
            self.hash_fn(pick(c, c, f)),   # 4: c c f
            self.hash_fn(pick(c, f, f)),   # 5: c f f
            self.hash_fn(f),               # 6: f f f
            self.hash_fn(pick(f, c, f)),   # 7: f c f
        ]
        return torch.stack(hashed, dim=-1), offset

    def forward(self, in_tensor: Tensor) -> Tensor:
        assert in_tensor.shape[-1] == 3
        hashed, offset = self.corner_indices(in_tensor)
        f = [self.hash_table[hashed[..., i]] for i in range(8)]  # each [..., L, F]
        ox, oy, oz = offset[..., 0:1], offset[..., 1:2], offset[..., 2:3]
        f_03 = f[0] * ox + f[3] * (1 - ox)
        f_12 = f[1] * ox + f[2] * (1 - ox)
        f_56 = f[5] * ox + f[6] * (1 - ox)
        f_47 = f[4] * ox + f[7] * (1 - ox)
        f0312 = f_03 * oy + f_12 * (1 - oy)
        f4756 = f_47 * oy + f_56 * (1 - oy)
        encoded = f0312 * oz + f4756 * (1 - oz)
        return torch.flatten(encoded, start_dim=-2, end_dim=-1)


class MLPRef(nn.Module):
    """nerfstudio field_components/mlp.py MLP, torch path: biased nn.Linear, ReLU between."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: int,
                 out_activation: Optional[nn.Module] = None):
        super().__init__()
        dims = [in_dim] + [layer_width] * (num_layers - 1) + [out_dim]
        self.layers = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))
        self.out_activation = out_activation

    def forward(self, x: Tensor) -> Tensor:
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i < len(self.layers) - 1:
                x = torch.relu(x)
        if self.out_activation is not None:
            x = self.out_activation(x)
        return x


def contract_linf(x: Tensor) -> Tensor:
    """SceneContraction(order=inf).forward."""
    mag = torch.linalg.norm(x, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, x, (2 - (1 / mag)) * (x / mag))


def sh_components_deg4(directions: Tensor) -> Tensor:
    """nerfstudio utils/math.py components_from_spherical_harmonics(levels=4)."""
    components = torch.zeros((*directions.shape[:-1], 16), device=directions.device)
    x, y, z = directions[..., 0], directions[..., 1], directions[..., 2]
    xx, yy, zz = x**2, y**2, z**2
    components[..., 0] = 0.28209479177387814
    components[..., 1] = 0.4886025119029199 * y
    components[..., 2] = 0.4886025119029199 * z
    components[..., 3] = 0.4886025119029199 * x
    components[..., 4] = 1.0925484305920792 * x * y
    components[..., 5] = 1.0925484305920792 * y * z
    components[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
    components[..., 7] = 1.0925484305920792 * x * z
    components[..., 8] = 0.5462742152960396 * (xx - yy)
    components[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
    components[..., 10] = 2.890611442640554 * x * y * z
    components[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
    components[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
    components[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
    components[..., 14] = 1.445305721320277 * z * (xx - yy)
    components[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return components


class DensityFieldRef(nn.Module):
    """fields/density_fields.py HashMLPDensityField (use_linear=False): proposal network."""

    def __init__(self, num_levels=5, max_res=128, base_res=16, log2_hashmap_size=17,
                 features_per_level=2, hidden_dim=16, num_layers=2, average_init_density=1.0):
        super().__init__()
        self.average_init_density = average_init_density
        self.encoding = HashEncodingRef(num_levels, base_res, max_res, log2_hashmap_size, features_per_level)
        self.mlp = MLPRef(self.encoding.get_out_dim(), num_layers, hidden_dim, 1)

    def density(self, positions: Tensor) -> Tensor:
        positions = contract_linf(positions)
        positions = (positions + 2.0) / 4.0
        selector = ((positions > 0.0) & (positions < 1.0)).all(dim=-1)
        positions = positions * selector[..., None]
        h = self.mlp(self.encoding(positions.view(-1, 3))).view(*positions.shape[:-1], -1)
        density = self.average_init_density * torch.exp(h)  # trunc_exp forward == exp
        return density * selector[..., None]


class NerfactoFieldRef(nn.Module):
    """fields/nerfacto_field.py NerfactoField, eval, no normals / transients / semantics."""

    def __init__(self, num_images=30, num_levels=16, base_res=16, max_res=2048, log2_hashmap_size=19,
                 features_per_level=2, hidden_dim=64, geo_feat_dim=15, hidden_dim_color=64,
                 appearance_embedding_dim=32, average_init_density=0.01, use_pred_normals=False):
        super().__init__()
        self.geo_feat_dim = geo_feat_dim
        self.average_init_density = average_init_density
        self.appearance_embedding_dim = appearance_embedding_dim
        self.encoding = HashEncodingRef(num_levels, base_res, max_res, log2_hashmap_size, features_per_level)
        self.mlp_base = MLPRef(self.encoding.get_out_dim(), 2, hidden_dim, 1 + geo_feat_dim)
        self.embedding_appearance = nn.Embedding(num_images, appearance_embedding_dim)
        self.mlp_head = MLPRef(16 + geo_feat_dim + appearance_embedding_dim, 3, hidden_dim_color, 3,
                               out_activation=nn.Sigmoid())
        # predict_normals=True (signerf_config.py:33): NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=1)
        # of the RAW positions (12 values) | geo features -> MLP(3 layers, 64 wide, out 64, no out activation) ->
        # PredNormalsFieldHead = Linear(64, 3) + Tanh, normalised (fields/nerfacto_field.py, field_components/field_heads.py)
        self.use_pred_normals = use_pred_normals
        if use_pred_normals:
            self.mlp_pred_normals = MLPRef(geo_feat_dim + 12, 3, 64, 64)
            self.field_head_pred_normals = nn.Linear(64, 3)

    def contracted_positions(self, positions: Tensor) -> Tuple[Tensor, Tensor]:
        positions = contract_linf(positions)
        positions = (positions + 2.0) / 4.0
        selector = ((positions > 0.0) & (positions < 1.0)).all(dim=-1)
        return positions * selector[..., None], selector

    def get_density(self, positions: Tensor) -> Tuple[Tensor, Tensor]:
        p, selector = self.contracted_positions(positions)
        h = self.mlp_base(self.encoding(p.view(-1, 3))).view(*p.shape[:-1], -1)
        density_before_activation, base_mlp_out = torch.split(h, [1, self.geo_feat_dim], dim=-1)
        density = self.average_init_density * torch.exp(density_before_activation)
        return density * selector[..., None], base_mlp_out

    def get_rgb(self, directions: Tensor, density_embedding: Tensor) -> Tensor:
        """directions [..., S, 3] already broadcast per sample (frustums.directions)."""
        d = (directions + 1.0) / 2.0  # get_normalized_directions
        d = sh_components_deg4(d.view(-1, 3))  # torch-path SHEncoding is fed the [0,1] tensor as-is
        app = torch.ones((*directions.shape[:-1], self.appearance_embedding_dim)) * self.embedding_appearance.weight.mean(dim=0)
        h = torch.cat([d, density_embedding.reshape(-1, self.geo_feat_dim),
                       app.view(-1, self.appearance_embedding_dim)], dim=-1)
        return self.mlp_head(h).view(*directions.shape[:-1], -1)


def nerf_encoding_2freq(x: Tensor) -> Tensor:
    """NeRFEncoding.pytorch_fwd, in_dim 3, frequencies 2^0 and 2^1, include_input=False: sin of [2 pi x f | 2 pi x f + pi/2]."""
    scaled = (2 * torch.pi * x)[..., None] * (2 ** torch.linspace(0.0, 1.0, 2))
    scaled = scaled.view(*scaled.shape[:-2], -1)
    return torch.sin(torch.cat([scaled, scaled + torch.pi / 2.0], dim=-1))


def normals_and_pred_normals(field: "NerfactoFieldRef", positions: Tensor):
    """Field.forward(compute_normals=True) + the PRED_NORMALS head.  -> (density [..,1], geo, analytic normals [..,3]
    WITHOUT a graph (Field.get_normals: autograd.grad of the density logit w.r.t. the contracted sample locations,
    retain_graph only, then -normalize), predicted normals [..,3] with a graph into mlp_pred_normals, its head and - through
    the geo features - the base MLP and the hash table)."""
    p, selector = field.contracted_positions(positions)
    p = p.detach().requires_grad_(True)
    with torch.enable_grad():
        h = field.mlp_base(field.encoding(p.view(-1, 3))).view(*p.shape[:-1], -1)
        logit, geo = torch.split(h, [1, field.geo_feat_dim], dim=-1)
        density = field.average_init_density * torch.exp(logit) * selector[..., None]
        (g,) = torch.autograd.grad(logit, p, grad_outputs=torch.ones_like(logit), retain_graph=True)
    normals = -torch.nn.functional.normalize(g, dim=-1)
    inp = torch.cat([nerf_encoding_2freq(positions.reshape(-1, 3)), geo.reshape(-1, field.geo_feat_dim)], dim=-1)
    x = field.mlp_pred_normals(inp).view(*positions.shape[:-1], -1)
    pred = torch.nn.functional.normalize(torch.tanh(field.field_head_pred_normals(x)), dim=-1)
    return density, geo, normals, pred


def orientation_loss(weights: Tensor, normals: Tensor, viewdirs: Tensor) -> Tensor:
    """losses.py orientation_loss (Ref-NeRF): back-facing normals along the ray, per ray."""
    n_dot_v = (normals * (-viewdirs)[..., None, :]).sum(dim=-1)
    return (weights[..., 0] * torch.fmin(torch.zeros_like(n_dot_v), n_dot_v) ** 2).sum(dim=-1)


def pred_normal_loss(weights: Tensor, normals: Tensor, pred_normals: Tensor) -> Tensor:
    """losses.py pred_normal_loss: predicted against analytic normals, per ray."""
    return (weights[..., 0] * (1.0 - torch.sum(normals * pred_normals, dim=-1))).sum(dim=-1)


class NerfactoRef(nn.Module):
    """models/nerfacto.py populate_modules with SIGNeRF's overrides (signerf_config.py:32-35)."""

    def __init__(self, num_images=30, average_init_density=0.01, near=0.05, far=1000.0,
                 num_proposal_samples=(256, 96), num_nerf_samples=48,
                 proposal_net_args=({"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128},
                                    {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256}),
                 log2_hashmap_size=19, num_levels=16, max_res=2048, predict_normals=False):
        super().__init__()
        self.near, self.far = near, far
        self.num_proposal_samples = tuple(num_proposal_samples)
        self.num_nerf_samples = num_nerf_samples
        self.field = NerfactoFieldRef(num_images=num_images, average_init_density=average_init_density,
                                      log2_hashmap_size=log2_hashmap_size, num_levels=num_levels, max_res=max_res,
                                      use_pred_normals=predict_normals)
        self.proposal_networks = nn.ModuleList(
            DensityFieldRef(**a, average_init_density=average_init_density) for a in proposal_net_args)


# --------------------------------------------------------------------------- rays
@dataclass
class RaysRef:
    origins: Tensor      # [N,3]
    directions: Tensor   # [N,3] unit
    pixel_area: Tensor   # [N,1]
    directions_norm: Tensor  # [N,1]


def generate_rays(c2w: Tensor, fx: float, fy: float, cx: float, cy: float, width: int, height: int) -> RaysRef:
    """Cameras.generate_rays(camera_indices=0) for a perspective camera without distortion.

    Ray id = y*W + x (row-major, `indexing="ij"`), pixel centre offset 0.5.
    """
    c2w = c2w[:3, :4].to(torch.float32)
    ys, xs = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    y = ys.to(torch.float32) + 0.5
    x = xs.to(torch.float32) + 0.5
    fx_t, fy_t, cx_t, cy_t = (torch.tensor(v, dtype=torch.float32) for v in (fx, fy, cx, cy))
    coord = torch.stack([(x - cx_t) / fx_t, -(y - cy_t) / fy_t], -1)
    coord_x = torch.stack([(x - cx_t + 1) / fx_t, -(y - cy_t) / fy_t], -1)
    coord_y = torch.stack([(x - cx_t) / fx_t, -(y - cy_t + 1) / fy_t], -1)
    coord_stack = torch.stack([coord, coord_x, coord_y], dim=0)  # [3,H,W,2]
    directions_stack = torch.empty((3, height, width, 3), dtype=torch.float32)
    directions_stack[..., 0] = coord_stack[..., 0]
    directions_stack[..., 1] = coord_stack[..., 1]
    directions_stack[..., 2] = -1.0
    rotation = c2w[:3, :3]
    directions_stack = torch.sum(directions_stack[..., None, :] * rotation, dim=-1)
    norm = torch.maximum(torch.linalg.vector_norm(directions_stack, dim=-1, keepdim=True),
                         torch.tensor([1e-6]))  # camera_utils.normalize_with_norm
    directions_stack = directions_stack / norm
    origins = c2w[:3, 3].expand(height, width, 3)
    directions = directions_stack[0]
    dx = torch.sqrt(torch.sum((directions - directions_stack[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((directions - directions_stack[2]) ** 2, dim=-1))
    return RaysRef(origins.reshape(-1, 3).contiguous(), directions.reshape(-1, 3).contiguous(),
                   (dx * dy).reshape(-1, 1), norm[0].reshape(-1, 1))


# --------------------------------------------------------------------------- samplers
def spacing_fn(x: Tensor) -> Tensor:  # UniformLinDispPiecewiseSampler
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def spacing_fn_inv(x: Tensor) -> Tensor:
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


@dataclass
class SamplesRef:
    starts: Tensor          # [N,S,1] euclid
    ends: Tensor
    spacing_starts: Tensor  # [N,S,1]
    spacing_ends: Tensor


def make_to_euclid(nears: Tensor, fars: Tensor):
    s_near, s_far = spacing_fn(nears), spacing_fn(fars)
    return lambda x: spacing_fn_inv(x * s_far + (1 - x) * s_near)


def initial_samples(num_rays: int, num_samples: int, to_euclid) -> SamplesRef:
    """SpacedSampler.generate_ray_samples, eval (no stratification)."""
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, ...]
    euclid = to_euclid(bins)  # nears/fars are [N,1] -> [N,S+1]
    bins = bins.expand(num_rays, -1)
    return SamplesRef(euclid[..., :-1, None], euclid[..., 1:, None], bins[..., :-1, None], bins[..., 1:, None])


def flat_bin_edges(num_samples: int, near: float, far: float) -> Tensor:
    """Euclidean bin edges [S+1] shared by every ray when near/far are the collider constants."""
    nears = torch.full((1, 1), near, dtype=torch.float32)
    fars = torch.full((1, 1), far, dtype=torch.float32)
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, ...]
    return make_to_euclid(nears, fars)(bins)[0]


def pdf_resample(samples: SamplesRef, weights: Tensor, num_samples: int, to_euclid,
                 histogram_padding: float = 0.01, eps: float = 1e-5, jitter: Optional[Tensor] = None) -> SamplesRef:
    """PDFSampler.generate_ray_samples, include_original=False.  jitter None: the eval branch (bin centres);
    jitter [N,1] in [0,1): the training branch with single_jitter=True, `u + rand((N,1)) / num_bins`, with the random
    numbers handed in so that the CUDA path can be checked on the same draws."""
    num_bins = num_samples + 1
    weights = weights[..., 0] + histogram_padding
    weights_sum = torch.sum(weights, dim=-1, keepdim=True)
    padding = torch.relu(eps - weights_sum)
    weights = weights + padding / weights.shape[-1]
    weights_sum += padding
    pdf = weights / weights_sum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins)
    if jitter is None:
        u = u + 1.0 / (2 * num_bins)
        u = u.expand(size=(*cdf.shape[:-1], num_bins)).contiguous()
    else:
        u = (u.expand(size=(*cdf.shape[:-1], num_bins)) + jitter / num_bins).contiguous()
    existing_bins = torch.cat([samples.spacing_starts[..., 0], samples.spacing_ends[..., -1:, 0]], dim=-1)
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, existing_bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, existing_bins.shape[-1] - 1)
    cdf_g0 = torch.gather(cdf, -1, below)
    bins_g0 = torch.gather(existing_bins, -1, below)
    cdf_g1 = torch.gather(cdf, -1, above)
    bins_g1 = torch.gather(existing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
    bins = bins_g0 + t * (bins_g1 - bins_g0)
    euclid = to_euclid(bins)
    return SamplesRef(euclid[..., :-1, None], euclid[..., 1:, None], bins[..., :-1, None], bins[..., 1:, None])


def positions_of(rays_o: Tensor, rays_d: Tensor, s: SamplesRef) -> Tensor:
    """Frustums.get_positions: origins + directions * (starts + ends) / 2."""
    return rays_o[:, None, :] + rays_d[:, None, :] * (s.starts + s.ends) / 2


def get_weights(s: SamplesRef, densities: Tensor) -> Tensor:
    """RaySamples.get_weights."""
    deltas = s.ends - s.starts
    delta_density = deltas * densities
    alphas = 1 - torch.exp(-delta_density)
    transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
    transmittance = torch.cat([torch.zeros((*transmittance.shape[:1], 1, 1)), transmittance], dim=-2)
    transmittance = torch.exp(-transmittance)
    weights = alphas * transmittance
    return torch.nan_to_num(weights)


# --------------------------------------------------------------------------- renderers
def render_rgb_last_sample(rgb: Tensor, weights: Tensor) -> Tensor:
    """RGBRenderer(background_color="last_sample"), eval: nan_to_num in, clamp out."""
    rgb = torch.nan_to_num(rgb)
    comp_rgb = torch.sum(weights * rgb, dim=-2)
    accumulated_weight = torch.sum(weights, dim=-2)
    comp_rgb = comp_rgb + rgb[..., -1, :] * (1.0 - accumulated_weight)
    return torch.clamp(comp_rgb, min=0.0, max=1.0)


def render_depth_median(weights: Tensor, s: SamplesRef) -> Tensor:
    """DepthRenderer(method="median")."""
    steps = (s.starts + s.ends) / 2
    cumulative_weights = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1)) * 0.5
    median_index = torch.searchsorted(cumulative_weights, split, side="left")
    median_index = torch.clamp(median_index, 0, steps.shape[-2] - 1)
    return torch.gather(steps[..., 0], dim=-1, index=median_index)


# --------------------------------------------------------------------------- model forward
@torch.no_grad()
def render_rays(model: NerfactoRef, rays_o: Tensor, rays_d: Tensor, mode: str = "flat",
                num_samples: Optional[int] = None, return_aux: bool = False) -> Dict[str, Tensor]:
    """NerfactoModel.get_outputs in eval for one chunk of rays.

    mode="flat":    BASELINE flat mode: the initial piecewise sampler with S bins feeds
                    the main field directly, no proposal networks.
    mode="cascade": the faithful 256 -> 96 -> 48 ProposalNetworkSampler (anneal = 1).
    """
    n = rays_o.shape[0]
    nears = torch.ones_like(rays_o[..., 0:1]) * model.near  # NearFarCollider
    fars = torch.ones_like(rays_o[..., 0:1]) * model.far
    to_euclid = make_to_euclid(nears, fars)
    aux = {}
    if mode == "flat":
        S = num_samples if num_samples is not None else model.num_nerf_samples
        samples = initial_samples(n, S, to_euclid)
    elif mode == "cascade":
        counts = list(model.num_proposal_samples) + [model.num_nerf_samples]
        samples = initial_samples(n, counts[0], to_euclid)
        for i, net in enumerate(model.proposal_networks):
            density = net.density(positions_of(rays_o, rays_d, samples))
            w = get_weights(samples, density)
            annealed = torch.pow(w, 1.0)
            samples = pdf_resample(samples, annealed, counts[i + 1], to_euclid)
            if return_aux:
                aux[f"prop_weights_{i}"] = w
    else:
        raise ValueError(mode)
    positions = positions_of(rays_o, rays_d, samples)
    density, geo = model.field.get_density(positions)
    dirs = rays_d[:, None, :].expand(-1, positions.shape[1], -1)
    rgb = model.field.get_rgb(dirs, geo)
    weights = get_weights(samples, density)
    out = {
        "rgb": render_rgb_last_sample(rgb, weights),
        "depth": render_depth_median(weights, samples),
        "accumulation": torch.sum(weights, dim=-2),
    }
    if return_aux:
        aux.update(positions=positions, density=density, rgb_samples=rgb, weights=weights,
                   starts=samples.starts, ends=samples.ends)
        out["aux"] = aux
    return out


@torch.no_grad()
def render_view(model: NerfactoRef, c2w: Tensor, fx: float, fy: float, cx: float, cy: float,
                width: int, height: int, mode: str = "flat", num_samples: Optional[int] = None,
                chunk: int = 1 << 15) -> Dict[str, Tensor]:
    """Model.get_outputs_for_camera_ray_bundle: row-major chunks of `eval_num_rays_per_chunk`
    (signerf_config.py:32 sets 1<<15), concatenated and viewed as (H, W, C)."""
    rays = generate_rays(c2w, fx, fy, cx, cy, width, height)
    outs: Dict[str, List[Tensor]] = {}
    for i in range(0, rays.origins.shape[0], chunk):
        o = render_rays(model, rays.origins[i:i + chunk], rays.directions[i:i + chunk], mode, num_samples)
        for k, v in o.items():
            outs.setdefault(k, []).append(v)
    res = {k: torch.cat(v).view(height, width, -1) for k, v in outs.items()}
    res["origins"] = rays.origins.view(height, width, 3)
    res["directions"] = rays.directions.view(height, width, 3)
    return res


def make_model(seed: int = 0, dense: bool = False, table_scale: Optional[float] = None,
               density_gain: Optional[float] = None, **kw) -> NerfactoRef:
    """Benchmark field: seeded random init; `dense` raises the density logit bias to +6
    (sigma ~ 4) so compositing is not degenerate; `table_scale` rescales the hash tables from the
    1e-3 init to a trained-like magnitude so density/colour vary in space (parity tests)."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = NerfactoRef(**kw)
    torch.random.set_rng_state(gen_state)
    if table_scale is not None:
        with torch.no_grad():
            for enc in [model.field.encoding] + [p.encoding for p in model.proposal_networks]:
                enc.hash_table.mul_(table_scale / 1e-3)
    if density_gain is not None:  # spatially varying density logit (test fields only)
        with torch.no_grad():
            model.field.mlp_base.layers[-1].weight[0].mul_(density_gain)
            for p in model.proposal_networks:
                p.mlp.layers[-1].weight.mul_(density_gain)
    if dense:
        with torch.no_grad():
            model.field.mlp_base.layers[-1].bias[0] = 6.0
            for p in model.proposal_networks:
                p.mlp.layers[-1].bias[0] = 6.0
    return model.eval()


# --------------------------------------------------------------------------- training step
def render_rays_train(model: NerfactoRef, rays_o: Tensor, rays_d: Tensor, num_samples: int) -> Tensor:
    """NerfactoModel.get_outputs WHILE TRAINING, restricted to what signerf_b200/train.py covers: the main field on the
    flat piecewise bins, autograd enabled.  RGBRenderer only applies nan_to_num / clamp when `not self.training`, so rgb is
    the raw composite.  -> rgb [N,3] with a graph back to hash table and MLP parameters."""
    n = rays_o.shape[0]
    nears = torch.ones_like(rays_o[..., 0:1]) * model.near
    fars = torch.ones_like(rays_o[..., 0:1]) * model.far
    samples = initial_samples(n, num_samples, make_to_euclid(nears, fars))
    positions = positions_of(rays_o, rays_d, samples)
    density, geo = model.field.get_density(positions)
    dirs = rays_d[:, None, :].expand(-1, positions.shape[1], -1)
    rgb = model.field.get_rgb(dirs, geo)
    weights = get_weights(samples, density)
    comp = torch.sum(weights * rgb, dim=-2)
    return comp + rgb[..., -1, :] * (1.0 - torch.sum(weights, dim=-2))


def signerf_rgb_loss(pred: Tensor, target: Tensor, use_l1: bool = True) -> Tensor:
    """signerf/signerf.py:36-47: `self.rgb_loss(image, output)` with nerfstudio's L1Loss = nn.L1Loss / MSELoss = nn.MSELoss."""
    return torch.nn.functional.l1_loss(target, pred) if use_l1 else torch.nn.functional.mse_loss(target, pred)


def initial_samples_train(num_rays: int, num_samples: int, to_euclid, t_rand: Tensor) -> SamplesRef:
    """SpacedSampler.generate_ray_samples with train_stratified while training, single_jitter=True: every ray shifts all
    its bins by one draw t_rand [N,1] between the neighbouring bin centres."""
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, ...]
    bin_centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
    bin_upper = torch.cat([bin_centers, bins[..., -1:]], -1)
    bin_lower = torch.cat([bins[..., :1], bin_centers], -1)
    bins = bin_lower + (bin_upper - bin_lower) * t_rand
    euclid = to_euclid(bins)
    return SamplesRef(euclid[..., :-1, None], euclid[..., 1:, None], bins[..., :-1, None], bins[..., 1:, None])


def sdist_of(s: SamplesRef) -> Tensor:
    """losses.py ray_samples_to_sdist: the S+1 spacing-domain bin edges of a ray."""
    return torch.cat([s.spacing_starts[..., 0], s.spacing_ends[..., -1:, 0]], dim=-1)


def outer(t0_starts: Tensor, t0_ends: Tensor, t1_starts: Tensor, t1_ends: Tensor, y1: Tensor) -> Tensor:
    """losses.py outer: for every interval of t0, the mass of the t1 histogram that could fall inside it (upper bound)."""
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    cy1_lo = torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)
    cy1_hi = torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1)
    return cy1_hi - cy1_lo


def lossfun_outer(t: Tensor, w: Tensor, t_env: Tensor, w_env: Tensor) -> Tensor:
    """losses.py lossfun_outer (mip-NeRF 360 proposal loss): penalise the proposal histogram where it under-covers w."""
    eps = torch.finfo(torch.float32).eps
    w_outer = outer(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + eps)


def interlevel_loss(weights_list: List[Tensor], samples_list: List[SamplesRef]) -> Tensor:
    """losses.py interlevel_loss: the final level's (detached) histogram against every proposal level's."""
    c = sdist_of(samples_list[-1]).detach()
    w = weights_list[-1][..., 0].detach()
    loss = 0.0
    for samples, weights in zip(samples_list[:-1], weights_list[:-1]):
        loss = loss + torch.mean(lossfun_outer(c, w, sdist_of(samples), weights[..., 0]))
    return loss


def lossfun_distortion(t: Tensor, w: Tensor) -> Tensor:
    """losses.py lossfun_distortion (mip-NeRF 360 eq. 15)."""
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return loss_inter + loss_intra


def distortion_loss(weights_list: List[Tensor], samples_list: List[SamplesRef]) -> Tensor:
    """losses.py distortion_loss: on the final level (metrics_dict["distortion"] of NerfactoModel while training)."""
    return torch.mean(lossfun_distortion(sdist_of(samples_list[-1]), weights_list[-1][..., 0]))


def samples_from_edges(spacing: Tensor, euclid: Tensor) -> SamplesRef:
    """[N,S+1] bin edges (spacing domain, metres) -> SamplesRef."""
    return SamplesRef(euclid[..., :-1, None], euclid[..., 1:, None], spacing[..., :-1, None], spacing[..., 1:, None])


def forward_train(model: NerfactoRef, rays_o: Tensor, rays_d: Tensor, jitter: Optional[Tensor] = None,
                  camera_indices: Optional[Tensor] = None, update_proposals: bool = True,
                  fixed_samples: Optional[List[SamplesRef]] = None, anneal: float = 1.0) -> Dict[str, object]:
    """NerfactoModel.get_outputs WHILE TRAINING (models/nerfacto.py get_outputs + ProposalNetworkSampler.generate_ray_samples
    with _anneal = 1, i.e. past proposal_weights_anneal_max_num_iters, as a fine-tune of a trained scene is): stratified
    initial bins, two proposal levels with PDF re-sampling, the main field on the last level, autograd on.
    jitter [3, N] = the three per-ray draws (initial sampler, PDF level 1, PDF level 2); None = bin centres (eval bins).
    camera_indices [N]: per-ray rows of the appearance embedding (training semantics); None = the mean (eval semantics).
    update_proposals False = a step between two proposal updates (`proposal_update_every`): densities under no_grad.
    anneal: ProposalNetworkSampler._anneal, `annealed_weights = torch.pow(weights, self._anneal)` in front of each PDF sampler.
    fixed_samples: the three levels' bins handed in instead of sampled (gradient checks on identical sample positions:
    the samplers carry no gradient, but a 1e-6 shift of a bin edge moves a sample across cells of the finest hash levels)."""
    n = rays_o.shape[0]
    nears = torch.ones_like(rays_o[..., 0:1]) * model.near
    fars = torch.ones_like(rays_o[..., 0:1]) * model.far
    to_euclid = make_to_euclid(nears, fars)
    counts = list(model.num_proposal_samples) + [model.num_nerf_samples]
    jit = (lambda i: None) if jitter is None else (lambda i: jitter[i].reshape(n, 1))
    if fixed_samples is not None:
        samples = fixed_samples[0]
    else:
        samples = initial_samples(n, counts[0], to_euclid) if jitter is None else initial_samples_train(n, counts[0], to_euclid, jit(0))
    weights_list, samples_list = [], []
    for i, net in enumerate(model.proposal_networks):
        positions = positions_of(rays_o, rays_d, samples)
        if update_proposals:
            density = net.density(positions)
        else:
            with torch.no_grad():
                density = net.density(positions)
        w = get_weights(samples, density)
        weights_list.append(w)
        samples_list.append(samples)
        if fixed_samples is not None:
            samples = fixed_samples[i + 1]
        else:
            samples = pdf_resample(samples, torch.pow(w.detach(), anneal), counts[i + 1], to_euclid, jitter=jit(i + 1))
    positions = positions_of(rays_o, rays_d, samples)
    normals = pred_normals = None
    if getattr(model.field, "use_pred_normals", False):
        density, geo, normals, pred_normals = normals_and_pred_normals(model.field, positions)
    else:
        density, geo = model.field.get_density(positions)
    dirs = rays_d[:, None, :].expand(-1, positions.shape[1], -1)
    if camera_indices is None:
        rgb = model.field.get_rgb(dirs, geo)
    else:
        d = sh_components_deg4(((dirs + 1.0) / 2.0).reshape(-1, 3))
        app = model.field.embedding_appearance(camera_indices)[:, None, :].expand(-1, positions.shape[1], -1)
        h = torch.cat([d, geo.reshape(-1, model.field.geo_feat_dim), app.reshape(-1, model.field.appearance_embedding_dim)], dim=-1)
        rgb = model.field.mlp_head(h).view(*dirs.shape[:-1], -1)
    weights = get_weights(samples, density)
    weights_list.append(weights)
    samples_list.append(samples)
    comp = torch.sum(weights * rgb, dim=-2)
    rgb_out = comp + rgb[..., -1, :] * (1.0 - torch.sum(weights, dim=-2))
    out = {"rgb": rgb_out, "weights_list": weights_list, "samples_list": samples_list}
    if normals is not None:   # models/nerfacto.py get_outputs, `if self.training and self.config.predict_normals`
        out["rendered_orientation_loss"] = orientation_loss(weights.detach(), normals, rays_d)
        out["rendered_pred_normal_loss"] = pred_normal_loss(weights.detach(), normals.detach(), pred_normals)
        out["normals"], out["pred_normals"] = normals, pred_normals
    return out


def signerf_loss_dict(out: Dict[str, object], target: Tensor, use_l1: bool = True, interlevel_loss_mult: float = 1.0,
                      distortion_loss_mult: float = 0.002, orientation_loss_mult: float = 0.0001,
                      pred_normal_loss_mult: float = 0.001) -> Dict[str, Tensor]:
    """SIGNeRFModel.get_loss_dict while training, without the LPIPS term (signerf/signerf.py:41-82; the multipliers are
    NerfactoModelConfig's defaults, which signerf_config.py leaves untouched)."""
    ld = {"rgb_loss": signerf_rgb_loss(out["rgb"], target, use_l1),
          "interlevel_loss": interlevel_loss_mult * interlevel_loss(out["weights_list"], out["samples_list"]),
          "distortion_loss": distortion_loss_mult * distortion_loss(out["weights_list"], out["samples_list"])}
    if "rendered_orientation_loss" in out:      # signerf.py:69-80, NerfactoModelConfig's multipliers 1e-4 / 1e-3
        ld["orientation_loss"] = orientation_loss_mult * torch.mean(out["rendered_orientation_loss"])
        ld["pred_normal_loss"] = pred_normal_loss_mult * torch.mean(out["rendered_pred_normal_loss"])
    return ld


def patch_sample_method(batch_size: int, num_images: int, image_height: int, image_width: int, patch_size: int,
                        generator: Optional[torch.Generator] = None) -> Tensor:
    """signerf/data/signerf_patch_pixel_sampler.py:60-78 (the unmasked branch), with an explicit RNG."""
    sub_bs = batch_size // (patch_size ** 2)
    indices = torch.rand((sub_bs, 3), generator=generator) * torch.tensor(
        [num_images, image_height - patch_size, image_width - patch_size])
    indices = indices.view(sub_bs, 1, 1, 3).broadcast_to(sub_bs, patch_size, patch_size, 3).clone()
    yys, xxs = torch.meshgrid(torch.arange(patch_size), torch.arange(patch_size), indexing="ij")
    indices[:, ..., 1] += yys
    indices[:, ..., 2] += xxs
    return torch.floor(indices).long().flatten(0, 2)
