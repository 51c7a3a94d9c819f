"""Synthetic nerfacto parameters and camera rings for benchmarks / smoke runs: random hash tables
U(-1,1)*1e-3, nn.Linear default init, N(0,1) appearance embedding, piecewise scalings.  No datasets or
checkpoints are available offline, so this is what "random hash-grid + MLP" in BASELINE.json's configs means."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from .field import HashGridParams, LinearParams, NerfactoFieldB200


def hash_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """HashEncoding.scalings with nerfstudio's own expression (float32 power: 16 -> 2047 at level 15)."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth**levels)


def _linear(i: int, o: int, gen: torch.Generator) -> LinearParams:
    bound = 1.0 / math.sqrt(i)  # kaiming_uniform(a=sqrt(5)) on weight and the default bias init share this bound
    w = (torch.rand(o, i, generator=gen) * 2 - 1) * bound
    b = (torch.rand(o, generator=gen) * 2 - 1) * bound
    return LinearParams(w, b)


def _grid(levels: int, max_res: int, log2: int, gen: torch.Generator, device, scale: float) -> HashGridParams:
    table = ((torch.rand(levels * (1 << log2), 2, generator=gen) * 2 - 1) * scale).to(device)
    return HashGridParams(table.contiguous(), hash_scalings(levels, 16, max_res), log2)


def random_field(seed: int = 0, device="cuda", dense: bool = False, with_proposals: bool = True,
                 table_scale: float = 1e-3, average_init_density: float = 0.01) -> NerfactoFieldB200:
    gen = torch.Generator().manual_seed(seed)
    grid = _grid(16, 2048, 19, gen, device, table_scale)
    base = [_linear(32, 64, gen), _linear(64, 16, gen)]
    head = [_linear(63, 64, gen), _linear(64, 64, gen), _linear(64, 3, gen)]
    app = torch.randn(30, 32, generator=gen).mean(dim=0)
    pg, pm = [], []
    if with_proposals:
        for max_res in (128, 256):
            pg.append(_grid(5, max_res, 17, gen, device, table_scale))
            pm.append([_linear(10, 16, gen), _linear(16, 1, gen)])
    if dense:  # sigma ~ 0.01 * e^6 ~ 4: compositing saturates inside the scene instead of at the far plane
        base[1].bias[0] = 6.0
        for m in pm:
            m[1].bias[0] = 6.0
    return NerfactoFieldB200(grid, base, head, app, average_init_density, pg, pm)


def camera_ring(n: int, width: int, height: int, radius: float = 0.5, theta_deg: float = 90.0,
                phi_deg: Tuple[float, float] = (0.0, 300.0)) -> Tuple[Tensor, Tensor]:
    """Look-at ring equivalent to the GUI default `circle_poses(size=n, radius=0.5, theta=90, phi=(0,300))`
    (reference signerf/utils/poses_generation.py:22-73, interface.py:62-71); fx = fy = W, principal point centred."""
    phis = torch.linspace(math.radians(phi_deg[0]), math.radians(phi_deg[1]), n)
    theta = torch.tensor(math.radians(theta_deg))
    pos = torch.stack([radius * torch.sin(theta) * torch.cos(phis), radius * torch.sin(theta) * torch.sin(phis),
                       radius * torch.cos(theta) * torch.ones_like(phis)], -1)
    z = pos / pos.norm(dim=-1, keepdim=True).clamp_min(1e-10)
    up = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    x = torch.linalg.cross(up, z)
    x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-10)
    y = torch.linalg.cross(z, x)
    c2w = torch.zeros(n, 3, 4)
    c2w[:, :, 0], c2w[:, :, 1], c2w[:, :, 2], c2w[:, :, 3] = x, y, z, pos
    intr = torch.tensor([[float(width), float(width), width / 2.0, height / 2.0]]).repeat(n, 1)
    return c2w, intr


def proxy_mesh(num_faces: int = 4968, bump: float = 0.15) -> Tuple[np.ndarray, np.ndarray]:
    """A closed, bumpy unit-sphere-like proxy with about `num_faces` triangles (vertices [Nv,3] fp32, faces [Nf,3] int32,
    outward winding) standing in for the reference's models/bunny.obj (4 968 faces) where that file is not available."""
    n_lon = max(8, int(round(math.sqrt(num_faces / 2.0 * 54.0 / 46.0))))
    n_lat = max(3, int(round(num_faces / (2.0 * n_lon))) + 1)
    verts = [(0.0, 0.0, 1.0)]
    for i in range(1, n_lat):
        th = math.pi * i / n_lat
        for j in range(n_lon):
            ph = 2.0 * math.pi * j / n_lon
            r = 1.0 + bump * math.sin(3.0 * th) * math.cos(4.0 * ph)
            verts.append((r * math.sin(th) * math.cos(ph), r * math.sin(th) * math.sin(ph), r * math.cos(th)))
    verts.append((0.0, 0.0, -1.0))
    south = len(verts) - 1
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)  # noqa: E731
    faces = []
    for j in range(n_lon):
        faces.append((0, ring(1, j), ring(1, j + 1)))
        faces.append((south, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j), ring(i + 1, j + 1)
            faces.append((a, c, d))
            faces.append((a, d, b))
    return np.asarray(verts, np.float32), np.asarray(faces, np.int32)


def random_pred_normals(seed: int = 5) -> dict:
    """nn.Linear-style random init of nerfacto's normal-prediction branch (`predict_normals=True`, signerf_config.py:33)
    under nerfstudio's parameter names: MLP 27 -> 64 -> 64 -> 64 and PredNormalsFieldHead's Linear(64, 3)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o: int, i: int):
        return (torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5, (torch.rand(o, generator=g) * 2 - 1) / i ** 0.5

    p = {}
    for j, (o_, i_) in enumerate(((64, 27), (64, 64), (64, 64))):
        p[f"field.mlp_pred_normals.layers.{j}.weight"], p[f"field.mlp_pred_normals.layers.{j}.bias"] = lin(o_, i_)
    p["field.field_head_pred_normals.net.weight"], p["field.field_head_pred_normals.net.bias"] = lin(3, 64)
    return p
