"""Test-side glue: turn the ORACLE's torch model into the product's field handle, shared metrics."""
import numpy as np
import torch

from signerf_b200.field import HashGridParams, LinearParams, NerfactoFieldB200


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _lin(layer):
    return LinearParams(layer.weight.detach().clone(), layer.bias.detach().clone())


def field_from_oracle(model, device="cuda", with_proposals=True) -> NerfactoFieldB200:
    f = model.field
    enc = f.encoding
    log2 = int(np.log2(enc.hash_table_size))
    grid = HashGridParams(enc.hash_table.detach().to(device).contiguous(), enc.scalings.detach().clone(), log2)
    pg, pm = [], []
    if with_proposals:
        for p in model.proposal_networks:
            pg.append(HashGridParams(p.encoding.hash_table.detach().to(device).contiguous(),
                                     p.encoding.scalings.detach().clone(), int(np.log2(p.encoding.hash_table_size))))
            pm.append([_lin(l) for l in p.mlp.layers])
    return NerfactoFieldB200(grid, [_lin(l) for l in f.mlp_base.layers], [_lin(l) for l in f.mlp_head.layers],
                             f.embedding_appearance.weight.detach().mean(dim=0), f.average_init_density, pg, pm)


def ring_cameras(n, width, height, radius=0.5):
    """Benchmark camera ring: circle_poses restated from the committed golden fixture when n == 16,
    otherwise an equivalent look-at ring built here (tests only)."""
    import math
    phis = torch.linspace(0.0, math.radians(300.0), n)
    theta = torch.tensor(math.radians(90.0))
    pos = torch.stack([radius * torch.sin(theta) * torch.cos(phis), radius * torch.sin(theta) * torch.sin(phis),
                       radius * torch.cos(theta) * torch.ones_like(phis)], -1)
    z = pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    x = torch.linalg.cross(up, z)
    x = x / x.norm(dim=-1, keepdim=True)
    y = torch.linalg.cross(z, x)
    c2w = torch.zeros(n, 3, 4)
    c2w[:, :, 0], c2w[:, :, 1], c2w[:, :, 2], c2w[:, :, 3] = x, y, z, pos
    intr = torch.tensor([[float(width), float(width), width / 2.0, height / 2.0]]).repeat(n, 1)
    return c2w, intr


def depth_agreement(depth: torch.Tensor, ref: torch.Tensor, same_bin_rtol: float = 1e-6):
    """Median depth is a discrete pick (the mid-point of the first bin whose cumulative weight reaches 0.5), so a
    1-ulp change of a weight can move a ray by a whole bin.  `same_bin_rtol` separates "same bin" from "another bin":
    1e-6 for the shared flat bins (identical edges), 1e-3 for PDF-resampled edges (they carry fp32 rounding).  Returns (fraction of rays that picked another bin,
    relative L2 over the rays that picked the same bin)."""
    d = depth.detach().double().cpu().flatten()
    r = ref.detach().double().cpu().flatten()
    moved = (d - r).abs() > same_bin_rtol * r.abs()
    same = ~moved
    err = float((d[same] - r[same]).norm() / r[same].norm().clamp_min(1e-30)) if bool(same.any()) else 0.0
    return float(moved.double().mean()), err
