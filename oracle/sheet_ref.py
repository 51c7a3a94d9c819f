"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the SIGNeRF-owned part of the hot
path: AABB mask, dilation, depth condition, sheet assembly, blends, uint8 quantisation.

PINNED: unlike nerfacto_ref.py, the pieces restated here live in /root/reference and are importable
without nerfstudio; tests/golden/make_golden.py runs the reference's own functions
(signerf/utils/intersection.py, image_tensor_converter.py, poses_generation.py) and cv2 4.13 and stores
their outputs in tests/golden/*.npz; tests/test_oracle_golden.py checks this file against those fixtures.
`render_camera_aabb` / `reference_sheet` re-type signerf/datasetgenerator/datasetgenerator.py:758-818 and
:498-539 (that module imports nerfstudio at the top and cannot be imported here).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


def intersect_with_aabb(rays_o: Tensor, rays_d: Tensor, aabb: Tensor) -> Tuple[Tensor, Tensor]:
    """signerf/utils/intersection.py:5-56 (slab test with the +1e-6 direction bias)."""
    shape = rays_o.shape
    o = rays_o.reshape(-1, 3)
    d = rays_d.reshape(-1, 3)
    dir_fraction = 1.0 / (d + 1e-6)
    t1 = (aabb[0, 0] - o[:, 0:1]) * dir_fraction[:, 0:1]
    t2 = (aabb[1, 0] - o[:, 0:1]) * dir_fraction[:, 0:1]
    t3 = (aabb[0, 1] - o[:, 1:2]) * dir_fraction[:, 1:2]
    t4 = (aabb[1, 1] - o[:, 1:2]) * dir_fraction[:, 1:2]
    t5 = (aabb[0, 2] - o[:, 2:3]) * dir_fraction[:, 2:3]
    t6 = (aabb[1, 2] - o[:, 2:3]) * dir_fraction[:, 2:3]
    nears = torch.max(torch.cat([torch.minimum(t1, t2), torch.minimum(t3, t4), torch.minimum(t5, t6)], dim=1), dim=1).values
    fars = torch.min(torch.cat([torch.maximum(t1, t2), torch.maximum(t3, t4), torch.maximum(t5, t6)], dim=1), dim=1).values
    return nears.reshape(shape[0], shape[1], 1), fars.reshape(shape[0], shape[1], 1)


def ellipse_kernel(ksize: Tuple[int, int]) -> np.ndarray:
    """cv2.getStructuringElement(MORPH_ELLIPSE, ksize) restated (OpenCV morph.dispatch.cpp)."""
    kw, kh = ksize
    r, c = kh // 2, kw // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    out = np.zeros((kh, kw), np.uint8)
    for i in range(kh):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * math.sqrt((r * r - dy * dy) * inv_r2)))
            out[i, max(c - dx, 0):min(c + dx + 1, kw)] = 1
    return out


def dilate(mask: np.ndarray, ksize: Tuple[int, int]) -> np.ndarray:
    """cv2.dilate(mask, ellipse) > 0 restated in numpy: anchor at (kw//2, kh//2), border ignored."""
    k = ellipse_kernel(ksize)
    kh, kw = k.shape
    ay, ax = kh // 2, kw // 2
    H, W = mask.shape
    m = mask > 0
    out = np.zeros((H, W), bool)
    for i in range(kh):
        js = np.nonzero(k[i])[0]
        if js.size == 0:
            continue
        j1, j2 = js[0], js[-1] + 1
        # dst(y, x) |= src(y + i - ay, x + j - ax) for j in [j1, j2)
        ys0, ys1 = max(0, -(i - ay)), min(H, H - (i - ay))
        if ys0 >= ys1:
            continue
        src_rows = m[ys0 + i - ay:ys1 + i - ay]
        csum = np.concatenate([np.zeros((src_rows.shape[0], 1), np.int32), np.cumsum(src_rows, axis=1, dtype=np.int32)], axis=1)
        xs = np.arange(W)
        lo = np.clip(xs + j1 - ax, 0, W)
        hi = np.clip(xs + j2 - ax, 0, W)
        hit = (csum[:, hi] - csum[:, lo]) > 0
        hit &= (hi > lo)[None, :]
        out[ys0:ys1] |= hit
    return out


def render_camera_aabb(rays_o: Tensor, rays_d: Tensor, depth: Tensor, aabb: Tensor, inverse_mask: bool = False,
                       mask_dilation: Optional[Tuple[int, int]] = (50, 50), additional_depth_radius: float = 0.1,
                       manual_depth: Optional[Tuple[float, float]] = None, use_cv2: bool = True):
    """datasetgenerator.py:758-818, masking_mode == "aabb", combine_shape_with_depth=False.
    rays_* [H,W,3], depth [H,W,1] -> (mask bool [H,W,1], cond float [H,W,1], stats dict)."""
    H, W = depth.shape[:2]
    nears, fars = intersect_with_aabb(rays_o, rays_d, aabb)
    non_empty_space = (nears < fars) & (nears > 0.0)
    visible_mask = (nears < depth) * (depth < fars) * non_empty_space
    visible_mask = ~visible_mask if inverse_mask else visible_mask
    is_visible = torch.sum(visible_mask) > 1e-6
    stats = {"is_visible": bool(is_visible), "count": int(visible_mask.sum())}
    if is_visible:
        if mask_dilation is not None:
            vis_np = visible_mask.cpu().numpy().astype(float)
            if use_cv2:
                import cv2
                vis_np = cv2.dilate(vis_np, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, mask_dilation))
            else:
                vis_np = dilate(vis_np[..., 0], mask_dilation).astype(float)
            mask_image = torch.tensor(vis_np, dtype=torch.float32).reshape(H, W, 1) > 0
        else:
            mask_image = visible_mask
        if manual_depth is not None:
            mn, mx = manual_depth
        else:
            masked = depth[(depth * visible_mask) > 0]
            mn = torch.min(masked[masked > 0]) - additional_depth_radius
            mx = torch.max(masked) + additional_depth_radius
        stats["min"], stats["max"] = float(mn), float(mx)
        depth_normalized = (depth - mn) / (mx - mn)
        condition_image = 1 - torch.clamp(depth_normalized, 0, 1)
    else:
        mask_image = torch.zeros(H, W, 1, dtype=torch.bool)
        condition_image = torch.zeros(H, W, 1, dtype=torch.float32)
    return mask_image, condition_image, stats


def _interp(x: Tensor, h: int, w: int) -> Tensor:
    return F.interpolate(x.permute(2, 0, 1).unsqueeze(0), (h, w), mode="bilinear", align_corners=False).squeeze(0).permute(1, 2, 0)


def sheet_size(rows: int, cols: int, th: int, tw: int, border: int) -> Tuple[int, int]:
    """datasetgenerator.py:498-503."""
    w = int(cols * tw) + int((cols - 1) * border)
    h = int(rows * th) + int((rows - 1) * border)
    return int(math.ceil(h / 8) * 8), int(math.ceil(w / 8) * 8)


def reference_sheet(renders: List[Tensor], masks: List[Tensor], conds: List[Tensor], rows: int, cols: int,
                    th: int, tw: int, border: int = 0):
    """datasetgenerator.py:506-539: white image sheet, zero mask/cond sheets, tiles pasted row-major."""
    Hs, Ws = sheet_size(rows, cols, th, tw, border)
    image = torch.ones((Hs, Ws, 3), dtype=torch.float32)
    mask = torch.zeros((Hs, Ws, 1), dtype=torch.float32)
    cond = torch.zeros((Hs, Ws, 1), dtype=torch.float32)
    for i, (r, m, c) in enumerate(zip(renders, masks, conds)):
        row, col = i // cols, i % cols
        r_s = _interp(r, th, tw)
        m_s = _interp(m.float(), th, tw) > 0.5
        c_s = _interp(c, th, tw)
        y0 = row * th + row * border
        x0 = col * tw + col * border
        image[y0:y0 + th, x0:x0 + tw, :] = r_s
        mask[y0:y0 + th, x0:x0 + tw, :] = m_s
        cond[y0:y0 + th, x0:x0 + tw, :] = c_s
    return image, mask, cond


def cut_tile(sheet: Tensor, cell: int, cols: int, th: int, tw: int, border: int, H: int, W: int) -> Tensor:
    """datasetgenerator.py:570-585: slice a tile, bilinear-resize to (H, W)."""
    row, col = cell // cols, cell % cols
    y0 = row * th + row * border
    x0 = col * tw + col * border
    return _interp(sheet[y0:y0 + th, x0:x0 + tw, :], H, W)


def blend(edited: Tensor, base: Tensor, mask: Tensor) -> Tensor:
    """datasetgenerator.py:562."""
    c = edited.shape[-1]
    return edited * mask.repeat(1, 1, c) + base * (1 - mask.repeat(1, 1, c))


def quantize_u8(x: Tensor) -> np.ndarray:
    """utils/image_tensor_converter.py:22-23,29-30: (x*255).astype(uint8) — truncation, no clamp."""
    a = x.detach().cpu().numpy()
    return (a * 255).astype(np.uint8)
