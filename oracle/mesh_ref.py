"""ORACLE (test infrastructure, NOT product code) — numpy restatement of the proxy-mesh depth the reference obtains from
pyrender / EGL (signerf/renderer/renderer.py:64-196) and of the masking_mode == "shape" branch of render_camera
(signerf/datasetgenerator/datasetgenerator.py:711-757).

parity unpinned: pyrender, trimesh and an EGL context are not installable here, and the result of a hardware
rasteriser is implementation-defined in its last bits anyway.  What is restated literally: the object pose
(renderer.py:80-116: Rz Ry Rx, scale x NERFSTUDIO_BLENDER_SCALE_RATIO = 10, translation), the Blender -> OpenGL axis swap
(:134-146) on both the object and the camera pose, pyrender 0.1.45's `IntrinsicsCamera.get_projection_matrix`
(znear 1e-4, zfar 10, :181), its depth read-back (`_read_main_framebuffer`: vertical flip, 2 d - 1,
2 n f / (f + n - d (f - n)), cleared pixels -> 0) and OpenGL's rasterisation rules (pixel centres at +0.5, 8 sub-pixel
bits, top-left rule, back-face culling for pyrender's single-sided default material, LESS on a 24-bit depth buffer).
Anchors: analytic depth of a fronto-parallel quad and of a sphere (tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np

NERFSTUDIO_BLENDER_SCALE_RATIO = 10.0
CONVERT = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, -1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
DEPTH_MAX = (1 << 24) - 1


def object_pose(position: Sequence[float], rotation_deg: Sequence[float], scale: Sequence[float]) -> np.ndarray:
    """renderer.py:80-131 -> 4x4 model matrix in OpenGL axes (convert @ [Rz Ry Rx diag(10 s) | position])."""
    rx, ry, rz = (math.radians(float(r)) for r in rotation_deg)
    Rx = np.array([[1, 0, 0], [0, math.cos(rx), -math.sin(rx)], [0, math.sin(rx), math.cos(rx)]])
    Ry = np.array([[math.cos(ry), 0, math.sin(ry)], [0, 1, 0], [-math.sin(ry), 0, math.cos(ry)]])
    Rz = np.array([[math.cos(rz), -math.sin(rz), 0], [math.sin(rz), math.cos(rz), 0], [0, 0, 1]])
    R = np.dot(Rz, np.dot(Ry, Rx))
    S = np.diag([float(s) * NERFSTUDIO_BLENDER_SCALE_RATIO for s in scale])
    pose = np.zeros((4, 4))
    pose[0:3, 0:3] = np.dot(R, S)
    pose[:, 3] = list(position) + [1]
    return CONVERT @ pose


def projection_matrix(fx, fy, cx, cy, width, height, znear=1e-4, zfar=10.0) -> np.ndarray:
    """pyrender camera.py IntrinsicsCamera.get_projection_matrix."""
    P = np.zeros((4, 4))
    P[0][0] = 2.0 * fx / width
    P[1][1] = 2.0 * fy / height
    P[0][2] = 1.0 - 2.0 * cx / width
    P[1][2] = 2.0 * cy / height - 1.0
    P[3][2] = -1.0
    P[2][2] = (zfar + znear) / (znear - zfar)
    P[2][3] = (2 * zfar * znear) / (znear - zfar)
    return P


def rasterize_depth(vertices: np.ndarray, faces: np.ndarray, model: np.ndarray, c2w: np.ndarray, intr: Sequence[float],
                    height: int, width: int, znear: float = 1e-4, zfar: float = 10.0, cull_back: bool = True) -> np.ndarray:
    """-> metric depth [H,W] float32 (0 = empty) of one view, as `Renderer.render_camera` returns it."""
    fx, fy, cx, cy = (float(v) for v in intr)
    cam = np.eye(4)
    cam[:3, :4] = np.asarray(c2w, dtype=np.float64)[:3, :4]
    cam = CONVERT @ cam
    R, t = cam[:3, :3], cam[:3, 3]
    vw = (model[:3, :3] @ vertices.astype(np.float64).T).T + model[:3, 3]
    eye = (vw - t) @ R                          # R^T (p - t)
    P = projection_matrix(fx, fy, cx, cy, width, height, znear, zfar)
    clip_x = P[0, 0] * eye[:, 0] + P[0, 2] * eye[:, 2]
    clip_y = P[1, 1] * eye[:, 1] + P[1, 2] * eye[:, 2]
    clip_z = P[2, 2] * eye[:, 2] + P[2, 3]
    clip_w = -eye[:, 2]
    ok = clip_w > 0
    w_safe = np.where(ok, clip_w, 1.0)
    X = np.rint((clip_x / w_safe + 1.0) * 0.5 * width * 256.0).astype(np.int64)
    Y = np.rint((clip_y / w_safe + 1.0) * 0.5 * height * 256.0).astype(np.int64)
    Z = (clip_z / w_safe + 1.0) * 0.5
    zbuf = np.full((height, width), DEPTH_MAX, dtype=np.int64)

    def edge(ax, ay, bx, by, px, py):
        return (bx - ax) * (py - ay) - (by - ay) * (px - ax)

    def owns(ax, ay, bx, by):
        dx, dy = bx - ax, by - ay
        return dy < 0 or (dy == 0 and dx < 0)

    for f in np.asarray(faces):
        i0, i1, i2 = (int(k) for k in f)
        if not (ok[i0] and ok[i1] and ok[i2]):
            continue
        area = edge(X[i0], Y[i0], X[i1], Y[i1], X[i2], Y[i2])
        if area == 0:
            continue
        if area < 0:
            if cull_back:
                continue
            i1, i2, area = i2, i1, -area
        xs, ys = (X[i0], X[i1], X[i2]), (Y[i0], Y[i1], Y[i2])
        x0, x1 = max(0, (min(xs) - 128 + 255) >> 8), min(width - 1, (max(xs) - 128) >> 8)
        y0, y1 = max(0, (min(ys) - 128 + 255) >> 8), min(height - 1, (max(ys) - 128) >> 8)
        if x1 < x0 or y1 < y0:
            continue
        px = (256 * np.arange(x0, x1 + 1, dtype=np.int64) + 128)[None, :]
        py = (256 * np.arange(y0, y1 + 1, dtype=np.int64) + 128)[:, None]
        e0 = edge(X[i1], Y[i1], X[i2], Y[i2], px, py)
        e1 = edge(X[i2], Y[i2], X[i0], Y[i0], px, py)
        e2 = edge(X[i0], Y[i0], X[i1], Y[i1], px, py)
        inside = ((e0 >= (0 if owns(X[i1], Y[i1], X[i2], Y[i2]) else 1)) & (e1 >= (0 if owns(X[i2], Y[i2], X[i0], Y[i0]) else 1))
                  & (e2 >= (0 if owns(X[i0], Y[i0], X[i1], Y[i1]) else 1)))
        zw = (e0.astype(np.float64) * Z[i0] + e1.astype(np.float64) * Z[i1] + e2.astype(np.float64) * Z[i2]) * (1.0 / float(area))
        inside &= (zw >= 0.0) & (zw <= 1.0)
        d = np.minimum(float(DEPTH_MAX), np.floor(zw * float(DEPTH_MAX) + 0.5)).astype(np.int64)
        rows = height - 1 - np.arange(y0, y1 + 1)
        sub = zbuf[rows[:, None], np.arange(x0, x1 + 1)[None, :]]
        zbuf[rows[:, None], np.arange(x0, x1 + 1)[None, :]] = np.where(inside & (d < sub), d, sub)
    # pyrender read-back (float32 numpy arithmetic)
    depth_im = (zbuf.astype(np.float32) / np.float32(DEPTH_MAX)).astype(np.float32)
    empty = zbuf == DEPTH_MAX
    depth_im = np.float32(2.0) * depth_im - np.float32(1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        metric = np.float32(2.0 * znear * zfar) / (np.float32(zfar + znear) - depth_im * np.float32(zfar - znear))
    metric = metric.astype(np.float32)
    metric[empty] = 0.0
    return metric


def shape_mask_condition(proxy: np.ndarray, nerf: np.ndarray, inverse_mask: bool = False,
                         dilation: Optional[Tuple[int, int]] = (50, 50), radius: float = 0.1,
                         manual_depth: Optional[Tuple[float, float]] = None):
    """datasetgenerator.py:711-757 for one view; proxy / nerf depth [H,W] float32 -> (mask bool [H,W], cond float32 [H,W],
    is_visible)."""
    from oracle import sheet_ref as S
    non_empty = proxy > 0
    visible = (proxy < nerf) & non_empty
    visible = ~visible if inverse_mask else visible
    if not visible.sum() > 1e-6:
        return np.zeros_like(visible), np.zeros(proxy.shape, np.float32), False
    mask = S.dilate(visible.astype(np.uint8), dilation) > 0 if dilation is not None else visible
    if manual_depth is not None:
        mn, mx = np.float32(manual_depth[0]), np.float32(manual_depth[1])
    else:
        mn = proxy[visible & (proxy > 0)].min() - np.float32(radius)
        mx = proxy.max() + np.float32(radius)
    on = (proxy - mn) / (mx - mn)
    nn = (nerf - mn) / (mx - mn)
    cond = visible.astype(np.float32) * on + (~visible).astype(np.float32) * nn
    return mask, (np.float32(1.0) - np.clip(cond, 0, 1)).astype(np.float32), True


def parse_obj(text: str):
    """Minimal Wavefront OBJ reader (v / f with v, v/vt, v//vn, v/vt/vn; polygons fan-triangulated; negative indices)."""
    verts, faces = [], []
    for line in text.splitlines():
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            verts.append([float(p[1]), float(p[2]), float(p[3])])
        elif p[0] == "f":
            idx = []
            for tok in p[1:]:
                k = int(tok.split("/")[0])
                idx.append(k - 1 if k > 0 else len(verts) + k)
            for j in range(1, len(idx) - 1):
                faces.append([idx[0], idx[j], idx[j + 1]])
    return np.asarray(verts, np.float32).reshape(-1, 3), np.asarray(faces, np.int32).reshape(-1, 3)


def uv_sphere(radius: float = 1.0, n_lat: int = 24, n_lon: int = 48):
    """Outward-facing (counter-clockwise) triangle sphere, for the analytic checks."""
    verts = [[0.0, 0.0, radius]]
    for i in range(1, n_lat):
        th = math.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * math.pi * j / n_lon
            verts.append([radius * math.sin(th) * math.cos(ph), radius * math.sin(th) * math.sin(ph), radius * math.cos(th)])
    verts.append([0.0, 0.0, -radius])
    faces = []
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)   # noqa: E731
    for j in range(n_lon):
        faces.append([0, ring(1, j), ring(1, j + 1)])
        faces.append([len(verts) - 1, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)])
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            faces.append([ring(i, j), ring(i + 1, j), ring(i + 1, j + 1)])
            faces.append([ring(i, j), ring(i + 1, j + 1), ring(i, j + 1)])
    return np.asarray(verts, np.float32), np.asarray(faces, np.int32)
