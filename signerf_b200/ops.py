"""Torch-tensor front end of the C ABI: every function here is a thin argument marshaller around one
`sgn_*` entry point (include/signerf_b200.h).  Tensors supply device pointers and the current CUDA
stream; all arithmetic happens in the sm_100a kernels.  Nothing here falls back to torch math.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .field import NerfactoFieldB200

MLP_FP16_MMA = _lib.SGN_MLP_FP16_MMA
MLP_FP32 = _lib.SGN_MLP_FP32


def _stream(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t: Optional[Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _req(t: Tensor, dtype: torch.dtype, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (signerf_b200 has no CPU path)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def piecewise_bin_edges(num_samples: int, near: float, far: float) -> Tensor:
    """Euclidean bin edges [S+1] of nerfstudio's UniformLinDispPiecewiseSampler in eval, shared by every
    ray because NearFarCollider sets constant near/far.  Same torch expression as the reference stack
    (SpacedSampler.generate_ray_samples), evaluated on the host and handed to the kernel verbatim."""
    bins = torch.linspace(0.0, 1.0, num_samples + 1)
    nears = torch.tensor(float(near), dtype=torch.float32)
    fars = torch.tensor(float(far), dtype=torch.float32)

    def s(x):
        return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))

    def s_inv(x):
        return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))

    s_near, s_far = s(nears), s(fars)
    return s_inv(bins * s_far + (1 - bins) * s_near).contiguous()


@dataclass
class RenderOptions:
    """Sampling options of one render call (SgnRenderOpts)."""
    mode: str = "flat"              # "flat" | "cascade"
    num_samples: int = 48           # flat: S bins; cascade: samples into the main field
    num_prop_samples: Tuple[int, int] = (256, 96)
    near_plane: float = 0.05
    far_plane: float = 1000.0
    mlp_mode: int = MLP_FP16_MMA

    def to_c(self, keep: list) -> _lib.SgnRenderOpts:
        if self.mode not in ("flat", "cascade"):
            raise ValueError(f"unknown render mode {self.mode!r}")
        n0 = self.num_samples if self.mode == "flat" else self.num_prop_samples[0]
        bins = piecewise_bin_edges(n0, self.near_plane, self.far_plane).numpy()
        keep.append(bins)
        o = _lib.SgnRenderOpts()
        o.near_plane, o.far_plane = self.near_plane, self.far_plane
        o.mode = 0 if self.mode == "flat" else 1
        o.num_samples = self.num_samples
        o.num_prop_samples[0], o.num_prop_samples[1] = self.num_prop_samples
        o.mlp_mode = self.mlp_mode
        o.h_bins = bins.ctypes.data_as(C.POINTER(C.c_float))
        return o


def render_views(fld: NerfactoFieldB200, c2w: Tensor, intr: Tensor, height: int, width: int,
                 opts: RenderOptions, want_acc: bool = False):
    """V views -> (rgb [V,H,W,3], depth [V,H,W,1][, acc [V,H,W,1]]) on the field's device."""
    dev = fld.device
    c2w = _req(c2w[..., :3, :4], torch.float32, "c2w")
    intr = _req(intr, torch.float32, "intr")
    V = c2w.shape[0]
    if intr.shape != (V, 4):
        raise ValueError(f"intr must be [V,4], got {tuple(intr.shape)}")
    rgb = torch.empty((V, height, width, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((V, height, width, 1), dtype=torch.float32, device=dev)
    acc = torch.empty((V, height, width, 1), dtype=torch.float32, device=dev) if want_acc else None
    keep: list = []
    o = opts.to_c(keep)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_render_views(fld.handle, _ptr(c2w), _ptr(intr), V, height, width, C.byref(o),
                                                _ptr(rgb), _ptr(depth), _ptr(acc), _stream(dev)))
    return (rgb, depth, acc) if want_acc else (rgb, depth)


def render_rays(fld: NerfactoFieldB200, origins: Tensor, directions: Tensor, opts: RenderOptions, want_acc: bool = False):
    """An explicit ray bundle (nerfstudio RayBundle.origins / .directions of any leading shape [..., 3]) ->
    (rgb [...,3], depth [...,1][, acc [...,1]]): `get_outputs_for_camera_ray_bundle`'s contract (datasetgenerator.py:694)."""
    dev = fld.device
    lead = tuple(origins.shape[:-1])
    o = _req(origins.reshape(-1, 3), torch.float32, "origins")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    if o.shape != d.shape:
        raise ValueError(f"origins {tuple(origins.shape)} and directions {tuple(directions.shape)} differ")
    n = o.shape[0]
    rgb = torch.empty(lead + (3,), dtype=torch.float32, device=dev)
    depth = torch.empty(lead + (1,), dtype=torch.float32, device=dev)
    acc = torch.empty(lead + (1,), dtype=torch.float32, device=dev) if want_acc else None
    keep: list = []
    co = opts.to_c(keep)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_render_rays(fld.handle, _ptr(o), _ptr(d), n, C.byref(co), _ptr(rgb), _ptr(depth), _ptr(acc),
                                               _stream(dev)))
    return (rgb, depth, acc) if want_acc else (rgb, depth)


def render_views_host(fld: NerfactoFieldB200, c2w: np.ndarray, intr: np.ndarray, height: int, width: int,
                      opts: RenderOptions, out_rgb: np.ndarray, out_depth: np.ndarray) -> None:
    """End-to-end entry with HOST buffers (cameras in, images out): H2D + kernels + D2H inside the call."""
    c2w = np.ascontiguousarray(c2w, dtype=np.float32)
    intr = np.ascontiguousarray(intr, dtype=np.float32)
    V = c2w.shape[0]
    assert c2w.shape == (V, 3, 4) and intr.shape == (V, 4)
    assert out_rgb.dtype == np.float32 and out_rgb.shape == (V, height, width, 3) and out_rgb.flags.c_contiguous
    assert out_depth.dtype == np.float32 and out_depth.shape == (V, height, width, 1) and out_depth.flags.c_contiguous
    keep: list = []
    o = opts.to_c(keep)
    with torch.cuda.device(fld.device):
        _lib.check(_lib.load().sgn_render_views_host(fld.handle, c2w.ctypes.data, intr.ctypes.data, V, height, width,
                                                     C.byref(o), out_rgb.ctypes.data, out_depth.ctypes.data, None))


def generate_rays(c2w: Tensor, intr: Tensor, height: int, width: int):
    """nerfstudio Cameras.generate_rays: (origins, directions [V,H,W,3], pixel_area, directions_norm [V,H,W,1])."""
    c2w = _req(c2w[..., :3, :4], torch.float32, "c2w")
    intr = _req(intr, torch.float32, "intr")
    dev, V = c2w.device, c2w.shape[0]
    o = torch.empty((V, height, width, 3), dtype=torch.float32, device=dev)
    d = torch.empty_like(o)
    area = torch.empty((V, height, width, 1), dtype=torch.float32, device=dev)
    nrm = torch.empty_like(area)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_generate_rays(_ptr(c2w), _ptr(intr), V, height, width, _ptr(o), _ptr(d), _ptr(area),
                                                 _ptr(nrm), _stream(dev)))
    return o, d, area, nrm


def hash_encode(fld: NerfactoFieldB200, positions01: Tensor, which: int = 0, want_indices: bool = True):
    """HashEncoding.pytorch_fwd probe: ([N,L,8] int64 table rows, [N,2L] features)."""
    p = _req(positions01.reshape(-1, 3), torch.float32, "positions")
    L = fld.grid.num_levels if which == 0 else fld.prop_grids[which - 1].num_levels
    N = p.shape[0]
    idx = torch.empty((N, L, 8), dtype=torch.int64, device=p.device) if want_indices else None
    feat = torch.empty((N, 2 * L), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().sgn_hash_encode(fld.handle, which, _ptr(p), N, _ptr(idx), _ptr(feat), _stream(p.device)))
    return idx, feat


def field_eval(fld: NerfactoFieldB200, positions: Tensor, directions: Tensor, mlp_mode: int = MLP_FP16_MMA):
    """NerfactoField density + rgb for world-space samples: (density [N], rgb [N,3])."""
    p = _req(positions.reshape(-1, 3), torch.float32, "positions")
    d = _req(directions.reshape(-1, 3), torch.float32, "directions")
    N = p.shape[0]
    den = torch.empty((N,), dtype=torch.float32, device=p.device)
    rgb = torch.empty((N, 3), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().sgn_field_eval(fld.handle, _ptr(p), _ptr(d), N, mlp_mode, _ptr(den), _ptr(rgb),
                                              _stream(p.device)))
    return den, rgb


@dataclass
class MaskOptions:
    """DatasetGenerator masking knobs (reference datasetgenerator.py:52-70)."""
    aabb: Sequence[float] = (-0.1, -0.1, -0.1, 0.1, 0.1, 0.1)
    inverse_mask: bool = False
    mask_dilation: Optional[Tuple[int, int]] = (50, 50)
    additional_depth_radius: float = 0.1
    manual_depth: Optional[Tuple[float, float]] = None

    def to_c(self) -> _lib.SgnMaskOpts:
        o = _lib.SgnMaskOpts()
        for i, v in enumerate(self.aabb):
            o.aabb[i] = float(v)
        o.inverse_mask = int(self.inverse_mask)
        o.dilate_w, o.dilate_h = self.mask_dilation if self.mask_dilation is not None else (0, 0)
        o.depth_radius = self.additional_depth_radius
        o.use_manual_depth = int(self.manual_depth is not None)
        if self.manual_depth is not None:
            o.manual_min, o.manual_max = self.manual_depth
        return o


def mask_condition(c2w: Tensor, intr: Tensor, depth: Tensor, opts: MaskOptions):
    """render_camera's "aabb" masking branch for V views: (mask uint8 [V,H,W,1], cond fp32 [V,H,W,1], stats [V,4])."""
    depth = _req(depth, torch.float32, "depth")
    V, H, W = depth.shape[0], depth.shape[1], depth.shape[2]
    dev = depth.device
    c2w = _req(c2w[..., :3, :4], torch.float32, "c2w")
    intr = _req(intr, torch.float32, "intr")
    mask = torch.empty((V, H, W, 1), dtype=torch.uint8, device=dev)
    cond = torch.empty((V, H, W, 1), dtype=torch.float32, device=dev)
    stats = torch.empty((V, 4), dtype=torch.float32, device=dev)
    o = opts.to_c()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_mask_condition(_ptr(c2w), _ptr(intr), V, H, W, _ptr(depth), C.byref(o), _ptr(mask),
                                                  _ptr(cond), _ptr(stats), _stream(dev)))
    return mask, cond, stats


def mask_condition_combined(c2w: Tensor, intr: Tensor, depth: Tensor, shape_depth: Tensor, shape_color: Tensor, opts: MaskOptions):
    """The "aabb" branch with combine_shape_with_depth (datasetgenerator.py:794-807): `shape_depth` [V,H,W,1] fp32 and
    `shape_color` [V,H,W,3] uint8 are Renderer.render_camera's outputs; same returns as mask_condition."""
    depth = _req(depth, torch.float32, "depth")
    shape_depth = _req(shape_depth, torch.float32, "shape_depth")
    if not shape_color.is_cuda or shape_color.dtype != torch.uint8:
        raise TypeError("shape_color must be a CUDA uint8 tensor [V,H,W,3]")
    shape_color = shape_color.contiguous()
    V, H, W = depth.shape[0], depth.shape[1], depth.shape[2]
    if tuple(shape_depth.shape[:3]) != (V, H, W) or tuple(shape_color.shape) != (V, H, W, 3):
        raise ValueError("shape depth / colour do not match the depth images")
    dev = depth.device
    c2w = _req(c2w[..., :3, :4], torch.float32, "c2w")
    intr = _req(intr, torch.float32, "intr")
    mask = torch.empty((V, H, W, 1), dtype=torch.uint8, device=dev)
    cond = torch.empty((V, H, W, 1), dtype=torch.float32, device=dev)
    stats = torch.empty((V, 4), dtype=torch.float32, device=dev)
    o = opts.to_c()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_mask_condition_combined(_ptr(c2w), _ptr(intr), V, H, W, _ptr(depth), _ptr(shape_depth),
                                                           _ptr(shape_color), C.byref(o), _ptr(mask), _ptr(cond), _ptr(stats),
                                                           _stream(dev)))
    return mask, cond, stats


def mask_condition_shape(proxy_depth: Tensor, depth: Tensor, opts: MaskOptions):
    """render_camera's "shape" masking branch for V views (proxy-mesh depth vs NeRF depth): same outputs as mask_condition."""
    depth = _req(depth, torch.float32, "depth")
    proxy_depth = _req(proxy_depth, torch.float32, "proxy_depth")
    if proxy_depth.shape != depth.shape:
        raise ValueError(f"proxy depth {tuple(proxy_depth.shape)} does not match depth {tuple(depth.shape)}")
    V, H, W = depth.shape[0], depth.shape[1], depth.shape[2]
    dev = depth.device
    mask = torch.empty((V, H, W, 1), dtype=torch.uint8, device=dev)
    cond = torch.empty((V, H, W, 1), dtype=torch.float32, device=dev)
    stats = torch.empty((V, 4), dtype=torch.float32, device=dev)
    o = opts.to_c()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().sgn_mask_condition_shape(_ptr(proxy_depth), _ptr(depth), V, H, W, C.byref(o), _ptr(mask),
                                                        _ptr(cond), _ptr(stats), _stream(dev)))
    return mask, cond, stats


def rasterize_depth(vertices: Tensor, faces: Tensor, model, c2w: Tensor, intr: Tensor, H: int, W: int, znear: float = 1e-4,
                    zfar: float = 10.0, cull_back: bool = True) -> Tensor:
    """Proxy-mesh metric depth [V,H,W,1] (0 = empty) with pyrender / OpenGL conventions; `model`: 4x4 object pose in
    OpenGL axes (16 floats, row-major, host)."""
    vertices = _req(vertices, torch.float32, "vertices")
    if not faces.is_cuda or faces.dtype != torch.int32:
        raise TypeError("faces must be a CUDA int32 tensor [Nf,3]")
    faces = faces.contiguous()
    c2w = _req(c2w[..., :3, :4], torch.float32, "c2w")
    intr = _req(intr, torch.float32, "intr")
    V, Nv, Nf = c2w.shape[0], vertices.shape[0], faces.shape[0]
    dev = c2w.device
    lib = _lib.load()
    ws = torch.empty(max(16, int(lib.sgn_rasterize_ws_bytes(Nv, V, H, W))), dtype=torch.uint8, device=dev)
    depth = torch.empty((V, H, W, 1), dtype=torch.float32, device=dev)
    m = (C.c_double * 16)(*[float(x) for x in (model.flatten().tolist() if hasattr(model, "flatten") else model)])
    with torch.cuda.device(dev):
        _lib.check(lib.sgn_rasterize_depth(_ptr(vertices), _ptr(faces), Nv, Nf, m, _ptr(c2w), _ptr(intr), V, H, W,
                                           float(znear), float(zfar), int(cull_back), _ptr(ws), _ptr(depth), _stream(dev)))
    return depth


def shape_color_u8(depth: Tensor, fg_rgb: Tuple[int, int, int], bg_rgb: Tuple[int, int, int] = (255, 255, 255)) -> Tensor:
    """Flat-ambient colour image of the proxy mesh: depth [...,1] fp32 (0 = empty) -> uint8 [...,3]."""
    depth = _req(depth, torch.float32, "depth")
    out = torch.empty(tuple(depth.shape[:-1]) + (3,), dtype=torch.uint8, device=depth.device)
    fg = (C.c_uint8 * 3)(*[int(c) for c in fg_rgb])
    bg = (C.c_uint8 * 3)(*[int(c) for c in bg_rgb])
    with torch.cuda.device(depth.device):
        _lib.check(_lib.load().sgn_shape_color_u8(_ptr(depth), depth.numel(), fg, bg, _ptr(out), _stream(depth.device)))
    return out


def dilate_ellipse(mask: Tensor, ksize: Tuple[int, int]) -> Tensor:
    """cv2.dilate(mask, getStructuringElement(MORPH_ELLIPSE, ksize)) > 0 for [V,H,W] uint8 masks."""
    m = _req(mask, torch.uint8, "mask")
    V, H, W = m.shape[0], m.shape[1], m.shape[2]
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _lib.check(_lib.load().sgn_dilate_ellipse(_ptr(m), V, H, W, int(ksize[0]), int(ksize[1]), _ptr(out),
                                                  _stream(m.device)))
    return out


@dataclass
class SheetLayout:
    """Reference-sheet geometry (reference datasetgenerator.py:498-503, :531-534)."""
    rows: int
    cols: int
    tile_h: int
    tile_w: int
    border: int = 0

    @property
    def height(self) -> int:
        h = self.rows * self.tile_h + (self.rows - 1) * self.border
        return -(-h // 8) * 8

    @property
    def width(self) -> int:
        w = self.cols * self.tile_w + (self.cols - 1) * self.border
        return -(-w // 8) * 8


def sheet_paste(src: Tensor, sheet: Tensor, layout: SheetLayout, first_cell: int = 0,
                threshold: Optional[float] = None) -> None:
    """Bilinear-resize tiles [V,H,W,C] to the layout's tile size and paste them into `sheet` [Hs,Ws,C] in place."""
    if src.dtype == torch.bool:
        src = src.to(torch.uint8)
    u8 = src.dtype == torch.uint8
    src = _req(src, torch.uint8 if u8 else torch.float32, "src")
    if not (sheet.is_cuda and sheet.dtype == torch.float32 and sheet.is_contiguous()):
        raise ValueError("sheet must be a contiguous float32 CUDA tensor")
    V, H, W, Cc = src.shape
    if tuple(sheet.shape) != (layout.height, layout.width, Cc):
        raise ValueError(f"sheet shape {tuple(sheet.shape)} != {(layout.height, layout.width, Cc)}")
    with torch.cuda.device(sheet.device):
        _lib.check(_lib.load().sgn_sheet_paste(_ptr(src), int(u8), V, H, W, Cc, _ptr(sheet), layout.height, layout.width,
                                               layout.rows, layout.cols, layout.border, layout.tile_h, layout.tile_w,
                                               first_cell, -1.0 if threshold is None else float(threshold),
                                               _stream(sheet.device)))


def sheet_cut(sheet: Tensor, layout: SheetLayout, cell: int, height: int, width: int) -> Tensor:
    """Cut tile `cell` out of the sheet and bilinear-resize it to (height, width)."""
    sheet = _req(sheet, torch.float32, "sheet")
    Cc = sheet.shape[2]
    out = torch.empty((height, width, Cc), dtype=torch.float32, device=sheet.device)
    with torch.cuda.device(sheet.device):
        _lib.check(_lib.load().sgn_sheet_cut(_ptr(sheet), sheet.shape[0], sheet.shape[1], Cc, layout.rows, layout.cols,
                                             layout.border, layout.tile_h, layout.tile_w, cell, _ptr(out), height, width,
                                             _stream(sheet.device)))
    return out


def blend_masked(edited: Tensor, base: Tensor, mask: Tensor) -> Tensor:
    """edited*mask + base*(1-mask) with a [..., 1] mask broadcast over the channels."""
    e = _req(edited, torch.float32, "edited")
    b = _req(base, torch.float32, "base")
    m = _req(mask, torch.float32, "mask")
    Cc = e.shape[-1]
    npix = e.numel() // Cc
    if b.shape != e.shape or m.numel() != npix:
        raise ValueError("shape mismatch in blend_masked")
    out = torch.empty_like(e)
    with torch.cuda.device(e.device):
        _lib.check(_lib.load().sgn_blend_masked(_ptr(e), _ptr(b), _ptr(m), npix, Cc, _ptr(out), _stream(e.device)))
    return out


def quantize_u8(x: Tensor) -> Tensor:
    """tensor_to_image's `np.uint8(x * 255)` truncation, on device."""
    x = _req(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().sgn_quantize_u8(_ptr(x), x.numel(), _ptr(out), _stream(x.device)))
    return out
