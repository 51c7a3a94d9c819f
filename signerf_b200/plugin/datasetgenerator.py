"""Mirror of the hot-path half of reference signerf/datasetgenerator/datasetgenerator.py: DatasetGeneratorConfig
(:33-81), DatasetGenerator.__init__ (:87-143), render_camera (:677-820), generate_reference_sheet (:470-593) and
generate_with_reference_sheet (:597-674) — same names, arguments, return structure, error behaviour and quirks
(SURVEY Appendix A), with every per-view Python loop replaced by one batched launch of the sm_100a kernels.
Directory layout / PNG dumps / transforms.json (init_directory, save_generated_images, generate_dataset) are the
§8(f) row-3 "dataset writer" and stay with the reference."""
from __future__ import annotations

import datetime
import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Dict, List, Literal, Optional, Tuple, Type

import torch
from torch import Tensor

from .. import ops
from .base import CameraBatch, InstantiateConfig, as_camera_batch, c2w_intr
from .diffuser import Diffuser, DiffuserConfig
from .renderer import Renderer, RendererConfig


@dataclass
class DatasetGeneratorConfig(InstantiateConfig):
    """datasetgenerator.py:33-81"""
    _target: Type = field(default_factory=lambda: DatasetGenerator)
    path: Path = field(default_factory=lambda: Path("./generations"))
    dataset_name: str = field(default_factory=lambda: "experiment-" + datetime.datetime.now().strftime("%Y%m%d-%H%M%S"))
    downscale_factor: int = 2
    fx: Optional[float] = None
    fy: Optional[float] = None
    cx: Optional[float] = None
    cy: Optional[float] = None
    width: Optional[int] = None
    height: Optional[int] = None
    masking_mode: Literal["shape", "aabb"] = "aabb"
    aabb_min: List[float] = field(default_factory=lambda: [-0.1, -0.1, -0.1])
    aabb_max: List[float] = field(default_factory=lambda: [0.1, 0.1, 0.1])
    rows: int = 2
    cols: int = 3
    mask_dialation: Optional[Tuple[int, int]] = (50, 50)
    additional_depth_radius: float = 0.1
    renderer: RendererConfig = field(default_factory=RendererConfig)
    diffuser: DiffuserConfig = field(default_factory=DiffuserConfig)
    border_width_between_images: int = 0
    inverse_mask: bool = False
    manual_depth: Optional[Tuple[int, int]] = None
    combine_shape_with_depth: bool = False


class DatasetGenerator:
    def __init__(self, config: DatasetGeneratorConfig, original_transform_matrix: Tensor, original_scale_factor: float,
                 transform_poses_to_original_space: Callable[[Tensor], Tensor], device: str) -> None:
        self.config, self.device = config, device
        self.original_transform_matrix = original_transform_matrix
        self.original_scale_factor = original_scale_factor
        self.transform_poses_to_original_space = transform_poses_to_original_space
        self.path, self.dataset_name = config.path, config.dataset_name
        self.fx, self.fy, self.cx, self.cy = config.fx, config.fy, config.cx, config.cy
        self.width, self.height = config.width, config.height
        self.downscale_factor = config.downscale_factor
        self.masking_mode = config.masking_mode
        self.aabb = torch.tensor([config.aabb_min, config.aabb_max], dtype=torch.float32, device=self.device)
        self.inverse_mask = config.inverse_mask
        self.combine_shape_with_depth = config.combine_shape_with_depth
        self.rows, self.cols = config.rows, config.cols
        self.border_width_between_images = config.border_width_between_images
        self.mask_dialation = config.mask_dialation
        self.additional_depth_radius = config.additional_depth_radius
        self.manual_depth = config.manual_depth
        self.renderer = Renderer(config.renderer, device=self.device)
        self.diffuser = Diffuser(config.diffuser, device=self.device)
        self.is_synthetic = False

    # ------------------------------------------------------------------ helpers
    def _mask_options(self) -> ops.MaskOptions:
        return ops.MaskOptions(aabb=self.aabb.flatten().tolist(), inverse_mask=self.inverse_mask,
                               mask_dilation=tuple(self.mask_dialation) if self.mask_dialation is not None else None,
                               additional_depth_radius=self.additional_depth_radius,
                               manual_depth=tuple(self.manual_depth) if self.manual_depth is not None else None)

    def _layout(self, scaled_w: int, scaled_h: int) -> ops.SheetLayout:
        return ops.SheetLayout(self.rows, self.cols, scaled_h, scaled_w, self.border_width_between_images)

    def render_views(self, graph, cameras) -> Tuple[Tensor, Tensor, Tensor]:
        """render_camera for V cameras at once (K1 + K2/K3, no host sync): rgb [V,H,W,3] fp32, mask [V,H,W,1] bool,
        condition [V,H,W,1] fp32."""
        if self.combine_shape_with_depth:
            raise NotImplementedError("combine_shape_with_depth (datasetgenerator.py:794-807) conditions on pyrender's SHADED "
                                      "colour image, which the depth rasteriser does not produce")
        if self.masking_mode == "shape" and self.renderer is None:
            raise ValueError("Renderer is None but masking mode is shape")
        cam = as_camera_batch(cameras)
        graph.eval()
        out = graph.render_cameras(cam)
        graph.train()
        if out is None:
            raise RuntimeError("Render thread did not return any outputs")
        if self.masking_mode == "shape":      # datasetgenerator.py:711-757, proxy depth from the CUDA rasteriser
            mask, cond, _ = ops.mask_condition_shape(self.renderer.render_depths(cam), out["depth"], self._mask_options())
            return out["rgb"], mask.bool(), cond
        c2w, intr = c2w_intr(cam, graph.device)
        mask, cond, _ = ops.mask_condition(c2w, intr, out["depth"], self._mask_options())
        return out["rgb"], mask.bool(), cond

    # ------------------------------------------------------------------ reference API
    def render_camera(self, graph, camera, with_mask: bool = True, with_condition: bool = True,
                      combine_shape_with_depth: bool = False):
        """datasetgenerator.py:677-820 incl. its early-return arity quirk (4-tuples when a part is skipped)."""
        cam = as_camera_batch(camera)
        if not with_mask:
            graph.eval()
            out = graph.render_cameras(cam)
            graph.train()
            return out["rgb"][0], None, None, None
        if combine_shape_with_depth:
            raise NotImplementedError("combine_shape_with_depth needs pyrender's shaded colour image (datasetgenerator.py:794-807)")
        rgb, mask, cond = self.render_views(graph, cam)
        if not with_condition:
            return rgb[0], mask[0], None, None
        return rgb[0], mask[0], cond[0]

    def generate_reference_sheet(self, graph, cameras, scaled_image_width: int, scaled_image_height: int):
        """datasetgenerator.py:470-593 -> (image_sheet, mask_sheet, condition_sheet, edited_sheet, references)."""
        cam = as_camera_batch(cameras)
        n = len(cam)
        if n != (self.rows * self.cols) - 1:
            raise ValueError(f"Camera count {n} is not equal to (rows * cols) - 1 = {(self.rows * self.cols) - 1}")
        lay = self._layout(scaled_image_width, scaled_image_height)
        dev = graph.device
        image_sheet = torch.ones((lay.height, lay.width, 3), dtype=torch.float32, device=dev)
        mask_sheet = torch.zeros((lay.height, lay.width, 1), dtype=torch.float32, device=dev)
        cond_sheet = torch.zeros((lay.height, lay.width, 1), dtype=torch.float32, device=dev)
        rgb, mask, cond = self.render_views(graph, cam)
        ops.sheet_paste(rgb, image_sheet, lay, 0)
        ops.sheet_paste(mask, mask_sheet, lay, 0, threshold=0.5)
        ops.sheet_paste(cond, cond_sheet, lay, 0)
        edited = self.diffuser.diffuse(image_sheet, image_sheet, mask_sheet, cond_sheet).to(dev)
        edited_sheet = ops.blend_masked(edited, image_sheet, mask_sheet)
        references: List[Dict[str, Tensor]] = []
        th, tw, b = scaled_image_height, scaled_image_width, self.border_width_between_images
        for i in range(n):
            r0, c0 = (i // self.cols) * (th + b), (i % self.cols) * (tw + b)
            sl = (slice(r0, r0 + th), slice(c0, c0 + tw))
            references.append({
                "render": rgb[i], "mask": mask[i], "condition": cond[i],
                "render_scaled": image_sheet[sl], "mask_scaled": mask_sheet[sl] > 0.5, "condition_scaled": cond_sheet[sl],
                "edited": ops.sheet_cut(edited_sheet, lay, i, self.height, self.width), "edited_scaled": edited_sheet[sl],
            })
        return image_sheet, mask_sheet, cond_sheet, edited_sheet, references

    def generate_with_reference_sheet(self, graph, camera, filename, scaled_image_width: int, scaled_image_height: int,
                                      image_reference_sheet: Tensor, condition_reference_sheet: Tensor) -> Dict[str, Tensor]:
        """datasetgenerator.py:597-674.  Mutates both sheet arguments in place (last tile), like the reference."""
        render, mask, condition = self.render_camera(graph, camera, combine_shape_with_depth=self.combine_shape_with_depth)
        if filename is not None:   # original-dataset cameras use the photo for the last tile (:628-630)
            from PIL import Image
            import numpy as np
            render = torch.from_numpy(np.array(Image.open(filename), dtype="float32") / 255.0).to(graph.device)
        lay = self._layout(scaled_image_width, scaled_image_height)
        last = self.rows * self.cols - 1
        th, tw, b = scaled_image_height, scaled_image_width, self.border_width_between_images
        r0, c0 = (self.rows - 1) * (th + b), (self.cols - 1) * (tw + b)
        sl = (slice(r0, r0 + th), slice(c0, c0 + tw))
        mask_sheet = torch.zeros_like(condition_reference_sheet)
        ops.sheet_paste(render[None].contiguous(), image_reference_sheet, lay, last)
        ops.sheet_paste(mask[None].contiguous(), mask_sheet, lay, last, threshold=0.5)
        ops.sheet_paste(condition[None].contiguous(), condition_reference_sheet, lay, last)
        render_scaled, mask_scaled = image_reference_sheet[sl].clone(), mask_sheet[sl] > 0.5
        edited_sheet = self.diffuser.diffuse(image_reference_sheet, image_reference_sheet, mask_sheet,
                                             condition_reference_sheet).to(graph.device)
        edited_scaled = ops.blend_masked(edited_sheet[sl].contiguous(), render_scaled, mask_scaled.float())
        blended_sheet = edited_sheet.clone()
        blended_sheet[sl] = edited_scaled
        return {"render": render, "mask": mask, "condition": condition,
                "edited": ops.sheet_cut(blended_sheet, lay, last, self.height, self.width),
                "render_scaled": render_scaled, "mask_scaled": mask_scaled,
                "condition_scaled": condition_reference_sheet[sl], "edited_scaled": edited_scaled}
