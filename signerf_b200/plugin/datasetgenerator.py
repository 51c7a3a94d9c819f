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

    # ------------------------------------------------------------------ NeRF forward behind the reference's model contract
    def _fused_graph(self, graph):
        """The fused sm_100a renderer for `graph`.  A graph that offers `render_cameras` (plugin.FusedNerfactoGraph, or a
        SIGNeRFModel carrying the `signerf.signerf` shim's override) is used as it is.  Any other nerfstudio `Model`
        only promises the reference contract (`render_aabb`, `eval / train`, `device`,
        `get_outputs_for_camera_ray_bundle`, datasetgenerator.py:691-701): its nerfacto parameters are packed into a
        FusedNerfactoGraph once per weight version and that renderer is attached behind the call.  Returns None when
        the parameters are not torch-fallback nerfacto tensors (tcnn `params` blobs, other model families)."""
        if hasattr(graph, "render_cameras"):
            return graph
        if not hasattr(graph, "state_dict"):
            return None
        from .model import FusedNerfactoGraph
        sd = graph.state_dict()
        version = tuple((k, int(v.data_ptr()), int(getattr(v, "_version", 0))) for k, v in sd.items() if isinstance(v, Tensor))
        cached = getattr(self, "_fused_cache", None)
        if cached is not None and cached[0] is graph and cached[1] == version:
            return cached[2]
        cfg = getattr(graph, "config", None)
        try:
            fused = FusedNerfactoGraph.from_state_dict(
                sd, device=graph.device, num_train_data=getattr(graph, "num_train_data", None),
                average_init_density=float(getattr(cfg, "average_init_density", 0.01)), near_plane=float(getattr(cfg, "near_plane", 0.05)),
                far_plane=float(getattr(cfg, "far_plane", 1000.0)))
        except KeyError as e:
            if not getattr(self, "_warned_no_fused", False):
                print(f"[signerf_b200] fused renderer not attached ({e}); rendering through the model's own "
                      "get_outputs_for_camera_ray_bundle")
                self._warned_no_fused = True
            fused = None
        if cached is not None and cached[2] is not None and cached[2] is not fused:
            cached[2].field.close()
        self._fused_cache = (graph, version, fused)
        return fused

    def _nerf_outputs(self, graph, cameras, cam: CameraBatch) -> Dict[str, Tensor]:
        """{"rgb" [V,H,W,3], "depth" [V,H,W,1]} of V cameras: datasetgenerator.py:691-701 for all of them at once."""
        if getattr(graph, "render_aabb", None) is not None:
            raise NotImplementedError("graph.render_aabb (the viewer's crop box, datasetgenerator.py:691) is not supported by the "
                                      "fused renderer: disable the crop before generating")
        fused = self._fused_graph(graph)
        graph.eval()
        try:
            if fused is not None:
                out = fused.render_cameras(cam)
            else:   # the model's own forward, view by view, exactly as the reference calls it
                views = [cameras] if not hasattr(cameras, "__len__") or getattr(cameras, "camera_to_worlds").dim() == 2 else \
                    [cameras[i] for i in range(len(cam))]
                outs = []
                for c in views:
                    if not hasattr(c, "generate_rays"):
                        raise TypeError("this graph needs nerfstudio `Cameras` (camera.generate_rays) - plugin.CameraBatch only "
                                        "drives the fused renderer")
                    bundle = c.generate_rays(camera_indices=0, aabb_box=None)
                    o = graph.get_outputs_for_camera_ray_bundle(bundle)
                    if o is None:
                        raise RuntimeError("Render thread did not return any outputs")
                    outs.append(o)
                out = {k: torch.stack([o[k] for o in outs]) for k in ("rgb", "depth")}
        finally:
            graph.train()
        if out is None:
            raise RuntimeError("Render thread did not return any outputs")
        return out

    def render_views(self, graph, cameras, combine_shape_with_depth: Optional[bool] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """render_camera for V cameras at once (K1 + K2/K3, no host sync): rgb [V,H,W,3] fp32, mask [V,H,W,1] bool,
        condition [V,H,W,1] fp32."""
        combine = self.combine_shape_with_depth if combine_shape_with_depth is None else combine_shape_with_depth
        if self.masking_mode == "shape" and self.renderer is None:
            raise ValueError("Renderer is None but masking mode is shape")
        cam = as_camera_batch(cameras)
        out = self._nerf_outputs(graph, cameras, cam)
        rgb, depth = out["rgb"].to(torch.float32).contiguous(), out["depth"].to(torch.float32).contiguous()
        if self.masking_mode == "shape":      # datasetgenerator.py:711-757, proxy depth from the CUDA rasteriser
            mask, cond, _ = ops.mask_condition_shape(self.renderer.render_depths(cam), depth, self._mask_options())
            return rgb, mask.bool(), cond
        c2w, intr = c2w_intr(cam, depth.device)
        if combine:                           # datasetgenerator.py:794-807
            if self.renderer is None:
                raise ValueError("Renderer is None but masking mode is shape")
            shape_depth = self.renderer.render_depths(cam)
            mask, cond, _ = ops.mask_condition_combined(c2w, intr, depth, shape_depth, self.renderer.render_colors(shape_depth),
                                                        self._mask_options())
        else:
            mask, cond, _ = ops.mask_condition(c2w, intr, depth, self._mask_options())
        return rgb, mask.bool(), cond

    # ------------------------------------------------------------------ reference API
    def render_camera(self, graph, camera, with_mask: bool = True, with_condition: bool = True,
                      combine_shape_with_depth: bool = False):
        """datasetgenerator.py:677-820 incl. its early-return arity quirk (4-tuples when a part is skipped)."""
        cam = as_camera_batch(camera)
        if not with_mask:
            return self._nerf_outputs(graph, camera, cam)["rgb"][0], None, None, None
        rgb, mask, cond = self.render_views(graph, camera, combine_shape_with_depth=combine_shape_with_depth)
        if not with_condition:
            return rgb[0], mask[0], None, None
        return rgb[0], mask[0], cond[0]

    # ------------------------------------------------------------------ multi-GPU (one process per GPU, torch.distributed)
    @staticmethod
    def _dist() -> Tuple[int, int]:
        """(rank, world) of the default process group; (0, 1) when the caller did not set one up."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _render_reference_views(self, graph, cameras, cam: CameraBatch) -> Tuple[Tensor, Tensor, Tensor]:
        """The rows*cols-1 reference views (hot loop #2, datasetgenerator.py:517-519).  With `world` ranks the views are
        independent units: rank r renders views v = r (mod world) and ONE all-gather of the packed tiles
        (rgb | condition | mask, 5 floats per pixel) hands every rank the full set - bit-identical to rendering them all
        on one GPU (pure partition, no reduction)."""
        rank, world = self._dist()
        if world == 1 or self._fused_graph(graph) is None:
            return self.render_views(graph, cameras)
        from .. import sharding
        n = len(cam)
        mine = sharding.views_of_rank(n, world, rank)
        h, w = cam.image_size()
        if mine:
            rgb, mask, cond = self.render_views(graph, cam[mine])
            local = torch.cat([rgb, cond, mask.to(torch.float32)], dim=-1).contiguous()
        else:
            local = torch.empty((0, h, w, 5), dtype=torch.float32, device=graph.device)
        full = sharding.all_gather_views(local, n, world, rank)
        return full[..., 0:3].contiguous(), full[..., 4:5] > 0.5, full[..., 3:4].contiguous()

    def _diffuse_reference_sheet(self, image_sheet: Tensor, mask_sheet: Tensor, cond_sheet: Tensor, dev) -> Tensor:
        """The ONE reference-sheet diffusion (datasetgenerator.py:558).  It does not shard (attention spans all tiles):
        rank 0 runs it and broadcasts the edited sheet, so every rank continues from the same pixels whatever the
        backend (a remote or non-deterministic diffuser is called once, as in the reference)."""
        rank, world = self._dist()
        if world == 1:
            return self.diffuser.diffuse(image_sheet, image_sheet, mask_sheet, cond_sheet).to(dev)
        import torch.distributed as dist
        edited = (self.diffuser.diffuse(image_sheet, image_sheet, mask_sheet, cond_sheet).to(dev, torch.float32).contiguous()
                  if rank == 0 else torch.empty_like(image_sheet))
        dist.broadcast(edited, src=0)
        return edited

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
        rgb, mask, cond = self._render_reference_views(graph, cameras, cam)
        ops.sheet_paste(rgb, image_sheet, lay, 0)
        ops.sheet_paste(mask, mask_sheet, lay, 0, threshold=0.5)
        ops.sheet_paste(cond, cond_sheet, lay, 0)
        edited = self._diffuse_reference_sheet(image_sheet, mask_sheet, cond_sheet, dev)
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
            photo = Image.open(filename)
            if photo.mode == "RGBA":       # image_to_tensor (utils/image_tensor_converter.py:46-47)
                photo = photo.convert("RGB")
            render = torch.from_numpy(np.array(photo, dtype="float32") / 255.0).to(graph.device)
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

    # ------------------------------------------------------------------ dataset on disk
    def init_directory(self) -> None:
        """datasetgenerator.py:146-182"""
        from .dataset_io import DatasetWriter
        self._writer = DatasetWriter(self.config.path, self.dataset_name, self.downscale_factor)
        w = self._writer
        self.dataset_path, self.transforms_path, self.references_path = w.dataset_path, w.transforms_path, w.references_path
        self.images_path, self.masks_path, self.conditions_path = w.images_path, w.masks_path, w.conditions_path
        self.rendered_path, self.originals_path = w.rendered_path, w.originals_path
        w.init_directory(self.config)

    def save_generated_images(self, idx: int, images: Dict[str, Tensor], camera, current_transforms: dict,
                              is_original: bool = False) -> dict:
        """datasetgenerator.py:398-468 (PNG encoding on the writer's thread pool)."""
        cam = as_camera_batch(camera)
        return self._writer.save_generated_images(idx, images, cam.camera_to_worlds[0], cam.fx, cam.fy, cam.cx, cam.cy,
                                                  cam.width, cam.height, current_transforms, is_original)

    def generate_dataset(self, graph, reference_camera_to_worlds: Tensor, original_dataset=None,
                         synthetic_camera_to_worlds: Optional[Tensor] = None, merge_with_original_dataset: bool = False) -> None:
        """datasetgenerator.py:185-393: reference sheet, then one sheet per dataset camera (hot loop #1), optionally the
        original photos merged in with inverted masks; directory + transforms.json in the reference's schema.
        `original_dataset` is duck-typed on `.cameras` (camera_to_worlds / fx / fy / cx / cy / width / height, `.size`),
        `._dataparser_outputs.image_filenames` and `.get_image_float32(idx)`."""
        from .dataset_io import DatasetWriter
        if original_dataset is None and synthetic_camera_to_worlds is None:
            raise ValueError("Either original dataset or camera_to_worlds must be given")
        if merge_with_original_dataset and (original_dataset is None or synthetic_camera_to_worlds is None):
            raise ValueError("Original dataset and camera_to_worlds must be given to merge with original dataset")
        self.init_directory()
        self.renderer.setup()
        if synthetic_camera_to_worlds is not None:
            self.is_synthetic = True
        tw, th = int(self.width // self.downscale_factor), int(self.height // self.downscale_factor)
        dev = graph.device

        def batch(c2w: Tensor) -> CameraBatch:
            return CameraBatch(c2w.to(dev), self.fx, self.fy, self.cx, self.cy, int(self.width), int(self.height))

        reference_cameras = batch(reference_camera_to_worlds)
        cameras, original_filenames = None, None
        if original_dataset is not None:
            cameras = as_camera_batch(original_dataset.cameras)
            original_filenames = list(original_dataset._dataparser_outputs.image_filenames)  # noqa: SLF001
        if synthetic_camera_to_worlds is not None:
            cameras = batch(synthetic_camera_to_worlds)
            original_filenames = [None] * synthetic_camera_to_worlds.shape[0]
        rank, world = self._dist()
        transforms = DatasetWriter.new_transforms(self.is_synthetic, merge_with_original_dataset, self.original_transform_matrix,
                                                  self.original_scale_factor)
        image_sheet, mask_sheet, cond_sheet, edited_sheet, references = self.generate_reference_sheet(graph, reference_cameras, tw, th)
        n_ref, n_gen = len(reference_cameras), len(cameras)
        transforms["reference_indices"] = list(range(n_ref))
        transforms["generated_indices"] = list(range(n_ref, n_ref + n_gen))
        if rank == 0:
            self._writer.save_reference_sheets(image_sheet, mask_sheet, cond_sheet, edited_sheet)
            for i in range(n_ref):
                transforms = self.save_generated_images(i, references[i], reference_cameras[i], transforms)
            if world == 1:
                self._writer.write_transforms(transforms)
        # Hot loop #1 (datasetgenerator.py:331-338): every dataset camera owns its sheet (the edited reference tiles + its
        # own last tile) and its own denoising trajectory - independent units, camera i -> rank i mod world, no data-path
        # collective.  Each rank writes its own PNGs (file names carry the global index); the frames are gathered once at
        # the end and rank 0 writes transforms.json in index order.
        mine: Dict[str, list] = {"frames": []}
        produced: List[Tuple[int, dict]] = []

        def keep(idx: int) -> None:
            produced.append((idx, mine["frames"].pop()))

        for i in range(rank, n_gen, world):
            filename = original_filenames[i]
            images = self.generate_with_reference_sheet(graph, cameras[i], filename, tw, th, edited_sheet, cond_sheet)
            self.save_generated_images(n_ref + i, images, cameras[i], mine, filename is not None)
            keep(n_ref + i)
        if world == 1:
            transforms["frames"] += [f for _, f in produced]
            produced = []
            self._writer.write_transforms(transforms)
        if merge_with_original_dataset:
            ocams = as_camera_batch(original_dataset.cameras)
            base = n_ref + n_gen
            transforms["original_indices"] = list(range(base, base + len(ocams)))
            lay = ops.SheetLayout(1, 1, th, tw, 0)      # a 1 x 1 "sheet" = K4's bilinear resize (F.interpolate, :355-358)

            def scaled(t: Tensor, threshold: Optional[float] = None) -> Tensor:
                out = torch.zeros((lay.height, lay.width, t.shape[2]), dtype=torch.float32, device=dev)
                ops.sheet_paste(t.to(dev, torch.float32)[None].contiguous(), out, lay, 0, **({} if threshold is None else {"threshold": threshold}))
                return out[:th, :tw]

            for i in range(rank, len(ocams), world):
                image = original_dataset.get_image_float32(i).to(dev)
                render, mask, condition = self.render_camera(graph, ocams[i], combine_shape_with_depth=self.combine_shape_with_depth)
                mask = ~mask                           # the photos do not contain the object (:351-352)
                images = {"render": render, "mask": mask, "condition": condition, "edited": image,
                          "render_scaled": scaled(render), "mask_scaled": scaled(mask.float(), 0.5) > 0.5,
                          "condition_scaled": scaled(condition), "edited_scaled": scaled(image)}
                self.save_generated_images(base + i, images, ocams[i], mine, True)
                keep(base + i)
            if world == 1:
                transforms["frames"] += [f for _, f in produced]
                produced = []
        if world > 1:
            import torch.distributed as dist
            from .. import sharding
            self._writer.writer.flush()                 # this rank's PNGs are on disk before anybody names them
            parts: List[Optional[list]] = [None] * world
            dist.all_gather_object(parts, produced)
            if rank == 0:
                transforms["frames"] += sharding.merge_frames(parts)
        if rank == 0:
            self._writer.write_transforms(transforms)
        self._writer.writer.close()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
