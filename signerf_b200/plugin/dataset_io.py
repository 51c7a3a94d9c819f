"""On-disk formats either side of the hot path: the generated-dataset directory and
`transforms.json` the reference writes in `DatasetGenerator.init_directory` / `save_generated_images` / `generate_dataset`
(signerf/datasetgenerator/datasetgenerator.py:146-182, :286-295, :398-468) and reads back in `SIGNeRFDataParser`
(signerf/data/signerf_dataparser.py:99-146) and `load_previous_experiment_cameras` (signerf/utils/...:26-52).

Same directory names, file names, JSON keys, key order and indentation, so `ns-train nerfacto / splatfacto --data <dir>` and
the reference's own dataparser consume the result unchanged.  What differs is HOW images get to disk: the uint8
truncation of `tensor_to_image` runs on the GPU (`sgn_quantize_u8`), one pinned device -> host copy per image, and the
PNG encoding happens on a small thread pool off the critical path (the reference encodes 8 PNGs per view synchronously
between two diffusion calls); `flush()` joins the pool before every `transforms.json` write."""
from __future__ import annotations

import json
from concurrent.futures import Future, ThreadPoolExecutor
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch
from torch import Tensor

from .. import ops


def quantize_image(t: Tensor) -> np.ndarray:
    """tensor_to_image's pixels (signerf/utils/image_tensor_converter.py:7-33): uint8(t * 255) by truncation, [H,W] for one
    channel, [H,W,3] for three.  CUDA tensors are quantised on the device; bool masks count as 0 / 1."""
    if t.dim() != 3 or t.shape[2] not in (1, 3):
        raise AssertionError("Tensor must be of shape (H, W, 1) or (H, W, 3)")
    x = t.detach()
    if x.is_cuda:
        q = ops.quantize_u8(x.to(torch.float32).contiguous()).cpu().numpy()
    else:
        q = (x.to(torch.float32).numpy() * 255).astype(np.uint8)
    return q[..., 0] if q.shape[2] == 1 else q


class AsyncImageWriter:
    """PNG encoding off the critical path."""

    def __init__(self, workers: int = 4):
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="sgn-png")
        self._pending: List[Future] = []

    @staticmethod
    def _save(arr: np.ndarray, path: Path) -> None:
        from PIL import Image
        Image.fromarray(arr, "L" if arr.ndim == 2 else None).save(path)

    def save(self, tensor: Tensor, path: Path) -> None:
        self._pending.append(self._pool.submit(self._save, quantize_image(tensor), Path(path)))

    def flush(self) -> None:
        pending, self._pending = self._pending, []
        for f in pending:
            f.result()          # re-raises an encoder / filesystem error

    def close(self) -> None:
        self.flush()
        self._pool.shutdown()


class DatasetWriter:
    """Directory layout + transforms.json of one generated dataset (datasetgenerator.py:146-182, :286-295, :398-468)."""

    def __init__(self, root: Path, dataset_name: str, downscale_factor: int, workers: int = 4):
        ds = int(downscale_factor) if float(downscale_factor).is_integer() else downscale_factor
        self.dataset_path = Path(root) / dataset_name
        p = self.dataset_path
        self.images_path, self.masks_path, self.conditions_path = p / "images", p / "masks", p / "conditions"
        self.rendered_path, self.originals_path = p / "rendered", p / "originals"
        self.images_scaled_path, self.masks_scaled_path = p / f"images_{ds}", p / f"masks_{ds}"
        self.conditions_scaled_path, self.rendered_path_scaled = p / f"conditions_{ds}", p / f"rendered_{ds}"
        self.originals_scaled_path, self.references_path = p / f"originals_{ds}", p / "references"
        self.transforms_path = p / "transforms.json"
        self.writer = AsyncImageWriter(workers)

    def init_directory(self, config: Any = None) -> None:
        for d in (self.dataset_path, self.images_path, self.masks_path, self.conditions_path, self.rendered_path,
                  self.originals_path, self.images_scaled_path, self.masks_scaled_path, self.conditions_scaled_path,
                  self.rendered_path_scaled, self.originals_scaled_path, self.references_path):
            d.mkdir(parents=True, exist_ok=True)
        if config is not None:
            import yaml
            (self.dataset_path / "config.yml").write_text(yaml.dump(config), "utf8")

    @staticmethod
    def new_transforms(is_synthetic: bool, is_combined: bool, original_transform_matrix: Tensor,
                       original_scale_factor: float) -> Dict[str, Any]:
        """datasetgenerator.py:286-295 (key order is part of the format)."""
        return {
            "camera_model": "OPENCV",
            "orientation_override": "none",
            "method": "SIGNeRF",
            "is_synthetic": is_synthetic,
            "is_combined": is_combined,
            "frames": [],
            "original_transform_matrix": original_transform_matrix.cpu().numpy().tolist(),
            "original_scale_factor": original_scale_factor,
        }

    def save_reference_sheets(self, image: Tensor, mask: Tensor, condition: Tensor, edited: Tensor) -> None:
        """datasetgenerator.py:300-304"""
        for t, n in ((image, "image"), (mask, "mask"), (condition, "condition"), (edited, "edited")):
            self.writer.save(t, self.references_path / f"{n}_reference_sheet.png")

    def save_generated_images(self, idx: int, images: Dict[str, Tensor], c2w: Tensor, fx: float, fy: float, cx: float,
                              cy: float, width: int, height: int, current_transforms: Dict[str, Any],
                              is_original: bool = False) -> Dict[str, Any]:
        """datasetgenerator.py:398-468 (including that `render_scaled` goes to rendered_<ds>/ for originals as well, and that
        `transform_matrix` repeats the scene matrix — the reference's FIXME)."""
        w = self.writer
        if "edited" in images:
            w.save(images["edited"], self.images_path / f"image_{idx}.png")
        if "render" in images:
            w.save(images["render"], (self.originals_path if is_original else self.rendered_path) / f"image_{idx}.png")
        if "mask" in images:
            w.save(images["mask"], self.masks_path / f"mask_{idx}.png")
        if "condition" in images:
            w.save(images["condition"], self.conditions_path / f"condition_{idx}.png")
        if "edited_scaled" in images:
            w.save(images["edited_scaled"], self.images_scaled_path / f"image_{idx}.png")
        if "render_scaled" in images:
            w.save(images["render_scaled"], self.rendered_path_scaled / f"image_{idx}.png")
        if "mask_scaled" in images:
            w.save(images["mask_scaled"], self.masks_scaled_path / f"mask_{idx}.png")
        if "condition_scaled" in images:
            w.save(images["condition_scaled"], self.conditions_scaled_path / f"condition_{idx}.png")
        scene = torch.cat([c2w.detach().cpu().to(torch.float32)[:3, :4], torch.tensor([[0.0, 0.0, 0.0, 1.0]])], dim=0)
        current_transforms["frames"].append({
            "fl_x": float(fx), "fl_y": float(fy), "cx": float(cx), "cy": float(cy), "w": int(width), "h": int(height),
            "file_path": f"./images/image_{idx}.png",
            "_mask_path": f"./masks/mask_{idx}.png",
            "transform_matrix": scene.numpy().tolist(),
            "scene_transform_matrix": scene.numpy().tolist(),
        })
        return current_transforms

    def write_transforms(self, transforms: Dict[str, Any]) -> None:
        self.writer.flush()                     # every image the JSON names is on disk before the JSON is
        with open(self.transforms_path, "w") as fh:
            json.dump(transforms, fh, indent=4)


def read_transforms(path) -> Dict[str, Any]:
    """The subset `SIGNeRFDataParser._generate_dataparser_outputs` (signerf_dataparser.py:99-146, :210-223) and
    `load_previous_experiment_cameras` take from a generated transforms.json: per-frame intrinsics, the scene-space pose
    (`scene_transform_matrix`, falling back to `transform_matrix`), image / mask paths of the frames whose image exists,
    the index lists and the original-space transform."""
    path = Path(path)
    meta = json.loads((path / "transforms.json" if path.is_dir() else path).read_text())
    root = path if path.is_dir() else path.parent
    out: Dict[str, Any] = {"image_filenames": [], "mask_filenames": [], "poses": [], "fx": [], "fy": [], "cx": [], "cy": [],
                           "width": [], "height": [], "num_skipped": 0}
    for frame in meta["frames"]:
        fname = root / Path(frame["file_path"])
        if not fname.exists():
            out["num_skipped"] += 1
            continue
        for k, src in (("fx", "fl_x"), ("fy", "fl_y"), ("cx", "cx"), ("cy", "cy")):
            assert src in frame, f"{k} not specified in frame"
            out[k].append(float(frame[src]))
        out["height"].append(int(frame["h"]))
        out["width"].append(int(frame["w"]))
        out["image_filenames"].append(fname)
        out["poses"].append(np.array(frame["scene_transform_matrix"] if "scene_transform_matrix" in frame else frame["transform_matrix"]))
        if "_mask_path" in frame:
            out["mask_filenames"].append(root / Path(frame["_mask_path"]))
    out["poses"] = np.stack(out["poses"]).astype(np.float32) if out["poses"] else np.zeros((0, 4, 4), np.float32)
    for k in ("reference_indices", "generated_indices", "original_indices", "is_synthetic", "is_combined", "method",
              "camera_model", "original_scale_factor"):
        out[k] = meta.get(k)
    out["original_transform_matrix"] = (np.array(meta["original_transform_matrix"], np.float32)
                                        if "original_transform_matrix" in meta else None)
    return out


# ---------------------------------------------------------------------------------------------- DataparserOutputs contract
from dataclasses import dataclass, field as _field  # noqa: E402


@dataclass
class GeneratedDataparserOutputs:
    """The fields of nerfstudio's `DataparserOutputs` as `SIGNeRFDataParser` fills them for a GENERATED dataset
    (signerf/data/signerf_dataparser.py:301-312) — what `SIGNeRFPipeline` / `DatasetGenerator` consume
    (`.dataparser_transform`, `.dataparser_scale`, `.image_filenames`, `.cameras`; signerf_pipeline.py:53-55)."""
    image_filenames: List[Path]
    cameras: Any                               # plugin.base.CameraBatch-like: camera_to_worlds [N,3,4] + per-frame intrinsics
    scene_box: Tensor                          # aabb [2,3] = +-scene_scale
    mask_filenames: Optional[List[Path]]
    dataparser_scale: float
    dataparser_transform: Tensor               # [3,4]
    metadata: Dict[str, Any] = _field(default_factory=dict)


@dataclass
class FrameCameras:
    """Per-frame pinhole cameras of a generated dataset (the subset of nerfstudio `Cameras` the path reads)."""
    camera_to_worlds: Tensor                   # [N,3,4]
    fx: Tensor
    fy: Tensor
    cx: Tensor
    cy: Tensor
    width: Tensor
    height: Tensor

    def __len__(self) -> int:
        return self.camera_to_worlds.shape[0]


def parse_generated_dataset(data_dir, scene_scale: float = 1.0, downscale_factor: int = 1,
                            depth_unit_scale_factor: float = 1e-3) -> GeneratedDataparserOutputs:
    """`SIGNeRFDataParser._generate_dataparser_outputs` for a dataset written by `generate_dataset`
    (signerf_dataparser.py:57-313): poses from `scene_transform_matrix` as they are (the file carries
    `original_transform_matrix` / `original_scale_factor`, so no re-orientation or re-scaling: :210-228), masks only for
    merged datasets (`original_indices` present; frames outside it get an all-white `white.png`, :156-169); with
    downscale_factor > 1 images / masks resolve to `images_<ds>/<name>` / `masks_<ds>/<name>` (`_get_fname`, :330-357:
    existence check, returned paths and `white.png` all live there); cameras
    rescaled by 1 / downscale_factor (nerfstudio `rescale_output_resolution`: focal lengths and principal point scale,
    sizes floor), scene box +-scene_scale."""
    root = Path(data_dir)
    meta = json.loads((root / "transforms.json").read_text())
    original_indices = meta.get("original_indices")
    image_filenames: List[Path] = []
    mask_filenames: List[Path] = []
    poses, fx, fy, cx, cy, hh, ww = [], [], [], [], [], [], []
    ds = int(downscale_factor)

    def get_fname(filepath: Path, prefix: str) -> Path:
        """`SIGNeRFDataParser._get_fname` (signerf_dataparser.py:330-357) for an explicit downscale factor: the
        down-sampled copies live in `<prefix><ds>/<file name>` next to the full-resolution folder."""
        return root / f"{prefix}{ds}" / filepath.name if ds > 1 else root / filepath

    for idx, frame in enumerate(meta["frames"]):
        fname = get_fname(Path(frame["file_path"]), "images_")
        if not fname.exists():
            continue
        for lst, key in ((fx, "fl_x"), (fy, "fl_y"), (cx, "cx"), (cy, "cy")):
            assert key in frame, f"{key} not specified in frame"
            lst.append(float(frame[key]))
        hh.append(int(frame["h"]))
        ww.append(int(frame["w"]))
        image_filenames.append(fname)
        poses.append(np.array(frame["scene_transform_matrix"] if "scene_transform_matrix" in frame else frame["transform_matrix"]))
        if "_mask_path" in frame:
            mask_fname = get_fname(Path(frame["_mask_path"]), "masks_")
            if original_indices is not None and idx not in original_indices:
                white = mask_fname.parent / "white.png"
                if not white.exists():
                    from PIL import Image
                    Image.new("L", (ww[-1], hh[-1]), color=255).save(white)
                mask_filenames.append(white)
            else:
                mask_filenames.append(mask_fname)
    assert len(image_filenames) != 0, "No image files found."
    assert len(mask_filenames) in (0, len(image_filenames)), "Different number of image and mask filenames."
    poses_t = torch.from_numpy(np.array(poses).astype(np.float32))
    if "original_transform_matrix" not in meta or "original_scale_factor" not in meta:
        raise ValueError("not a SIGNeRF-generated dataset: original_transform_matrix / original_scale_factor missing "
                         "(auto-orientation of foreign datasets is nerfstudio's job)")
    if "original_indices" not in meta:
        mask_filenames = []
    s = 1.0 / float(downscale_factor)
    cams = FrameCameras(poses_t[:, :3, :4], torch.tensor(fx) * s, torch.tensor(fy) * s, torch.tensor(cx) * s, torch.tensor(cy) * s,
                        torch.floor(torch.tensor(ww, dtype=torch.float32) * s).to(torch.int32),
                        torch.floor(torch.tensor(hh, dtype=torch.float32) * s).to(torch.int32))
    a = float(scene_scale)
    return GeneratedDataparserOutputs(
        image_filenames=image_filenames, cameras=cams,
        scene_box=torch.tensor([[-a, -a, -a], [a, a, a]], dtype=torch.float32),
        mask_filenames=mask_filenames if len(mask_filenames) > 0 else None,
        dataparser_scale=float(meta["original_scale_factor"]),
        dataparser_transform=torch.tensor(meta["original_transform_matrix"], dtype=torch.float32),
        metadata={"depth_filenames": None, "depth_unit_scale_factor": depth_unit_scale_factor})
