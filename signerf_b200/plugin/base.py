"""Config base class + the camera view the hot path needs."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Tuple, Union

import torch
from torch import Tensor

try:  # the real base class when the plugin runs inside nerfstudio
    from nerfstudio.configs.base_config import InstantiateConfig  # type: ignore
except Exception:  # nerfstudio absent (build container): same `_target` / `setup()` contract
    @dataclass
    class InstantiateConfig:  # type: ignore
        _target: Any = None

        def setup(self, **kwargs) -> Any:
            return self._target(self, **kwargs)


@dataclass
class CameraBatch:
    """What `render_camera` reads from nerfstudio `Cameras` (datasetgenerator.py:267,281,691): c2w [N,3,4], pinhole
    intrinsics, image size.  Built from a real `Cameras` (same attribute names) or directly.  Every intrinsic is either
    one number shared by all N cameras or a per-camera [N] tensor (original datasets carry per-frame intrinsics and
    sizes: datasetgenerator.py:331-334, :354-356, :456-461); indexing keeps each camera's own values."""
    camera_to_worlds: Tensor
    fx: Union[float, Tensor]
    fy: Union[float, Tensor]
    cx: Union[float, Tensor]
    cy: Union[float, Tensor]
    width: Union[int, Tensor]
    height: Union[int, Tensor]

    def __post_init__(self) -> None:
        n = len(self)
        for name in ("fx", "fy", "cx", "cy", "width", "height"):
            v = getattr(self, name)
            if isinstance(v, Tensor):
                v = v.detach().reshape(-1).cpu()
                if v.numel() == 1:
                    v = v[0].item()
                elif v.numel() != n:
                    raise ValueError(f"{name} has {v.numel()} entries for {n} cameras")
                elif bool((v == v[0]).all()):
                    v = v[0].item()
            if not isinstance(v, Tensor):
                v = int(v) if name in ("width", "height") else float(v)
            setattr(self, name, v)

    def __len__(self) -> int:
        return self.camera_to_worlds.shape[0]

    def _at(self, name: str, i):
        v = getattr(self, name)
        return v[i] if isinstance(v, Tensor) else v

    def __getitem__(self, i) -> "CameraBatch":
        c = self.camera_to_worlds[i]
        single = c.dim() == 2
        pick = (lambda name: self._at(name, i).reshape(-1) if isinstance(getattr(self, name), Tensor) else getattr(self, name))
        return CameraBatch(c[None] if single else c, pick("fx"), pick("fy"), pick("cx"), pick("cy"), pick("width"), pick("height"))

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def to(self, device) -> "CameraBatch":
        return CameraBatch(self.camera_to_worlds.to(device), self.fx, self.fy, self.cx, self.cy, self.width, self.height)

    def intrinsics(self) -> Tensor:
        """[N,4] fp32 (fx, fy, cx, cy) per camera, on the host."""
        n = len(self)
        cols = [getattr(self, k).to(torch.float32) if isinstance(getattr(self, k), Tensor)
                else torch.full((n,), float(getattr(self, k)), dtype=torch.float32) for k in ("fx", "fy", "cx", "cy")]
        return torch.stack(cols, dim=1)

    def image_size(self) -> Tuple[int, int]:
        """(height, width) shared by all cameras of the batch; one launch renders one image size, so a batch with
        mixed sizes must be split first (`size_groups`)."""
        if isinstance(self.width, Tensor) or isinstance(self.height, Tensor):
            raise ValueError("cameras of different image sizes in one batch: render them per size group (CameraBatch.size_groups)")
        return int(self.height), int(self.width)

    def size_groups(self) -> List[Tuple[Tuple[int, int], List[int]]]:
        """[((height, width), camera indices)] in first-appearance order."""
        n = len(self)
        hs = self.height.tolist() if isinstance(self.height, Tensor) else [self.height] * n
        ws = self.width.tolist() if isinstance(self.width, Tensor) else [self.width] * n
        groups: dict = {}
        for i, hw in enumerate(zip(hs, ws)):
            groups.setdefault((int(hw[0]), int(hw[1])), []).append(i)
        return list(groups.items())


def as_camera_batch(cameras) -> CameraBatch:
    """Accept a CameraBatch or a nerfstudio `Cameras` (duck-typed on its attribute names: `camera_to_worlds` [..,3,4],
    `fx / fy / cx / cy / width / height` as [N,1] tensors or numbers)."""
    if isinstance(cameras, CameraBatch):
        return cameras
    c2w = cameras.camera_to_worlds
    c2w = c2w[None] if c2w.dim() == 2 else c2w
    return CameraBatch(c2w[..., :3, :4], cameras.fx, cameras.fy, cameras.cx, cameras.cy, cameras.width, cameras.height)


def c2w_intr(cam: CameraBatch, device) -> Tuple[Tensor, Tensor]:
    c2w = cam.camera_to_worlds[..., :3, :4].to(device=device, dtype=torch.float32).contiguous()
    return c2w, cam.intrinsics().to(device)
