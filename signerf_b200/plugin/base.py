"""Config base class + the camera view the hot path needs."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Tuple

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
    intrinsics, image size.  Built from a real `Cameras` (same attribute names) or directly."""
    camera_to_worlds: Tensor
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int

    def __len__(self) -> int:
        return self.camera_to_worlds.shape[0]

    def __getitem__(self, i) -> "CameraBatch":
        c = self.camera_to_worlds[i]
        return CameraBatch(c[None] if c.dim() == 2 else c, self.fx, self.fy, self.cx, self.cy, self.width, self.height)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def to(self, device) -> "CameraBatch":
        return CameraBatch(self.camera_to_worlds.to(device), self.fx, self.fy, self.cx, self.cy, self.width, self.height)


def _scalar(v) -> float:
    return float(v.reshape(-1)[0]) if isinstance(v, Tensor) else float(v)


def as_camera_batch(cameras) -> CameraBatch:
    """Accept a CameraBatch or a nerfstudio `Cameras` (duck-typed on its attribute names)."""
    if isinstance(cameras, CameraBatch):
        return cameras
    c2w = cameras.camera_to_worlds
    c2w = c2w[None] if c2w.dim() == 2 else c2w
    return CameraBatch(c2w[..., :3, :4], _scalar(cameras.fx), _scalar(cameras.fy), _scalar(cameras.cx), _scalar(cameras.cy),
                       int(_scalar(cameras.width)), int(_scalar(cameras.height)))


def c2w_intr(cam: CameraBatch, device) -> Tuple[Tensor, Tensor]:
    c2w = cam.camera_to_worlds[..., :3, :4].to(device=device, dtype=torch.float32).contiguous()
    intr = torch.tensor([[cam.fx, cam.fy, cam.cx, cam.cy]], dtype=torch.float32, device=device).repeat(c2w.shape[0], 1)
    return c2w, intr
