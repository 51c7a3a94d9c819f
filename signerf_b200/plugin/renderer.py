"""Mirror of reference signerf/renderer/renderer.py (proxy-mesh depth through pyrender / EGL).  Only the surface is
kept in this round: the mesh rasteriser is SURVEY §8(f) row 2 and is needed for masking_mode="shape" /
combine_shape_with_depth only (datasetgenerator.py:711-716, :794-798)."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Type

from .base import InstantiateConfig


@dataclass
class RendererConfig(InstantiateConfig):
    """renderer.py:23-39"""
    _target: Type = field(default_factory=lambda: Renderer)
    position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    rotation: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    scale: List[float] = field(default_factory=lambda: [1.0, 1.0, 1.0])
    color: List[float] = field(default_factory=lambda: [1.0, 1.0, 1.0, 1.0])
    object_path: Path = field(default_factory=lambda: Path("./models/bunny.obj"))


class Renderer:
    """renderer.py:44-196.  Attributes are mutable, as the GUI writes them (interface.py:344-359)."""

    def __init__(self, config: RendererConfig, device: str) -> None:
        self.config, self.device = config, device
        self.position, self.rotation, self.scale = config.position, config.rotation, config.scale
        self.color, self.object_path = config.color, config.object_path
        self.scene = None

    def setup(self) -> None:
        """renderer.py:64-131 prints and returns when the mesh cannot be loaded, leaving scene=None; same here."""
        self.scene = None

    def render_camera(self, camera):
        raise NotImplementedError("proxy-mesh depth rasteriser (reference renderer.py:149-196) is a §8(f) 'next' row; "
                                  "use masking_mode='aabb'")
