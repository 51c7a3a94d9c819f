"""Mirror of reference signerf/renderer/renderer.py (proxy-mesh depth through pyrender / EGL): same config fields,
attributes, `setup()` / `render_camera()` signatures and error behaviour, with the depth coming from the CUDA z-buffer
rasteriser `sgn_rasterize_depth` instead of an OpenGL context — no trimesh / pyrender / EGL needed,
no GPU -> CPU -> GPU copy.  The colour image (only consumed by combine_shape_with_depth, datasetgenerator.py:794-807) is
what pyrender produces for this scene: ONE mesh lit by ambient light 1.0 and nothing else (renderer.py:129-130), i.e. the
material's flat base colour on covered pixels over pyrender's default white background.  `RendererConfig.color` is never
handed to pyrender by the reference (renderer.py:55 stores it, nothing reads it); the material is whatever trimesh finds
in the OBJ — for the shipped models/bunny.obj the `bunny1.mtl` it names does not exist, so pyrender's default material
for meshes without one applies ([EXT] pyrender `Mesh.from_trimesh`: baseColorFactor 0.3 -> 77/255; parity unpinned,
`Renderer.material_base_color` is the knob)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Tuple, Type

import numpy as np
import torch
from torch import Tensor

from .. import ops
from .base import InstantiateConfig, as_camera_batch, c2w_intr

NERFSTUDIO_BLENDER_SCALE_RATIO: float = 10.0
_CONVERT = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, -1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]])


@dataclass
class RendererConfig(InstantiateConfig):
    """renderer.py:23-39"""
    _target: Type = field(default_factory=lambda: Renderer)
    position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    rotation: List[float] = field(default_factory=lambda: [0, 0, 0])
    scale: List[float] = field(default_factory=lambda: [0.1, 0.1, 0.1])
    color: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 1.0])
    object_path: str = field(default_factory=lambda: "models/bunny.obj")


def load_obj(path) -> Tuple[np.ndarray, np.ndarray]:
    """Wavefront OBJ geometry: (vertices [Nv,3] float32, faces [Nf,3] int32); polygons are fan-triangulated, texture /
    normal indices ignored (depth only needs positions)."""
    verts, faces = [], []
    with open(path, "r") as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif p[0] == "f":
                idx = []
                for tok in p[1:]:
                    k = int(tok.split("/")[0])
                    idx.append(k - 1 if k > 0 else len(verts) + k)
                for j in range(1, len(idx) - 1):
                    faces.append((idx[0], idx[j], idx[j + 1]))
    return np.asarray(verts, np.float32).reshape(-1, 3), np.asarray(faces, np.int32).reshape(-1, 3)


class Renderer:
    """renderer.py:44-196.  Attributes are mutable, as the GUI writes them (interface.py:344-359); call setup() after
    changing them, as the reference does (interface.py:375-377, :424-426)."""

    def __init__(self, config: RendererConfig, device: str) -> None:
        self.config, self.device = config, device
        self.position, self.rotation, self.scale = config.position, config.rotation, config.scale
        self.color, self.object_path = config.color, config.object_path
        self.material_base_color = [0.3, 0.3, 0.3]   # [EXT] pyrender's default material (see the module docstring)
        self.scene = None          # (vertices, faces, model) once setup() succeeded
        self.mesh = None

    def set_mesh(self, vertices, faces) -> None:
        """Use an in-memory mesh instead of object_path (tests, procedural proxies); call setup() afterwards."""
        self.mesh = (torch.as_tensor(np.asarray(vertices, np.float32)).to(self.device),
                     torch.as_tensor(np.asarray(faces, np.int32)).to(self.device))

    def object_pose(self) -> np.ndarray:
        """renderer.py:80-131: convert @ [Rz Ry Rx diag(10 * scale) | position]."""
        rx, ry, rz = (math.radians(float(r)) for r in self.rotation)
        Rx = np.array([[1, 0, 0], [0, math.cos(rx), -math.sin(rx)], [0, math.sin(rx), math.cos(rx)]])
        Ry = np.array([[math.cos(ry), 0, math.sin(ry)], [0, 1, 0], [-math.sin(ry), 0, math.cos(ry)]])
        Rz = np.array([[math.cos(rz), -math.sin(rz), 0], [math.sin(rz), math.cos(rz), 0], [0, 0, 1]])
        RS = np.dot(np.dot(Rz, np.dot(Ry, Rx)), np.diag([float(s) * NERFSTUDIO_BLENDER_SCALE_RATIO for s in self.scale]))
        pose = np.zeros((4, 4))
        pose[0:3, 0:3] = RS
        pose[:, 3] = list(self.position) + [1]
        return _CONVERT @ pose

    def setup(self) -> None:
        """renderer.py:64-131: prints and returns (scene stays None) when the path is not an existing .obj file."""
        if self.mesh is None or getattr(self, "_mesh_path", None) is not None:
            object_path = Path(self.object_path)
            if object_path.suffix != ".obj":
                print(f"Path {object_path} is not an obj file")
                return
            if not object_path.exists():
                print(f"Path {object_path} does not exist")
                print("Be sure that the path exists on the server not client")
                return
            if getattr(self, "_mesh_path", None) != str(object_path):
                v, f = load_obj(object_path)
                self.set_mesh(v, f)
                self._mesh_path = str(object_path)
        self.scene = (self.mesh[0], self.mesh[1], self.object_pose())

    def render_depths(self, cameras) -> Tensor:
        """Proxy depth [V,H,W,1] fp32 on the device for a batch of cameras (0 = empty)."""
        if self.scene is None:
            raise AttributeError("Renderer.setup() has not loaded a mesh (reference: self.scene is None)")
        cam = as_camera_batch(cameras)
        c2w, intr = c2w_intr(cam, self.device)
        v, f, model = self.scene
        h, w = cam.image_size()
        return ops.rasterize_depth(v, f, model, c2w, intr, h, w, znear=0.0001, zfar=10.0)

    def render_colors(self, depths: Tensor) -> Tensor:
        """Colour images uint8 [V,H,W,3] that go with `render_depths`' output (flat ambient shading, white background)."""
        fg = tuple(int(math.floor(min(max(float(c), 0.0), 1.0) * 255.0 + 0.5)) for c in self.material_base_color[:3])
        return ops.shape_color_u8(depths, fg, (255, 255, 255))

    def render_camera(self, camera) -> Tuple[Tensor, Tensor]:
        """-> (color uint8 [H,W,3], depth fp32 [H,W,1]) on the device, as renderer.py:149-196."""
        depth = self.render_depths(camera)
        return self.render_colors(depth)[0], depth[0]
