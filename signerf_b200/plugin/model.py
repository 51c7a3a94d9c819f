"""The `graph` object DatasetGenerator.render_camera talks to (datasetgenerator.py:691-701): `render_aabb`, `eval()`,
`train()`, `device`, `get_outputs_for_camera_ray_bundle(bundle) -> {"rgb", "depth", ...}` — backed by the fused
sm_100a renderer instead of nerfstudio's chunked torch forward."""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import torch
from torch import Tensor

from .. import ops
from ..field import HashGridParams, LinearParams, NerfactoFieldB200
from ..synthetic import hash_scalings
from .base import CameraBatch, c2w_intr


class FusedNerfactoGraph:
    def __init__(self, field: NerfactoFieldB200, render_opts: Optional[ops.RenderOptions] = None):
        self.field = field
        self.render_opts = render_opts or ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96))
        self.render_aabb = None
        self.training = False

    @property
    def device(self):
        return self.field.device

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def render_cameras(self, cam: CameraBatch) -> Dict[str, Tensor]:
        """All views of `cam` in one launch: rgb [V,H,W,3], depth [V,H,W,1], accumulation [V,H,W,1]."""
        c2w, intr = c2w_intr(cam, self.device)
        rgb, depth, acc = ops.render_views(self.field, c2w, intr, cam.height, cam.width, self.render_opts, want_acc=True)
        return {"rgb": rgb, "depth": depth, "accumulation": acc}

    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle) -> Dict[str, Tensor]:
        """nerfstudio passes a RayBundle; the fused path re-derives the rays from the camera the bundle was generated
        from (`bundle.camera` attached by plugin.DatasetGenerator.render_camera), exactly as generate_rays does."""
        cam = getattr(camera_ray_bundle, "camera", camera_ray_bundle)
        out = self.render_cameras(cam)
        return {k: v[0] for k, v in out.items()}

    # nerfacto checkpoint -> field (names per SURVEY §8c; torch-fallback `implementation="torch"` checkpoints)
    @classmethod
    def from_state_dict(cls, sd: Mapping[str, Tensor], device="cuda", num_train_data: Optional[int] = None,
                        average_init_density: float = 0.01, render_opts: Optional[ops.RenderOptions] = None):
        def grid(prefix: str, levels: int, max_res: int) -> HashGridParams:
            table = sd[prefix + ".hash_table"].detach().float()
            log2 = (table.shape[0] // levels).bit_length() - 1
            return HashGridParams(table.to(device).contiguous(), hash_scalings(levels, 16, max_res), log2)

        def lin(prefix: str) -> LinearParams:
            return LinearParams(sd[prefix + ".weight"].detach().float().cpu(), sd[prefix + ".bias"].detach().float().cpu())

        g = grid("field.mlp_base_grid", 16, 2048)
        base = [lin("field.mlp_base_mlp.layers.0"), lin("field.mlp_base_mlp.layers.1")]
        head = [lin(f"field.mlp_head.layers.{i}") for i in range(3)]
        emb = sd.get("field.embedding_appearance.embedding.weight")
        if emb is None:  # SIGNeRFPipeline.load_state_dict drops it (signerf_pipeline.py:110-111): fresh N(0,1) rows
            gen = torch.Generator().manual_seed(0)
            emb = torch.randn(num_train_data or 30, 32, generator=gen)
        app = emb.detach().float().cpu().mean(dim=0)
        pg, pm = [], []
        for i, max_res in enumerate((128, 256)):
            key = f"proposal_networks.{i}.encoding.hash_table"
            if key not in sd:
                break
            pg.append(grid(f"proposal_networks.{i}.encoding", 5, max_res))
            pm.append([lin(f"proposal_networks.{i}.mlp_base.layers.0"), lin(f"proposal_networks.{i}.mlp_base.layers.1")])
        fld = NerfactoFieldB200(g, base, head, app, average_init_density, pg, pm)
        if render_opts is None and not pg:
            render_opts = ops.RenderOptions(mode="flat", num_samples=48)
        return cls(fld, render_opts)
