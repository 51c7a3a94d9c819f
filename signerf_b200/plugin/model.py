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
        # nerfstudio is not installable here, so the parameter names cannot be checked against a real 1.0.2 checkpoint:
        # both layouts the torch-fallback modules have used are accepted (MLPWithHashEncoding: `mlp_base.encoding` /
        # `mlp_base.mlp`; older split modules: `mlp_base_grid` / `mlp_base_mlp`, `encoding` / `mlp_base`).
        def pick(*cands: str) -> str:
            for c in cands:
                if c + ".hash_table" in sd or c + ".layers.0.weight" in sd:
                    return c
            raise KeyError(f"none of {cands} found in the state_dict (tcnn checkpoints store a flat `params` vector and "
                           "are not supported: torch-fallback semantics are the contract, SURVEY §7)")

        def grid(prefix: str, levels: int, max_res: int) -> HashGridParams:
            table = sd[prefix + ".hash_table"].detach().float()
            log2 = (table.shape[0] // levels).bit_length() - 1
            return HashGridParams(table.to(device).contiguous(), hash_scalings(levels, 16, max_res), log2)

        def mlp(prefix: str, n: int):
            return [LinearParams(sd[f"{prefix}.layers.{i}.weight"].detach().float().cpu(),
                                 sd[f"{prefix}.layers.{i}.bias"].detach().float().cpu()) for i in range(n)]

        g = grid(pick("field.mlp_base.encoding", "field.mlp_base_grid"), 16, 2048)
        base = mlp(pick("field.mlp_base.mlp", "field.mlp_base_mlp"), 2)
        head = mlp(pick("field.mlp_head"), 3)
        emb = sd.get("field.embedding_appearance.embedding.weight")
        if emb is None:  # SIGNeRFPipeline.load_state_dict drops it (signerf_pipeline.py:110-111): fresh N(0,1) rows
            gen = torch.Generator().manual_seed(0)
            emb = torch.randn(num_train_data or 30, 32, generator=gen)
        app = emb.detach().float().cpu().mean(dim=0)
        pg, pm = [], []
        for i, max_res in enumerate((128, 256)):
            pre = f"proposal_networks.{i}"
            try:
                enc = pick(pre + ".mlp_base.encoding", pre + ".encoding")
            except KeyError:
                break   # load_model_with_proposal_weights=False drops the proposal networks (signerf_pipeline.py:113-118)
            pg.append(grid(enc, 5, max_res))
            pm.append(mlp(pick(pre + ".mlp_base.mlp", pre + ".mlp_base.1", pre + ".mlp_base"), 2))
        fld = NerfactoFieldB200(g, base, head, app, average_init_density, pg, pm)
        if render_opts is None and not pg:
            render_opts = ops.RenderOptions(mode="flat", num_samples=48)
        return cls(fld, render_opts)
