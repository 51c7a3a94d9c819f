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
from .base import CameraBatch, as_camera_batch, c2w_intr


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
        h, w = cam.image_size()
        rgb, depth, acc = ops.render_views(self.field, c2w, intr, h, w, self.render_opts, want_acc=True)
        return {"rgb": rgb, "depth": depth, "accumulation": acc}

    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle) -> Dict[str, Tensor]:
        """nerfstudio's `Model.get_outputs_for_camera_ray_bundle` (the call at datasetgenerator.py:694): a RayBundle with
        `origins` / `directions` [H,W,3] (what `camera.generate_rays(camera_indices=0)` returns) -> {"rgb" [H,W,3],
        "depth" [H,W,1], "accumulation" [H,W,1]} through `sgn_render_rays`.  A CameraBatch / Cameras passed instead of a
        bundle is rendered from its pose (rays generated inside the kernel)."""
        if hasattr(camera_ray_bundle, "origins") and hasattr(camera_ray_bundle, "directions"):
            if getattr(camera_ray_bundle, "nears", None) is not None or getattr(camera_ray_bundle, "fars", None) is not None:
                raise NotImplementedError("per-ray nears / fars (aabb_box colliders) are not supported: NearFarCollider planes only")
            rgb, depth, acc = ops.render_rays(self.field, camera_ray_bundle.origins.to(self.device),
                                              camera_ray_bundle.directions.to(self.device), self.render_opts, want_acc=True)
            return {"rgb": rgb, "depth": depth, "accumulation": acc}
        out = self.render_cameras(as_camera_batch(camera_ray_bundle))
        return {k: v[0] for k, v in out.items()}

    @classmethod
    def from_checkpoint(cls, path, device="cuda", **kw):
        """A nerfstudio checkpoint file (`step-%09d.ckpt`, signerf_trainer.py:278-306: {"step", "pipeline", "optimizers", ...}
        with the model under `pipeline["_model.*"]`, DDP's `module.` prefix possibly in front) -> fused renderer.  `kw` as
        `from_state_dict`; the per-image appearance table stays in (eval uses its mean)."""
        ckpt = torch.load(str(path), map_location="cpu", weights_only=False)
        pipe = ckpt["pipeline"] if isinstance(ckpt, dict) and "pipeline" in ckpt else ckpt
        model = {k[len("_model."):]: v for k, v in pipe.items() if k.startswith("_model.")}
        if model and all(k.startswith("module.") for k in model):
            model = {k[len("module."):]: v for k, v in model.items()}
        if not model:
            raise KeyError("no `_model.` tensors in the checkpoint's pipeline state")
        return cls.from_state_dict(model, device=device, **kw)

    # nerfacto checkpoint -> field (names per SURVEY §8c; torch-fallback `implementation="torch"` checkpoints)
    @classmethod
    def from_state_dict(cls, sd: Mapping[str, Tensor], device="cuda", num_train_data: Optional[int] = None,
                        average_init_density: float = 0.01, render_opts: Optional[ops.RenderOptions] = None,
                        near_plane: float = 0.05, far_plane: float = 1000.0, allow_flat: bool = False):
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
        if len(pg) != 2 and render_opts is None and not allow_flat:
            # With load_model_with_proposal_weights=False the reference still samples 256 -> 96 -> 48 through FRESHLY
            # initialised proposal networks (signerf_pipeline.py:113-120, strict=False): flat sampling would give other
            # depths / masks without a word.  The caller either hands over the model's live proposal networks or asks
            # for flat sampling explicitly.
            raise KeyError("proposal_networks.{0,1}.* missing from the state_dict: pass the live model's state_dict (its "
                           "re-initialised proposal networks included), or allow_flat=True / explicit render_opts")
        if "field.embedding_appearance.embedding.weight" not in sd:
            import warnings
            warnings.warn("appearance embedding missing from the state_dict: using the mean of a seeded N(0,1) table with "
                          f"{num_train_data or 30} rows; the reference's fresh nn.Embedding has its own random rows")
        fld = NerfactoFieldB200(g, base, head, app, average_init_density, pg, pm)
        if render_opts is None:
            render_opts = (ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96), near_plane=near_plane,
                                             far_plane=far_plane) if len(pg) == 2
                           else ops.RenderOptions(mode="flat", num_samples=48, near_plane=near_plane, far_plane=far_plane))
        return cls(fld, render_opts)
