"""Mirror of reference signerf/signerf_pipeline.py: `SIGNeRFPipelineConfig` / `SIGNeRFPipeline` — the object that owns the
datamanager, the model (the `graph` of the hot path) and the `DatasetGenerator`, and whose `load_state_dict` decides
which nerfacto tensors the renderer sees (signerf_pipeline.py:93-132).

Inside nerfstudio the classes derive from `VanillaPipeline(Config)` exactly as the reference's do; in a container
without nerfstudio (this one) a minimal stand-in with the same constructor contract is used, so that the plugin glue —
`DatasetGeneratorConfig.setup(original_transform_matrix=..., original_scale_factor=...,
transform_poses_to_original_space=..., device=...)` from the train `DataparserOutputs` (:52-57), the intrinsics
back-fill from the first train camera (:60-88), and the checkpoint filter — is exercised by the CPU tests."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Tuple, Type

from .base import InstantiateConfig
from .datasetgenerator import DatasetGenerator, DatasetGeneratorConfig

try:  # the real base classes when the plugin runs inside nerfstudio
    from nerfstudio.pipelines.base_pipeline import VanillaPipeline, VanillaPipelineConfig  # type: ignore
    HAVE_NERFSTUDIO = True
except Exception:  # nerfstudio absent: same constructor / attribute contract, nothing else
    HAVE_NERFSTUDIO = False

    @dataclass
    class VanillaPipelineConfig(InstantiateConfig):  # type: ignore
        _target: Type = field(default_factory=lambda: VanillaPipeline)
        datamanager: Any = None
        model: Any = None

    class VanillaPipeline:  # type: ignore
        """[EXT] nerfstudio VanillaPipeline.__init__: datamanager from its config, model from its config with the
        train scene box / image count, `.model` property, `load_state_dict` fanning out to the model."""

        def __init__(self, config, device: str, test_mode: str = "val", world_size: int = 1, local_rank: int = 0,
                     grad_scaler=None) -> None:
            self.config, self.device, self.test_mode, self.world_size = config, device, test_mode, world_size
            self.datamanager = config.datamanager.setup(device=device, test_mode=test_mode, world_size=world_size,
                                                        local_rank=local_rank)
            ds = self.datamanager.train_dataset
            self._model = config.model.setup(scene_box=getattr(ds, "scene_box", None), num_train_data=len(ds),
                                             metadata=getattr(ds, "metadata", {}), device=device, grad_scaler=grad_scaler)

        @property
        def model(self):
            return self._model

        def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
            own = getattr(self.datamanager, "load_state_dict", None)
            if own is not None:
                own({k[len("datamanager."):]: v for k, v in state_dict.items() if k.startswith("datamanager.")}, strict=False)


# ---------------------------------------------------------------------------------------------- checkpoint filter
DROPPED_MODEL_KEYS = ("field.embedding_appearance.embedding.weight", "camera_optimizer.pose_adjustment")
DROPPED_PIPELINE_KEYS = ("datamanager.train_camera_optimizer.pose_adjustment",
                         "datamanager.train_ray_generator.pose_optimizer.pose_adjustment")


def split_checkpoint_state(state_dict: Mapping[str, Any], keep_proposal_weights: bool) -> Tuple[Dict[str, Any], Dict[str, Any]]:
    """signerf_pipeline.py:93-120 as a pure function: (model_state, pipeline_state) of a nerfstudio pipeline checkpoint.
    `_model.` prefixes go (and DDP's `module.` when EVERY model key carries it); the per-image appearance table and all
    camera-pose refinements are dropped (the generated dataset has other images and cameras); proposal networks go too
    when the sampler is to be re-learnt (`load_model_with_proposal_weights=False`)."""
    model_state = {k[len("_model."):]: v for k, v in state_dict.items() if k.startswith("_model.")}
    if all(k.startswith("module.") for k in model_state):        # vacuously true for no model keys, like the reference
        model_state = {k[len("module."):]: v for k, v in model_state.items()}
    pipeline_state = {k: v for k, v in state_dict.items() if not k.startswith("_model.")}
    for k in DROPPED_MODEL_KEYS:
        model_state.pop(k, None)
    for k in DROPPED_PIPELINE_KEYS:
        pipeline_state.pop(k, None)
    return model_state, pipeline_state


def drop_proposal_weights(model_state: Dict[str, Any]) -> None:
    for k in [k for k in model_state if k.startswith("proposal")]:
        del model_state[k]


@dataclass
class SIGNeRFPipelineConfig(VanillaPipelineConfig):
    """signerf_pipeline.py:20-29 (the reference defaults `datamanager` to SIGNeRFDataManagerConfig, which lives in the
    out-of-scope data stack: here the field is whatever the method config passes, as nerfstudio resolves it)."""
    _target: Type = field(default_factory=lambda: SIGNeRFPipeline)
    datamanager: Any = None
    dataset_generator: DatasetGeneratorConfig = field(default_factory=DatasetGeneratorConfig)


class SIGNeRFPipeline(VanillaPipeline):
    """signerf_pipeline.py:31-157"""

    config: SIGNeRFPipelineConfig

    def __init__(self, config: SIGNeRFPipelineConfig, device: str, base_dir: Optional[Path] = None, test_mode: str = "val",
                 world_size: int = 1, local_rank: int = 0, grad_scaler=None, load_model_with_proposal_weights: bool = True):
        super().__init__(config, device, test_mode, world_size, local_rank)
        outputs = self.datamanager.train_dataparser_outputs          # the DataparserOutputs contract
        self.dataset_generator: DatasetGenerator = config.dataset_generator.setup(
            original_transform_matrix=outputs.dataparser_transform, original_scale_factor=outputs.dataparser_scale,
            transform_poses_to_original_space=outputs.transform_poses_to_original_space, device=device)
        cams = self.datamanager.train_dataset.cameras
        for name in ("fx", "fy", "cx", "cy", "width", "height"):       # unset intrinsics come from the first train camera
            if getattr(self.dataset_generator, name) is None:
                first = getattr(cams, name)[0].item()
                setattr(self.dataset_generator, name, first)
                setattr(self.dataset_generator.config, name, first)
        self.load_model_with_proposal_weights = load_model_with_proposal_weights
        self.model_state_dict: Optional[Dict[str, Any]] = None

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):  # noqa: ARG002 - always non-strict, as the reference
        model_state, pipeline_state = split_checkpoint_state(state_dict, self.load_model_with_proposal_weights)
        self.model_state_dict = model_state                         # kept for reload_model_state_dict_without_proposal_weights
        if not self.load_model_with_proposal_weights:
            drop_proposal_weights(model_state)
        self.model.load_state_dict(model_state, strict=False)
        super().load_state_dict(pipeline_state, strict=False)

    def reload_model_state_dict_without_proposal_weights(self) -> None:
        print("Reloading model without proposal weights")
        if self.model_state_dict is not None:
            drop_proposal_weights(self.model_state_dict)
            self.model.load_state_dict(self.model_state_dict, strict=False)

    def get_training_callbacks(self, training_callback_attributes) -> List[Any]:
        return (self.datamanager.get_training_callbacks(training_callback_attributes)
                + self.model.get_training_callbacks(training_callback_attributes))

    def forward(self):
        pass

    # ------------------------------------------------------------------ the training step on the CUDA kernels (opt-in)
    def enable_fused_training(self, **kw) -> None:
        """Route `get_train_loss_dict` ([EXT] VanillaPipeline.get_train_loss_dict: next_train -> model forward ->
        get_metrics_dict -> get_loss_dict) through `signerf_b200.train.FusedTrainingStep`: same batches from the same
        datamanager, the same loss-dict keys, gradients into the model's own parameters - the trainer's optimizers,
        schedulers, GradScaler and checkpoints do not notice.  `kw`: FusedTrainingStep options."""
        from ..train import FusedTrainingStep
        cfg = getattr(self.model, "config", None)
        opts = dict(near=float(getattr(cfg, "near_plane", 0.05)), far=float(getattr(cfg, "far_plane", 1000.0)),
                    average_init_density=float(getattr(cfg, "average_init_density", 0.01)), use_l1=bool(getattr(cfg, "use_l1", True)),
                    interlevel_loss_mult=float(getattr(cfg, "interlevel_loss_mult", 1.0)),
                    distortion_loss_mult=float(getattr(cfg, "distortion_loss_mult", 0.002)),
                    predict_normals=bool(getattr(cfg, "predict_normals", False)) or None,
                    orientation_loss_mult=float(getattr(cfg, "orientation_loss_mult", 0.0001)),
                    pred_normal_loss_mult=float(getattr(cfg, "pred_normal_loss_mult", 0.001)),
                    counts=tuple(getattr(cfg, "num_proposal_samples_per_ray", (256, 96))) + (int(getattr(cfg, "num_nerf_samples_per_ray", 48)),))
        opts.update(kw)
        self._fused_step = FusedTrainingStep(self.model, **opts)

    def get_train_loss_dict(self, step: int):
        fused = getattr(self, "_fused_step", None)
        if fused is None:
            return super().get_train_loss_dict(step)
        ray_bundle, batch = self.datamanager.next_train(step)
        cams = ray_bundle.camera_indices.reshape(-1).to(dtype=__import__("torch").int32)
        image = batch["image"].to(ray_bundle.origins.device)[..., :3]
        # the LPIPS term of signerf.py:49-60 stays the MODEL's own module (torchmetrics' pretrained network, a host-side torch
        # term): it sees the rendered rgb as patches and its gradient joins the CUDA backward through `extra_loss`
        extra, cfg = None, getattr(self.model, "config", None)
        if getattr(cfg, "use_lpips", False) and getattr(self.model, "lpips", None) is not None:
            ps, mult, lpips = int(getattr(cfg, "patch_size", 32)), float(getattr(cfg, "lpips_loss_mult", 1.0)), self.model.lpips

            def patches(t):
                return (t.reshape(-1, ps, ps, 3).permute(0, 3, 1, 2) * 2 - 1).clamp(-1, 1)

            gt = patches(image.reshape(-1, 3))
            extra = lambda rgb: mult * lpips(patches(rgb), gt)                     # noqa: E731
        loss_dict = fused.loss_dict(ray_bundle.origins.reshape(-1, 3), ray_bundle.directions.reshape(-1, 3), image, cams,
                                    extra_loss=extra, step=step)
        return {}, loss_dict, {"distortion": loss_dict["distortion_loss"].detach() / max(fused.trainer.distortion_mult, 1e-30)}
