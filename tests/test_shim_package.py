"""CPU: the `signerf` overlay package — the reference's module paths and entry points resolve to this repo's classes,
the rest of the reference falls through — and the SIGNeRFPipeline mirror's plugin glue (signerf_pipeline.py:36-144) on
duck-typed nerfstudio stand-ins (datamanager / DataparserOutputs / Cameras / model)."""
import inspect
import json
import os
import subprocess
import sys
import tomllib

import pytest
import torch

import signerf_b200.plugin as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def surface(golden_dir):
    return json.load(open(os.path.join(golden_dir, "plugin_surface.json")))


def test_reference_module_paths_resolve_to_the_b200_classes():
    import signerf.datasetgenerator.datasetgenerator as dg
    import signerf.diffuser.diffuser as df
    import signerf.renderer.renderer as rd
    import signerf.signerf_pipeline as pl
    assert dg.DatasetGenerator is P.DatasetGenerator and dg.DatasetGeneratorConfig is P.DatasetGeneratorConfig
    assert df.Diffuser is P.Diffuser and df.DiffuserConfig is P.DiffuserConfig
    assert rd.Renderer is P.Renderer and rd.RendererConfig is P.RendererConfig and rd.NERFSTUDIO_BLENDER_SCALE_RATIO == 10.0
    assert pl.SIGNeRFPipeline is P.SIGNeRFPipeline and pl.SIGNeRFPipelineConfig is P.SIGNeRFPipelineConfig


def test_entry_points_are_the_references(surface):
    proj = tomllib.load(open(os.path.join(ROOT, "pyproject.toml"), "rb"))
    assert proj["project"]["entry-points"]["nerfstudio.method_configs"] == surface["entry_points"]   # reference pyproject.toml:44-46
    assert proj["tool"]["setuptools"]["packages"]["find"]["include"] == ["signerf*"]


def test_overlay_falls_through_to_a_reference_checkout(tmp_path):
    """Modules this repo does not provide (trainer, data, interface, utils, method configs) import from the reference's
    `signerf/` directory; the four hot-path modules keep shadowing their namesakes there."""
    ref = tmp_path / "checkout" / "signerf"
    (ref / "utils").mkdir(parents=True)
    (ref / "renderer").mkdir()
    (ref / "signerf_trainer.py").write_text("WHO = 'reference trainer'\n")
    (ref / "utils" / "__init__.py").write_text("")
    (ref / "utils" / "poses_generation.py").write_text("WHO = 'reference utils'\n")
    (ref / "renderer" / "__init__.py").write_text("")
    (ref / "renderer" / "renderer.py").write_text("WHO = 'reference renderer (must be shadowed)'\n")
    (ref / "signerf_config.py").write_text("from signerf.signerf_pipeline import SIGNeRFPipelineConfig\nsignerf_method = SIGNeRFPipelineConfig\n")
    code = ("import signerf, signerf.signerf_trainer as t, signerf.utils.poses_generation as u, signerf.renderer.renderer as r, "
            "signerf.signerf_config as c; import signerf_b200.plugin as P; "
            "assert t.WHO == 'reference trainer' and u.WHO == 'reference utils'; assert r.Renderer is P.Renderer; "
            "assert c.signerf_method is P.SIGNeRFPipelineConfig; print('ok', len(signerf.__path__))")
    for env_extra, path_extra in (({"SIGNERF_REFERENCE_DIR": str(ref)}, []), ({}, [str(tmp_path / "checkout")])):
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT] + path_extra), **env_extra)
        env.pop("SIGNERF_REFERENCE_DIR", None) if not env_extra else None
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.strip() == "ok 2", out.stderr[-2000:]


def test_pipeline_surface_matches_reference(surface):
    own = [k for k in P.SIGNeRFPipelineConfig.__annotations__]
    assert own == surface["SIGNeRFPipelineConfig"]["fields"]
    for m, args in surface["SIGNeRFPipeline"]["methods"].items():
        assert list(inspect.signature(getattr(P.SIGNeRFPipeline, m)).parameters) == args, m


# ---------------------------------------------------------------------------------------------- nerfstudio stand-ins
class _Cameras:
    def __init__(self, n):
        self.camera_to_worlds = torch.eye(4)[:3].repeat(n, 1, 1)
        self.fx, self.fy = torch.full((n, 1), 500.0), torch.full((n, 1), 510.0)
        self.cx, self.cy = torch.full((n, 1), 320.0), torch.full((n, 1), 240.0)
        self.width, self.height = torch.full((n, 1), 640), torch.full((n, 1), 480)


class _DataparserOutputs:
    dataparser_transform = torch.tensor([[0.0, 1, 0, 0.5], [1, 0, 0, 0.25], [0, 0, 1, 0.125]])
    dataparser_scale = 0.37

    @staticmethod
    def transform_poses_to_original_space(poses):
        return poses


class _Dataset:
    def __init__(self):
        self.cameras, self.scene_box, self.metadata = _Cameras(7), None, {}

    def __len__(self):
        return 7


class _DataManager:
    def __init__(self, **kw):
        self.kw, self.train_dataset, self.train_dataparser_outputs = kw, _Dataset(), _DataparserOutputs()
        self.loaded = None

    def load_state_dict(self, sd, strict=True):
        self.loaded = dict(sd)

    def get_training_callbacks(self, attrs):
        return ["dm"]


class _Model(torch.nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.kw = kw
        self.field = torch.nn.Linear(2, 2)
        self.proposal_networks = torch.nn.ModuleList([torch.nn.Linear(2, 2)])
        self.loads = []

    def load_state_dict(self, sd, strict=True):
        self.loads.append((sorted(sd), strict))

    def get_training_callbacks(self, attrs):
        return ["model"]


class _Cfg:
    def __init__(self, cls):
        self.cls = cls

    def setup(self, **kw):
        return self.cls(**kw)


def _pipeline(**kw):
    cfg = P.SIGNeRFPipelineConfig(datamanager=_Cfg(_DataManager), model=_Cfg(_Model),
                                  dataset_generator=P.DatasetGeneratorConfig(fx=None, width=320))
    return P.SIGNeRFPipeline(cfg, "cpu", **kw)


@pytest.mark.skipif(P.pipeline.HAVE_NERFSTUDIO, reason="stand-in base classes are only used without nerfstudio")
def test_pipeline_wires_the_dataset_generator_from_the_dataparser_outputs():
    pipe = _pipeline(test_mode="inference", world_size=1, local_rank=0)
    gen = pipe.dataset_generator
    assert isinstance(gen, P.DatasetGenerator)
    assert torch.equal(gen.original_transform_matrix, _DataparserOutputs.dataparser_transform) and gen.original_scale_factor == 0.37
    assert gen.transform_poses_to_original_space is _DataparserOutputs.transform_poses_to_original_space
    # unset intrinsics are back-filled from the first train camera, on the object AND its config (signerf_pipeline.py:60-88)
    assert (gen.fx, gen.fy, gen.cx, gen.cy, gen.width, gen.height) == (500.0, 510.0, 320.0, 240.0, 320, 480)
    assert (gen.config.fx, gen.config.height, gen.config.width) == (500.0, 480, 320)
    assert pipe.model.kw["num_train_data"] == 7 and pipe.load_model_with_proposal_weights and pipe.model_state_dict is None
    assert pipe.get_training_callbacks(None) == ["dm", "model"] and pipe.forward() is None


@pytest.mark.skipif(P.pipeline.HAVE_NERFSTUDIO, reason="stand-in base classes are only used without nerfstudio")
@pytest.mark.parametrize("ddp", [False, True])
@pytest.mark.parametrize("keep_proposals", [True, False])
def test_pipeline_load_state_dict_filters_like_the_reference(ddp, keep_proposals):
    pre = "_model.module." if ddp else "_model."
    ckpt = {pre + "field.mlp_base.encoding.hash_table": 1, pre + "field.embedding_appearance.embedding.weight": 2,
            pre + "proposal_networks.0.mlp_base.encoding.hash_table": 3, pre + "camera_optimizer.pose_adjustment": 4,
            "datamanager.train_camera_optimizer.pose_adjustment": 5,
            "datamanager.train_ray_generator.pose_optimizer.pose_adjustment": 6, "datamanager.other": 7}
    pipe = _pipeline(load_model_with_proposal_weights=keep_proposals)
    pipe.load_state_dict(ckpt)
    keys, strict = pipe.model.loads[-1]
    want = ["field.mlp_base.encoding.hash_table"] + (["proposal_networks.0.mlp_base.encoding.hash_table"] if keep_proposals else [])
    assert keys == want and strict is False
    assert pipe.datamanager.loaded == {"other": 7}
    assert sorted(pipe.model_state_dict) == want
    pipe.reload_model_state_dict_without_proposal_weights()          # signerf_pipeline.py:135-144
    assert pipe.model.loads[-1] == (["field.mlp_base.encoding.hash_table"], False)


def test_checkpoint_split_keeps_module_prefix_when_not_ddp():
    # a model attribute that happens to be called `module` is not a DDP wrapper unless EVERY key carries the prefix
    ms, ps = P.pipeline.split_checkpoint_state({"_model.module.a": 1, "_model.b": 2, "step": 3}, True)
    assert ms == {"module.a": 1, "b": 2} and ps == {"step": 3}
