"""CPU: the host-side mirror keeps the reference's plugin surface — dataclass fields and method argument names extracted
from the reference sources (tests/golden/plugin_surface.json, made by tests/golden/make_plugin_surface.py)."""
import dataclasses
import inspect
import json
import os

import pytest
import torch

import signerf_b200.plugin as P


@pytest.fixture(scope="module")
def surface(golden_dir):
    return json.load(open(os.path.join(golden_dir, "plugin_surface.json")))


@pytest.mark.parametrize("name", ["DatasetGeneratorConfig", "DiffuserConfig", "RendererConfig"])
def test_config_fields_match_reference(surface, name):
    mine = [f.name for f in dataclasses.fields(getattr(P, name))]
    assert mine == surface[name]["fields"]


@pytest.mark.parametrize("cls,methods", [("DatasetGenerator", ["__init__", "render_camera", "generate_reference_sheet",
                                                               "generate_with_reference_sheet"]),
                                         ("Diffuser", ["__init__", "diffuse", "_custom"]),
                                         ("Renderer", ["__init__", "setup", "render_camera"])])
def test_method_signatures_match_reference(surface, cls, methods):
    for m in methods:
        mine = list(inspect.signature(getattr(getattr(P, cls), m)).parameters)
        assert mine == surface[cls]["methods"][m], (cls, m)


def test_defaults_and_error_behaviour():
    c = P.DatasetGeneratorConfig()
    assert (c.rows, c.cols, c.downscale_factor, c.mask_dialation, c.additional_depth_radius) == (2, 3, 2, (50, 50), 0.1)
    d = P.DiffuserConfig()
    assert (d.num_inference_steps, d.denoising_strength, d.guidance_scale, d.seed, d.controlnet_conditioning_scale) == (20, 0.9, 7, 1, 0.8)
    gen = c.setup(original_transform_matrix=torch.eye(4)[:3], original_scale_factor=1.0,
                  transform_poses_to_original_space=lambda x: x, device="cpu")
    assert isinstance(gen, P.DatasetGenerator) and gen.aabb.shape == (2, 3) and gen.diffuser.url == "http://127.0.0.1:5000"
    cams = P.CameraBatch(torch.zeros(3, 3, 4), 64.0, 64.0, 32.0, 32.0, 64, 64)
    with pytest.raises(ValueError, match="Camera count 3 is not equal"):      # datasetgenerator.py:494-495
        gen.generate_reference_sheet(None, cams, 32, 32)
    diff = P.Diffuser(P.DiffuserConfig(mode="custom"), "cpu")
    with pytest.raises(RuntimeError, match="backend"):
        diff.diffuse(torch.zeros(8, 8, 3), torch.zeros(8, 8, 3), torch.zeros(8, 8, 1), torch.zeros(8, 8, 1))
    r = P.Renderer(P.RendererConfig(object_path="no/such/mesh.obj"), "cpu")
    assert (r.scale, r.color, r.rotation) == ([0.1, 0.1, 0.1], [0.0, 0.0, 0.0, 1.0], [0, 0, 0])     # renderer.py:30-37
    r.setup()                                   # renderer.py:68-75: prints and returns on a bad path
    assert r.scene is None
    with pytest.raises(AttributeError):         # the reference dereferences self.scene = None (renderer.py:183)
        r.render_camera(cams)
    r.object_path = "mesh.ply"
    r.setup()
    assert r.scene is None


def test_renderer_pose_and_obj_loader_match_oracle(tmp_path):
    """plugin.Renderer's object pose (renderer.py:80-131) and OBJ reader against oracle/mesh_ref.py, on the CPU."""
    import numpy as np
    from oracle import mesh_ref as M
    from signerf_b200.plugin.renderer import load_obj
    r = P.Renderer(P.RendererConfig(position=[0.1, -0.2, 0.3], rotation=[10, 20, 30], scale=[0.1, 0.2, 0.3]), "cpu")
    assert np.allclose(r.object_pose(), M.object_pose([0.1, -0.2, 0.3], [10, 20, 30], [0.1, 0.2, 0.3]), atol=1e-12)
    r.rotation, r.scale = [0, 0, 90], [0.1, 0.1, 0.1]                       # GUI edit
    pose = r.object_pose()
    assert np.allclose(pose[:3, :3] @ np.array([1.0, 0, 0]), [0.0, 0.0, -1.0])   # x -> y (Rz 90), y -> -z_gl (axis swap), scale 10 * 0.1
    text = "o quad\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 3/1/1 4/1/1\n"
    (tmp_path / "q.obj").write_text(text)
    v, f = load_obj(tmp_path / "q.obj")
    vo, fo = M.parse_obj(text)
    assert np.array_equal(v, vo) and np.array_equal(f, fo) and f.tolist() == [[0, 1, 2], [0, 2, 3]]
