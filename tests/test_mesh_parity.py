"""GPU: SURVEY §8(f) row 2 — proxy-mesh depth rasteriser and render_camera's masking_mode == "shape" branch through
the C ABI against oracle/mesh_ref.py."""
import numpy as np
import pytest
import torch

from oracle import mesh_ref as M
from oracle import sheet_ref as S
from tests.helpers import ring_cameras

pytestmark = pytest.mark.gpu


def _views(n, W, H):
    c2w, intr = ring_cameras(n, W, H)
    return c2w, intr


@pytest.mark.parametrize("cull", [True, False])
def test_rasteriser_matches_oracle_bit_for_bit(cull):
    from signerf_b200 import ops
    H, W, V = 72, 96, 4
    c2w, intr = _views(V, W, H)
    v, f = M.uv_sphere(1.0, 16, 24)
    model = M.object_pose([0.02, -0.03, 0.01], [20, -35, 60], [0.012, 0.02, 0.016])
    got = ops.rasterize_depth(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), model, c2w.cuda(), intr.cuda(), H, W,
                              cull_back=cull).cpu().numpy()
    covered = 0
    for i in range(V):
        ref = M.rasterize_depth(v, f, model, c2w[i].numpy(), intr[i].tolist(), H, W, cull_back=cull)
        assert np.array_equal(got[i, ..., 0], ref)
        covered += int((ref > 0).sum())
    assert covered > 1000


def test_rasteriser_edge_cases():
    from signerf_b200 import ops
    H, W = 32, 48
    c2w, intr = _views(2, W, H)
    empty_v, empty_f = torch.zeros(0, 3).cuda(), torch.zeros(0, 3, dtype=torch.int32).cuda()
    d = ops.rasterize_depth(empty_v, empty_f, np.eye(4), c2w.cuda(), intr.cuda(), H, W)
    assert tuple(d.shape) == (2, H, W, 1) and float(d.abs().max()) == 0.0
    # a triangle behind the camera and one far outside the frustum leave the image empty
    v = torch.tensor([[0, 0, 5.0], [1, 0, 5.0], [0, 1, 5.0], [50, 50, -1.0], [51, 50, -1.0], [50, 51, -1.0]])
    f = torch.tensor([[0, 1, 2], [3, 4, 5]], dtype=torch.int32)
    cam = torch.eye(4)[None, :3]
    d = ops.rasterize_depth(v.cuda(), f.cuda(), M.CONVERT, cam.cuda(), intr[:1].cuda(), H, W, cull_back=False)
    assert float(d.abs().max()) == 0.0
    with pytest.raises(TypeError):
        ops.rasterize_depth(v.cuda(), f.long().cuda(), np.eye(4), cam.cuda(), intr[:1].cuda(), H, W)


@pytest.mark.parametrize("inverse,dilation,manual", [(False, (7, 7), None), (True, (5, 9), None), (False, None, (0.2, 0.9))])
def test_shape_mask_condition_matches_oracle(inverse, dilation, manual):
    from signerf_b200 import ops
    H, W, V = 64, 80, 3
    g = torch.Generator().manual_seed(4)
    c2w, intr = _views(V, W, H)
    v, f = M.uv_sphere(1.0, 12, 20)
    model = M.object_pose([0.0, 0.0, 0.0], [0, 0, 0], [0.015, 0.015, 0.015])
    proxy = ops.rasterize_depth(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), model, c2w.cuda(), intr.cuda(), H, W)
    nerf = (0.3 + 0.4 * torch.rand(V, H, W, 1, generator=g)).cuda()
    nerf[2] = 0.01                                       # view 2: the NeRF surface hides the whole proxy
    opts = ops.MaskOptions(inverse_mask=inverse, mask_dilation=dilation, additional_depth_radius=0.1, manual_depth=manual)
    mask, cond, stats = ops.mask_condition_shape(proxy, nerf, opts)
    for i in range(V):
        rm, rc, vis = M.shape_mask_condition(proxy[i, ..., 0].cpu().numpy(), nerf[i, ..., 0].cpu().numpy(), inverse, dilation, 0.1, manual)
        assert bool(stats[i, 0]) == vis
        assert np.array_equal(mask[i, ..., 0].cpu().numpy().astype(bool), rm)
        assert np.allclose(cond[i, ..., 0].cpu().numpy(), rc, atol=1e-6)
    if not inverse:
        assert not bool(stats[2, 0]) and float(cond[2].abs().max()) == 0.0 and int(mask[2].sum()) == 0


def test_plugin_renderer_and_shape_mode(tmp_path):
    """plugin.Renderer from an .obj on disk + DatasetGenerator(masking_mode='shape').render_camera."""
    from signerf_b200 import plugin as P
    from signerf_b200 import ops
    v, f = M.uv_sphere(1.0, 10, 16)
    obj = tmp_path / "ball.obj"
    obj.write_text("".join(f"v {a} {b} {c}\n" for a, b, c in v.tolist()) + "".join(f"f {a + 1}//1 {b + 1}//1 {c + 1}//1\n" for a, b, c in f.tolist()))
    r = P.Renderer(P.RendererConfig(object_path=str(obj), scale=[0.01, 0.01, 0.01], color=[1.0, 0.0, 0.0, 1.0]), "cuda")
    r.setup()
    assert r.scene is not None
    H = W = 64
    c2w, intr = _views(1, W, H)
    cam = P.CameraBatch(c2w, float(intr[0, 0]), float(intr[0, 1]), float(intr[0, 2]), float(intr[0, 3]), W, H)
    color, depth = r.render_camera(cam)
    ref = M.rasterize_depth(v, f, M.object_pose([0, 0, 0], [0, 0, 0], [0.01, 0.01, 0.01]), c2w[0].numpy(), intr[0].tolist(), H, W)
    assert color.dtype == torch.uint8 and tuple(color.shape) == (H, W, 3) and tuple(depth.shape) == (H, W, 1)
    assert np.array_equal(depth[..., 0].cpu().numpy(), ref) and (ref > 0).sum() > 50
    # pyrender lights the mesh with ambient 1.0 only: flat material colour (its default baseColorFactor 0.3 -> 77) on covered
    # pixels, white background elsewhere; RendererConfig.color is never handed to pyrender by the reference
    assert color[depth[..., 0] > 0].unique().tolist() == [77]
    assert int(color[depth[..., 0] == 0].min()) == 255
    r.position = [0.05, 0.0, 0.0]                # GUI edit + setup(), as interface.py:375-377
    r.setup()
    moved = r.render_camera(cam)[1]
    assert not torch.equal(moved, depth)
