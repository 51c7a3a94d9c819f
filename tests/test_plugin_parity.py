"""GPU: the plugin-level entry points (plugin.DatasetGenerator.generate_reference_sheet / generate_with_reference_sheet
over the fused kernels) against the oracle's restatement of the reference's per-view Python loops
(oracle/nerfacto_ref.py + oracle/sheet_ref.py), with a deterministic stand-in for the diffuser so that the sheet
assembly, blending and tile cut-out are compared exactly."""
import pytest
import torch

import signerf_b200.plugin as P
from oracle import nerfacto_ref as R
from oracle import sheet_ref as S
from signerf_b200 import ops
from tests.helpers import field_from_oracle, rel_l2, ring_cameras

pytestmark = pytest.mark.gpu


class _StubDiffuser(P.Diffuser):
    """edited = 1 - 0.5 * image where masked; returned on the CPU like the reference's HTTP client does."""

    def diffuse(self, original_image, rendered_image, mask_image=None, condition_image=None):
        return (1.0 - 0.5 * original_image).cpu()


def _generator(rows, cols, H, W, ds):
    cfg = P.DatasetGeneratorConfig(rows=rows, cols=cols, width=W, height=H, downscale_factor=ds, mask_dialation=(7, 7),
                                   fx=float(W), fy=float(W), cx=W / 2, cy=H / 2)
    gen = cfg.setup(original_transform_matrix=torch.eye(4)[:3], original_scale_factor=1.0,
                    transform_poses_to_original_space=lambda x: x, device="cuda")
    gen.diffuser = _StubDiffuser(cfg.diffuser, "cuda")
    return gen


def _oracle_views(m, c2w, intr, H, W, gen, S_samples):
    renders, masks, conds = [], [], []
    for v in range(c2w.shape[0]):
        ref = R.render_view(m, c2w[v], *intr[v].tolist(), W, H, "flat", S_samples)
        rays = R.generate_rays(c2w[v], *intr[v].tolist(), W, H)
        mk, cd, _ = S.render_camera_aabb(rays.origins.view(H, W, 3), rays.directions.view(H, W, 3), ref["depth"],
                                         gen.aabb.cpu(), False, (7, 7), 0.1, None)
        renders.append(ref["rgb"]), masks.append(mk), conds.append(cd)
    return renders, masks, conds


def test_generate_reference_sheet_and_with_reference_sheet():
    rows, cols, H, W, ds, Ssamp = 2, 2, 48, 40, 2, 24
    th, tw = H // ds, W // ds
    m = R.make_model(0, dense=True, table_scale=0.5, density_gain=20.0)
    graph = P.FusedNerfactoGraph(field_from_oracle(m), ops.RenderOptions(mode="flat", num_samples=Ssamp, mlp_mode=ops.MLP_FP32))
    gen = _generator(rows, cols, H, W, ds)
    c2w, intr = ring_cameras(rows * cols, W, H)
    cams = P.CameraBatch(c2w[:-1], float(W), float(W), W / 2, H / 2, W, H)
    img, msk, cnd, edited, refs = gen.generate_reference_sheet(graph, cams, tw, th)
    renders, masks, conds = _oracle_views(m, c2w[:-1], intr[:-1], H, W, gen, Ssamp)
    img_r, msk_r, cnd_r = S.reference_sheet(renders, masks, conds, rows, cols, th, tw, 0)
    assert img.shape == img_r.shape == (2 * th, 2 * tw, 3)
    assert rel_l2(img, img_r) < 2e-5 and rel_l2(cnd, cnd_r) < 1e-3
    assert float((msk.cpu() != msk_r).float().mean()) < 2e-3            # median-depth bin flips at the box boundary
    edited_r = S.blend(1.0 - 0.5 * img_r, img_r, msk_r)
    same = (msk.cpu() == msk_r).expand_as(edited_r)
    assert torch.allclose(edited.cpu()[same], edited_r[same], atol=1e-4)
    assert len(refs) == 3 and set(refs[0]) == {"render", "mask", "condition", "render_scaled", "mask_scaled",
                                               "condition_scaled", "edited", "edited_scaled"}
    for i in range(3):
        assert rel_l2(refs[i]["render"], renders[i]) < 2e-5
        assert refs[i]["edited"].shape == (H, W, 3) and refs[i]["edited_scaled"].shape == (th, tw, 3)
        assert torch.allclose(refs[i]["edited"].cpu(), S.cut_tile(edited.cpu(), i, cols, th, tw, 0, H, W), atol=1e-6)
    # per-dataset-camera pass: the last tile is overwritten IN PLACE in both sheet arguments (Appendix A.2)
    cam_last = P.CameraBatch(c2w[-1:], float(W), float(W), W / 2, H / 2, W, H)
    img_arg, cnd_arg = edited.clone(), cnd.clone()
    out = gen.generate_with_reference_sheet(graph, cam_last, None, tw, th, img_arg, cnd_arg)
    r_l, m_l, c_l = _oracle_views(m, c2w[-1:], intr[-1:], H, W, gen, Ssamp)
    rs = S._interp(r_l[0], th, tw)
    assert rel_l2(img_arg[th:, tw:], rs) < 2e-5 and torch.equal(img_arg[:th], edited[:th])
    assert rel_l2(out["render_scaled"], rs) < 2e-5 and out["edited"].shape == (H, W, 3)
    ms = S._interp(m_l[0].float(), th, tw) > 0.5
    exp = (1.0 - 0.5 * rs) * ms + rs * (~ms)
    ok = (out["mask_scaled"].cpu() == ms).expand_as(exp)
    assert float(ok.float().mean()) > 0.995 and torch.allclose(out["edited_scaled"].cpu()[ok], exp[ok], atol=1e-4)


def test_render_camera_return_arity_quirk():
    m = R.make_model(0, dense=True)
    graph = P.FusedNerfactoGraph(field_from_oracle(m), ops.RenderOptions(mode="flat", num_samples=8))
    gen = _generator(2, 2, 16, 16, 1)
    c2w, _ = ring_cameras(1, 16, 16)
    cam = P.CameraBatch(c2w, 16.0, 16.0, 8.0, 8.0, 16, 16)
    assert len(gen.render_camera(graph, cam)) == 3
    assert len(gen.render_camera(graph, cam, with_mask=False)) == 4          # datasetgenerator.py:708
    assert len(gen.render_camera(graph, cam, with_condition=False)) == 4     # :783
    rgb, mask, cond = gen.render_camera(graph, cam)
    assert rgb.shape == (16, 16, 3) and mask.dtype == torch.bool and cond.shape == (16, 16, 1)


@pytest.mark.parametrize("style", ["mlp_with_hash_encoding", "split_modules"])
def test_graph_from_nerfacto_state_dict(style):
    """FusedNerfactoGraph.from_state_dict maps a torch-fallback nerfacto checkpoint (both parameter layouts) onto the
    fused field: renders are bit-identical to the field built directly from the same tensors."""
    m = R.make_model(0, dense=True, table_scale=0.5, density_gain=20.0)
    f = m.field
    a, b, c, d = (("field.mlp_base.encoding", "field.mlp_base.mlp", "proposal_networks.%d.mlp_base.encoding",
                   "proposal_networks.%d.mlp_base.mlp") if style == "mlp_with_hash_encoding" else
                  ("field.mlp_base_grid", "field.mlp_base_mlp", "proposal_networks.%d.encoding", "proposal_networks.%d.mlp_base"))
    sd = {a + ".hash_table": f.encoding.hash_table, "field.embedding_appearance.embedding.weight": f.embedding_appearance.weight}
    for i, l in enumerate(f.mlp_base.layers):
        sd[f"{b}.layers.{i}.weight"], sd[f"{b}.layers.{i}.bias"] = l.weight, l.bias
    for i, l in enumerate(f.mlp_head.layers):
        sd[f"field.mlp_head.layers.{i}.weight"], sd[f"field.mlp_head.layers.{i}.bias"] = l.weight, l.bias
    for j, pn in enumerate(m.proposal_networks):
        sd[(c % j) + ".hash_table"] = pn.encoding.hash_table
        for i, l in enumerate(pn.mlp.layers):
            sd[f"{d % j}.layers.{i}.weight"], sd[f"{d % j}.layers.{i}.bias"] = l.weight, l.bias
    graph = P.FusedNerfactoGraph.from_state_dict(sd, average_init_density=f.average_init_density)
    direct = P.FusedNerfactoGraph(field_from_oracle(m))
    assert graph.render_opts.mode == "cascade" and len(graph.field.prop_grids) == 2
    c2w, _ = ring_cameras(2, 24, 16)
    cam = P.CameraBatch(c2w, 24.0, 24.0, 12.0, 8.0, 24, 16)
    o1, o2 = graph.render_cameras(cam), direct.render_cameras(cam)
    assert torch.equal(o1["rgb"], o2["rgb"]) and torch.equal(o1["depth"], o2["depth"])
    # proposal networks dropped (load_model_with_proposal_weights=False): flat sampling of the main field
    sd_np = {k: v for k, v in sd.items() if not k.startswith("proposal_networks")}
    assert P.FusedNerfactoGraph.from_state_dict(sd_np).render_opts.mode == "flat"


def test_render_camera_shape_mode_matches_oracle():
    """masking_mode='shape' (datasetgenerator.py:711-757): proxy-mesh depth from the CUDA rasteriser against the NeRF
    depth, mask / condition against oracle/mesh_ref.py on the same depths."""
    import numpy as np
    from oracle import mesh_ref as M
    H, W, Ssamp = 48, 40, 24
    m = R.make_model(0, dense=True, table_scale=0.5, density_gain=20.0)
    graph = P.FusedNerfactoGraph(field_from_oracle(m), ops.RenderOptions(mode="flat", num_samples=Ssamp, mlp_mode=ops.MLP_FP32))
    gen = _generator(2, 2, H, W, 1)
    gen.masking_mode = "shape"
    v, f = M.uv_sphere(1.0, 10, 16)
    c2w, intr = ring_cameras(2, W, H)
    pos = (0.8 * c2w[0, :3, 3]).tolist()          # a small ball close to camera 0: in front of the dense field's surface
    gen.renderer.scale = [0.002, 0.002, 0.002]
    gen.renderer.position = pos
    gen.renderer.set_mesh(v, f)
    gen.renderer.setup()
    cam = P.CameraBatch(c2w[:1], float(W), float(W), W / 2, H / 2, W, H)
    rgb, mask, cond = gen.render_camera(graph, cam)
    assert tuple(rgb.shape) == (H, W, 3) and tuple(mask.shape) == (H, W, 1) and mask.dtype == torch.bool
    nerf = graph.render_cameras(cam)["depth"][0, ..., 0].cpu().numpy()
    proxy = M.rasterize_depth(v, f, M.object_pose(pos, [0, 0, 0], [0.002] * 3), c2w[0].numpy(), intr[0].tolist(), H, W)
    rm, rc, vis = M.shape_mask_condition(proxy, nerf, False, (7, 7), 0.1, None)
    assert vis and np.array_equal(mask[..., 0].cpu().numpy(), rm) and np.allclose(cond[..., 0].cpu().numpy(), rc, atol=1e-6)
    assert len(gen.render_camera(graph, cam, with_condition=False)) == 4       # the reference's arity quirk survives
    gen.renderer = None
    with pytest.raises(ValueError, match="Renderer is None"):
        gen.render_camera(graph, cam)
