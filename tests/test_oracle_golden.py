"""CPU: pin oracle/sheet_ref.py against fixtures produced by the reference's own functions and cv2
(tests/golden/make_golden.py).  nerfacto_ref.py has no reference fixtures (parity unpinned, see its header);
its self-consistency properties are checked here instead."""
import os

import numpy as np
import torch

from oracle import nerfacto_ref as R
from oracle import sheet_ref as S


def test_aabb_matches_reference(golden_dir):
    a = np.load(os.path.join(golden_dir, "aabb.npz"))
    n, f = S.intersect_with_aabb(torch.tensor(a["o"]), torch.tensor(a["d"]), torch.tensor(a["aabb"]))
    assert np.array_equal(n.numpy(), a["nears"]) and np.array_equal(f.numpy(), a["fars"])


def test_ellipse_and_dilate_match_cv2(golden_dir):
    d = np.load(os.path.join(golden_dir, "dilate.npz"))
    for key, ks in (("se50", (50, 50)), ("se_7x11", (7, 11)), ("se_4x4", (4, 4))):
        assert np.array_equal(S.ellipse_kernel(ks), d[key])
    assert int(d["se50"].sum()) == 1995 and int(d["se50"][0].sum()) == 1
    for i in range(d["masks"].shape[0]):
        assert np.array_equal(S.dilate(d["masks"][i], (50, 50)), d["out50"][i])
        assert np.array_equal(S.dilate(d["masks"][i], (7, 11)), d["out7x11"][i])


def test_quantize_matches_tensor_to_image(golden_dir):
    q = np.load(os.path.join(golden_dir, "quantize.npz"))
    assert np.array_equal(S.quantize_u8(torch.tensor(q["x"])), q["q"])
    # truncation, not rounding (0.999 -> 254)
    assert S.quantize_u8(torch.tensor([[[0.999]]]))[0, 0, 0] == 254


def test_resize_matches_interpolate(golden_dir):
    r = np.load(os.path.join(golden_dir, "resize.npz"))
    img = torch.tensor(r["img"])
    for key, (h, w) in (("down", (10, 14)), ("odd", (13, 9)), ("up", (40, 56)), ("same", (20, 28))):
        assert np.allclose(S._interp(img, h, w).numpy(), r[key], atol=1e-6)
    assert np.array_equal(r["same"], r["img"])


def test_poses_fixture_is_a_look_at_ring(golden_dir):
    p = np.load(os.path.join(golden_dir, "poses.npz"))["poses16"]
    assert p.shape == (16, 4, 4)
    assert np.allclose(np.linalg.norm(p[:, :3, 3], axis=-1), 0.5, atol=1e-6)
    rot = p[:, :3, :3]
    assert np.allclose(rot @ rot.transpose(0, 2, 1), np.eye(3), atol=1e-5)


def test_sheet_size_padding():
    assert S.sheet_size(4, 4, 512, 512, 0) == (2048, 2048)
    assert S.sheet_size(2, 3, 250, 333, 5) == (512, 1016)


# ---- nerfacto restatement: internal consistency (no reference fixtures exist)
def test_hash_scalings_float32_quirk():
    assert R.hash_scalings(16, 16, 2048).tolist()[-1] == 2047.0
    assert R.hash_scalings(5, 16, 128).tolist()[-1] == 128.0


def test_hash_indices_are_int64_and_in_level_range():
    enc = R.HashEncodingRef(num_levels=4, min_res=16, max_res=128, log2_hashmap_size=10)
    idx, off = enc.corner_indices(torch.rand(100, 3))
    assert idx.dtype == torch.int64
    for l in range(4):
        assert int(idx[:, l].min()) >= l * 1024 and int(idx[:, l].max()) < (l + 1) * 1024
    assert float(off.min()) >= 0 and float(off.max()) < 1


def test_flat_bins_and_median_depth_properties():
    edges = R.flat_bin_edges(32, 0.05, 1000.0)
    assert edges.shape == (33,) and abs(float(edges[0]) - 0.05) < 1e-6 and abs(float(edges[-1]) - 1000.0) < 1e-1
    assert bool((edges[1:] > edges[:-1]).all())
    m = R.make_model(0, dense=True, log2_hashmap_size=12)
    o = torch.zeros(8, 3)
    d = torch.nn.functional.normalize(torch.randn(8, 3), dim=-1)
    out = R.render_rays(m, o, d, "flat", 32, return_aux=True)
    mids = (edges[:-1] + edges[1:]) / 2
    for k in range(8):  # median depth is one of the bin mid-points
        assert float((mids - out["depth"][k]).abs().min()) < 1e-6
    assert float(out["rgb"].min()) >= 0 and float(out["rgb"].max()) <= 1


def test_cascade_runs_and_is_deterministic():
    m = R.make_model(0, dense=True, log2_hashmap_size=12, table_scale=0.3, density_gain=10.0)
    o = torch.zeros(4, 3)
    d = torch.nn.functional.normalize(torch.randn(4, 3), dim=-1)
    a = R.render_rays(m, o, d, "cascade")
    b = R.render_rays(m, o, d, "cascade")
    assert torch.equal(a["rgb"], b["rgb"]) and torch.equal(a["depth"], b["depth"])


# ---------------------------------------------------------------------------------------------- A1111 inpaint pre / post
def test_inpaint_oracle_matches_cv2_and_pillow(golden_dir):
    """oracle/inpaint_ref.py bit for bit against cv2.GaussianBlur / PIL resize / paste / alpha_composite outputs
    (tests/golden/make_inpaint_golden.py): mask blur, overlay mask, latent mask, overlay compositing."""
    from oracle import inpaint_ref as I
    g = np.load(os.path.join(golden_dir, "inpaint.npz"))
    for ci in range(4):
        m = g[f"c{ci}_mask"]
        for blur in (4, 2):
            ks = 2 * int(2.5 * blur + 0.5) + 1
            bx = I.gaussian_blur_u8_1d(m, ks, float(blur), 1)
            assert np.array_equal(bx, g[f"c{ci}_blur{blur}_x"])
            assert np.array_equal(I.gaussian_blur_u8_1d(bx, ks, float(blur), 0), g[f"c{ci}_blur{blur}_xy"])
        blurred, overlay = I.a1111_mask_blur(m, 4)
        assert np.array_equal(blurred, g[f"c{ci}_blur4_xy"]) and np.array_equal(overlay, g[f"c{ci}_overlay_mask"])
        lat = g[f"c{ci}_lat_u8"]
        assert np.array_equal(I.pil_resize_bicubic_u8(blurred, lat.shape), lat)
        assert np.array_equal(I.a1111_latent_mask(blurred, lat.shape), g[f"c{ci}_latmask"])
        assert np.array_equal(I.a1111_apply_overlay(g[f"c{ci}_gen"], g[f"c{ci}_orig"], overlay), g[f"c{ci}_composited"])
    for ri in range(4):
        assert np.array_equal(I.pil_resize_bicubic_u8(g[f"r{ri}_in"], g[f"r{ri}_out"].shape), g[f"r{ri}_out"])


def test_gaussian_taps_sum_to_one_and_match_impulse(golden_dir):
    from oracle import inpaint_ref as I
    g = np.load(os.path.join(golden_dir, "inpaint.npz"))
    assert I.gaussian_kernel_q8(21, 4.0).tolist() == [1, 2, 4, 5, 9, 11, 16, 19, 23, 25, 26, 25, 23, 19, 16, 11, 9, 5, 4, 2, 1]
    for sigma, ks in ((4.0, 21), (2.0, 11), (1.5, 9)):
        k = I.gaussian_kernel_q8(ks, sigma)
        assert k.sum() == 256 and np.array_equal(k, k[::-1])
        resp = g[f"impulse_s{sigma}_k{ks}"][0, 2 * ks - ks // 2: 2 * ks + ks // 2 + 1]
        assert np.array_equal(resp, ((k * 255 + 128) >> 8).astype(np.uint8))


def test_vae_oracle_shapes_and_parameter_count():
    """oracle/vae_ref.py (parity unpinned): SDXL-VAE parameter count, /8 latent geometry, posterior sampling."""
    from oracle import vae_ref as VR
    assert sum(p.numel() for p in VR.AutoencoderKL().parameters()) == 83_653_863
    m = VR.make_vae(VR.tiny_vae_config(), seed=0)
    x = torch.rand(1, 3, 32, 48) * 2 - 1
    mom = m.moments(x)
    assert tuple(mom.shape) == (1, 8, 4, 6)
    n = torch.randn(1, 4, 4, 6)
    z0, z1 = m.encode(x), m.encode(x, n)
    mean, logvar = mom.chunk(2, 1)
    assert torch.allclose(z0, VR.SCALE_FACTOR * mean) and torch.allclose(z1, VR.SCALE_FACTOR * (mean + torch.exp(0.5 * logvar) * n))
    assert tuple(m.decode(z0).shape) == (1, 3, 32, 48)


# ---------------------------------------------------------------------------------------------- proxy-mesh depth (row (f)2)
def test_mesh_oracle_against_analytic_depth():
    """oracle/mesh_ref.py (parity unpinned: no pyrender here): a sphere and a fronto-parallel quad against their
    analytic z-depth through the same camera model as K1 (pixel centres at +0.5, -z forward, y up)."""
    from oracle import mesh_ref as M
    H = W = 96
    fx = fy = 96.0
    cx = cy = 48.0
    c2w = np.eye(4)[:3]
    v, f = M.uv_sphere(1.0, 48, 96)
    # object and camera poses both go through the Blender -> OpenGL swap, so their relative geometry is the nerfstudio
    # one: with an identity c2w (looking down -z, y up) an object at world position p sits at camera-space p
    model = M.object_pose([0.1, -0.05, -1.5], [0, 0, 0], [0.05, 0.05, 0.05])     # radius 10 * 0.05 = 0.5
    cam = M.CONVERT @ np.eye(4)
    centre_cam = cam[:3, :3].T @ (model[:3, 3] - cam[:3, 3])
    assert np.allclose(centre_cam, [0.1, -0.05, -1.5])
    d = M.rasterize_depth(v, f, model, c2w, (fx, fy, cx, cy), H, W)
    ys, xs = np.mgrid[0:H, 0:W]
    dirs = np.stack([(xs + 0.5 - cx) / fx, -(ys + 0.5 - cy) / fy, -np.ones_like(xs, dtype=np.float64)], -1)
    a = (dirs ** 2).sum(-1)
    b = -2 * (dirs @ centre_cam)
    c = centre_cam @ centre_cam - 0.25
    disc = b * b - 4 * a * c
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0)
    hit = (disc > 0) & (t > 0)
    assert hit.sum() > 500 and ((d > 0) == hit).mean() > 0.995
    both = (d > 0) & hit
    err = np.abs(d - t)[both]                      # z-depth = t * |dir_z| = t
    assert np.median(err) < 2e-3 and err.max() < 2e-2   # 24-bit buffer with znear 1e-4: ~1e-3 at depth 1.5; facets at the rim
    # quad at camera-space depth 0.8 facing the camera: constant depth, exact coverage of its projected square
    z0 = 0.8
    q_cam = np.array([[-0.2, -0.2, -z0], [0.2, -0.2, -z0], [0.2, 0.2, -z0], [-0.2, 0.2, -z0]])
    q_gl = (cam[:3, :3] @ q_cam.T).T + cam[:3, 3]
    dq = M.rasterize_depth(q_gl.astype(np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.int32), np.eye(4), c2w, (fx, fy, cx, cy), H, W)
    assert (dq > 0).sum() == 48 * 48 and np.abs(dq[dq > 0] - z0).max() < 5e-4
    back = M.rasterize_depth(q_gl.astype(np.float32), np.array([[0, 2, 1], [0, 3, 2]], np.int32), np.eye(4), c2w, (fx, fy, cx, cy), H, W)
    assert (back > 0).sum() == 0                   # clockwise = back face, culled (pyrender's single-sided default)


def test_obj_parser_and_shape_condition():
    from oracle import mesh_ref as M
    v, f = M.parse_obj("# c\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n")
    assert v.shape == (4, 3) and f.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    proxy = np.zeros((40, 40), np.float32)
    proxy[10:20, 10:20] = 0.5
    nerf = np.full((40, 40), 0.7, np.float32)
    nerf[15:20, 10:20] = 0.3                      # NeRF surface in front of the proxy there
    mask, cond, vis = M.shape_mask_condition(proxy, nerf, dilation=(3, 3), radius=0.1)
    assert vis and mask[10:15, 10:20].all() and not mask[30, 30]
    mn, mx = np.float32(0.5) - np.float32(0.1), np.float32(0.5) + np.float32(0.1)
    assert np.isclose(cond[12, 12], 1 - (0.5 - mn) / (mx - mn)) and cond[30, 30] == 0.0     # far NeRF depth clamps to 0
    assert not M.shape_mask_condition(np.zeros_like(proxy), nerf)[2]
