"""GPU parity of K2-K4 (mask / dilation / condition / sheet assembly / quantisation) against the oracle and
the committed reference fixtures, through the C ABI.  Integer / boolean outputs are bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import nerfacto_ref as R
from oracle import sheet_ref as S
from signerf_b200 import ops
from tests.helpers import rel_l2, ring_cameras

pytestmark = pytest.mark.gpu


def test_dilate_matches_cv2_fixture(golden_dir):
    d = np.load(os.path.join(golden_dir, "dilate.npz"))
    m = torch.tensor(d["masks"]).cuda()
    assert np.array_equal(ops.dilate_ellipse(m, (50, 50)).cpu().numpy().astype(bool), d["out50"])
    assert np.array_equal(ops.dilate_ellipse(m, (7, 11)).cpu().numpy().astype(bool), d["out7x11"])
    # 1x1 kernel is the identity; empty batch is a no-op
    assert torch.equal(ops.dilate_ellipse(m, (1, 1)), m)
    assert ops.dilate_ellipse(m[:0], (50, 50)).shape[0] == 0


def test_dilate_full_size_idempotence_property():
    """512x512 x 16 views: dilating with the 1x1 kernel is identity, dilation is extensive (out >= in) and
    monotone, and agrees with the numpy restatement on one full-size view."""
    g = torch.Generator().manual_seed(0)
    m = (torch.rand(16, 512, 512, generator=g) < 1e-4).to(torch.uint8).cuda()
    out = ops.dilate_ellipse(m, (50, 50))
    assert bool((out >= m).all())
    assert np.array_equal(out[3].cpu().numpy().astype(bool), S.dilate(m[3].cpu().numpy(), (50, 50)))


def test_quantize_matches_tensor_to_image_fixture(golden_dir):
    q = np.load(os.path.join(golden_dir, "quantize.npz"))
    assert np.array_equal(ops.quantize_u8(torch.tensor(q["x"]).cuda()).cpu().numpy(), q["q"])


@pytest.mark.parametrize("inverse,dilation,manual", [(False, (50, 50), None), (False, None, None),
                                                       (True, (7, 11), None), (False, (50, 50), (0.2, 0.9))])
def test_mask_condition_matches_oracle(inverse, dilation, manual):
    H, W, V = 96, 128, 4
    c2w, intr = ring_cameras(V, W, H)
    g = torch.Generator().manual_seed(5)
    depth = 0.3 + 0.4 * torch.rand(V, H, W, 1, generator=g)
    depth[3] = 5.0  # beyond the box everywhere -> not visible (unless inverted)
    aabb = torch.tensor([[-0.1, -0.1, -0.1], [0.1, 0.1, 0.1]])
    opts = ops.MaskOptions(aabb=aabb.flatten().tolist(), inverse_mask=inverse, mask_dilation=dilation, manual_depth=manual)
    mask, cond, stats = ops.mask_condition(c2w.cuda(), intr.cuda(), depth.cuda(), opts)
    for v in range(V):
        rays = R.generate_rays(c2w[v], *intr[v].tolist(), W, H)
        m_ref, c_ref, st = S.render_camera_aabb(rays.origins.view(H, W, 3), rays.directions.view(H, W, 3), depth[v], aabb,
                                                inverse, dilation, 0.1, manual)
        assert bool(stats[v, 0] > 0) == st["is_visible"]
        assert int(stats[v, 3]) == st["count"]
        assert np.array_equal(mask[v].cpu().numpy().astype(bool), m_ref.numpy()), f"mask differs for view {v}"
        assert torch.allclose(cond[v].cpu(), c_ref, atol=1e-6)
    if not inverse:
        assert float(stats[3, 0]) == 0.0 and float(mask[3].sum()) == 0 and float(cond[3].abs().sum()) == 0


@pytest.mark.parametrize("rows,cols,th,tw,border", [(4, 4, 32, 32, 0), (2, 3, 24, 20, 5)])
def test_sheet_paste_cut_blend_match_oracle(rows, cols, th, tw, border):
    H, W = 2 * th, 2 * tw
    n = rows * cols - 1
    g = torch.Generator().manual_seed(9)
    renders = torch.rand(n, H, W, 3, generator=g)
    masks = torch.rand(n, H, W, 1, generator=g) > 0.6
    conds = torch.rand(n, H, W, 1, generator=g)
    img_ref, mask_ref, cond_ref = S.reference_sheet(list(renders), list(masks), list(conds), rows, cols, th, tw, border)
    lay = ops.SheetLayout(rows, cols, th, tw, border)
    assert (lay.height, lay.width) == S.sheet_size(rows, cols, th, tw, border)
    img = torch.ones(lay.height, lay.width, 3, device="cuda")
    msk = torch.zeros(lay.height, lay.width, 1, device="cuda")
    cnd = torch.zeros(lay.height, lay.width, 1, device="cuda")
    ops.sheet_paste(renders.cuda(), img, lay)
    ops.sheet_paste(masks.cuda(), msk, lay, threshold=0.5)
    ops.sheet_paste(conds.cuda(), cnd, lay)
    assert torch.allclose(img.cpu(), img_ref, atol=1e-6)
    assert torch.equal(msk.cpu(), mask_ref)
    assert torch.allclose(cnd.cpu(), cond_ref, atol=1e-6)
    # blend + cut back out
    edited = torch.rand(lay.height, lay.width, 3, generator=g)
    b_ref = S.blend(edited, img_ref, mask_ref)
    b = ops.blend_masked(edited.cuda(), img_ref.cuda(), mask_ref.cuda())  # same inputs as the oracle: bit-exact
    assert torch.equal(b.cpu(), b_ref)
    for cell in (0, n - 1):
        t_ref = S.cut_tile(b_ref, cell, cols, th, tw, border, H, W)
        t = ops.sheet_cut(b, lay, cell, H, W)
        assert torch.allclose(t.cpu(), t_ref, atol=1e-6)


def test_resize_matches_interpolate_fixture(golden_dir):
    r = np.load(os.path.join(golden_dir, "resize.npz"))
    img = torch.tensor(r["img"]).cuda()[None]
    for key, (h, w) in (("down", (10, 14)), ("odd", (13, 9)), ("up", (40, 56)), ("same", (20, 28))):
        lay = ops.SheetLayout(1, 1, h, w, 0)
        sheet = torch.zeros(lay.height, lay.width, 3, device="cuda")
        ops.sheet_paste(img, sheet, lay)
        assert np.allclose(sheet[:h, :w].cpu().numpy(), r[key], atol=1e-6), key
    # identity size is a bit-exact copy (BASELINE C2: downscale 1)
    lay = ops.SheetLayout(1, 1, 20, 28, 0)
    sheet = torch.zeros(lay.height, lay.width, 3, device="cuda")
    ops.sheet_paste(img, sheet, lay)
    assert np.array_equal(sheet[:20, :28].cpu().numpy(), r["img"])
