"""SURVEY §8(f) row 4 — the fine-tune step: patch sampler (CPU), and on the GPU the main field's forward / BACKWARD, the
image loss and the Adam update against torch autograd through the oracle's restatement of nerfstudio's torch modules
(oracle/nerfacto_ref.py render_rays_train) on identical parameters.  Gradients carry fp32 atomics (summation order
varies run to run): tolerance 1e-3 relative L2 per parameter tensor, the north star's bar."""
import pytest
import torch

from oracle import nerfacto_ref as R
from signerf_b200 import train as T
from tests.helpers import rel_l2


# ---------------------------------------------------------------------------------------------- CPU: patch sampler
def test_patch_sampler_matches_reference_restatement():
    s = T.PatchPixelSampler(T.PatchPixelSamplerConfig(patch_size=8, num_rays_per_batch=1000))
    assert s.num_rays_per_batch == 960                                  # floored to whole patches (:36-42)
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    idx = s.sample_method(s.num_rays_per_batch, 7, 48, 64, generator=g1)
    ref = R.patch_sample_method(960, 7, 48, 64, 8, generator=g2)
    assert idx.dtype == torch.int64 and torch.equal(idx, ref)
    p = idx.view(15, 8, 8, 3)
    assert bool((p[..., 0] == p[:, :1, :1, 0]).all())                    # one image per patch
    assert bool((p[:, :, :, 1] - p[:, :1, :, 1] == torch.arange(8)[None, :, None]).all())       # rows, then columns
    assert bool((p[:, :, :, 2] - p[:, :, :1, 2] == torch.arange(8)[None, None, :]).all())
    assert int(p[..., 1].max()) < 48 and int(p[..., 2].max()) < 64 and int(p[..., 0].max()) < 7
    # with a mask the reference falls back to per-pixel sampling inside the mask
    mask = torch.zeros(7, 48, 64, 1, dtype=torch.bool)
    mask[2, 10:20, 5:9] = True
    m = s.sample_method(128, 7, 48, 64, mask=mask, generator=g1)
    assert m.shape == (128, 3) and bool(mask[m[:, 0], m[:, 1], m[:, 2], 0].all())


def test_mlp_block_layout_matches_the_library():
    from signerf_b200 import _lib
    n = int(_lib.load().sgn_mlp_param_count())
    v = T.mlp_block_views(torch.arange(n, dtype=torch.float32))
    assert v["w_base0"].shape == (64, 32) and v["w_head2"].shape == (3, 64) and v["tail"].shape == (4,)
    assert float(v["b_base0"][0]) == 64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 3 * 64


# ---------------------------------------------------------------------------------------------- GPU
def _setup(seed=0, n_rays=192, S=24, log2=14):
    from tests.helpers import field_from_oracle, ring_cameras
    m = R.make_model(seed, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=log2)
    m.train()
    fld = field_from_oracle(m, with_proposals=False)
    c2w, intr = ring_cameras(4, 32, 24)
    rays = [R.generate_rays(c2w[v], *intr[v].tolist(), 32, 24) for v in range(4)]
    g = torch.Generator().manual_seed(1)
    pick = torch.randperm(4 * 32 * 24, generator=g)[:n_rays]
    o = torch.cat([r.origins for r in rays])[pick].contiguous()
    d = torch.cat([r.directions for r in rays])[pick].contiguous()
    target = torch.rand(n_rays, 3, generator=g)
    bins = R.flat_bin_edges(S, m.near, m.far)
    return m, fld, o, d, target, bins, S


@pytest.mark.gpu
@pytest.mark.parametrize("use_l1", [True, False])
def test_forward_loss_and_gradients_match_autograd(use_l1):
    m, fld, o, d, target, bins, S = _setup()
    rgb_ref = R.render_rays_train(m, o, d, S)
    loss_ref = R.signerf_rgb_loss(rgb_ref, target, use_l1)
    loss_ref.backward()
    tr = T.FieldTrainer(fld, use_l1=use_l1)
    rgb, acc, saved = T.train_forward(fld, o.cuda(), d.cuda(), bins.cuda())
    assert rel_l2(rgb, rgb_ref) < 2e-5
    loss, grad = T.rgb_loss(rgb, target.cuda(), use_l1)
    assert abs(float(loss) - float(loss_ref.detach())) < 1e-5 * max(1.0, abs(float(loss_ref.detach())))
    tr.zero_grad()
    tr.backward(o.cuda(), d.cuda(), bins.cuda(), saved, grad)
    torch.cuda.synchronize()
    ours = T.nerfstudio_gradients(tr.grad_table, tr.grad_mlp, m.field.embedding_appearance.weight.detach().mean(dim=0))
    f = m.field
    ref = {"field.mlp_base.encoding.hash_table": f.encoding.hash_table.grad}
    for i, l in enumerate(f.mlp_base.layers):
        ref[f"field.mlp_base.mlp.layers.{i}.weight"], ref[f"field.mlp_base.mlp.layers.{i}.bias"] = l.weight.grad, l.bias.grad
    for i, l in enumerate(f.mlp_head.layers):
        ref[f"field.mlp_head.layers.{i}.weight"], ref[f"field.mlp_head.layers.{i}.bias"] = l.weight.grad, l.bias.grad
    errs = {k: rel_l2(ours[k], ref[k]) for k in ref}
    print("gradient rel-L2:", {k.replace("field.", ""): f"{v:.1e}" for k, v in errs.items()})
    assert float(ref["field.mlp_base.encoding.hash_table"].abs().max()) > 0
    assert max(errs.values()) < 1e-3, errs


@pytest.mark.gpu
def test_autograd_function_and_per_ray_bins():
    """render_rays_train composes with any torch loss on rgb (the LPIPS hook); per-ray bins [N,S+1] = shared bins repeated
    give the same gradients as the shared form."""
    m, fld, o, d, target, bins, S = _setup(n_rays=96, S=16)
    tr = T.FieldTrainer(fld)
    oc, dc, bc, tc = o.cuda(), d.cuda(), bins.cuda(), target.cuda()
    tr.zero_grad()
    rgb = T.render_rays_train(tr, oc, dc, bc)
    loss = ((rgb - tc) ** 2).mean() + 0.1 * rgb.abs().mean()
    loss.backward()
    g_shared = (tr.grad_table.clone(), tr.grad_mlp.clone())
    rgb_ref = R.render_rays_train(m, o, d, S)
    (((rgb_ref - target) ** 2).mean() + 0.1 * rgb_ref.abs().mean()).backward()
    assert rel_l2(g_shared[0], m.field.encoding.hash_table.grad) < 1e-3
    tr.zero_grad()
    rgb2 = T.render_rays_train(tr, oc, dc, bc[None].repeat(oc.shape[0], 1).contiguous())
    (((rgb2 - tc) ** 2).mean() + 0.1 * rgb2.abs().mean()).backward()
    assert torch.equal(rgb2.detach(), rgb.detach())
    assert rel_l2(tr.grad_table, g_shared[0]) < 1e-5 and rel_l2(tr.grad_mlp, g_shared[1]) < 1e-5


@pytest.mark.gpu
def test_adam_steps_match_torch_and_reduce_the_loss():
    """Three full fine-tune steps (forward, L1, backward, Adam lr 1e-2 / eps 1e-15 as signerf_config.py:43-50) against
    torch.optim.Adam on the oracle: parameters stay within 1e-3, the loss goes down, and the fp16 renderer sees the new
    weights after refresh_renderer()."""
    from signerf_b200 import ops
    m, fld, o, d, target, bins, S = _setup(n_rays=256, S=16)
    params = [m.field.encoding.hash_table] + [p for l in list(m.field.mlp_base.layers) + list(m.field.mlp_head.layers) for p in (l.weight, l.bias)]
    opt = torch.optim.Adam(params, lr=1e-2, eps=1e-15)
    tr = T.FieldTrainer(fld, lr=1e-2, eps=1e-15)
    oc, dc, bc, tc = o.cuda(), d.cuda(), bins.cuda(), target.cuda()
    before = ops.render_rays(fld, oc, dc, ops.RenderOptions(mode="flat", num_samples=S))[0].clone()
    losses, losses_ref = [], []
    for _ in range(3):
        opt.zero_grad()
        lr_ = R.signerf_rgb_loss(R.render_rays_train(m, o, d, S), target, True)
        lr_.backward()
        opt.step()
        losses_ref.append(float(lr_))
        losses.append(float(tr.step(oc, dc, bc, tc)))
    assert all(abs(a - b) < 2e-4 for a, b in zip(losses, losses_ref)), (losses, losses_ref)
    assert losses[-1] < losses[0]
    v = T.mlp_block_views(tr.mlp)
    h0 = m.field.mlp_head.layers[0]
    folded = h0.bias + h0.weight[:, 31:] @ m.field.embedding_appearance.weight.mean(dim=0)
    # Adam's first steps move every touched parameter by ~lr * sign(g): an entry whose gradient is a sum that cancels to
    # rounding can take the other sign (fp32 atomics vs torch's summation order) and then sits 2 * lr away.  Such entries
    # are counted; all the others must agree closely.
    def off(a, b, tol=2e-3):
        return float(((a.detach().cpu() - b.detach().cpu()).abs() > tol).float().mean())
    stats = {"table": off(tr.table, m.field.encoding.hash_table), "w_base0": off(v["w_base0"], m.field.mlp_base.layers[0].weight),
             "w_head1": off(v["w_head1"], m.field.mlp_head.layers[1].weight), "b_head0'": off(v["b_head0"], folded),
             "w_app": off(tr.w_app, h0.weight[:, 31:]), "b_head2": off(v["b_head2"][:3], m.field.mlp_head.layers[2].bias)}
    print("fraction of parameters more than 2e-3 from torch.optim.Adam after 3 steps:", {k: f"{x:.2%}" for k, x in stats.items()})
    assert max(stats.values()) < 0.01, stats
    moved = float((tr.table.cpu() - R.make_model(0, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=14)
                   .field.encoding.hash_table.detach()).abs().gt(1e-3).float().mean())
    assert moved > 0.05                                    # the batch really updated a good part of the table
    tr.refresh_renderer()
    after = ops.render_rays(fld, oc, dc, ops.RenderOptions(mode="flat", num_samples=S))[0]
    assert not torch.equal(after, before)
    with torch.no_grad():
        ref_after = R.render_rays_train(m, o, d, S).clamp(0, 1)
    assert rel_l2(after, ref_after) < 2e-3        # fp16 tensor-core renderer on the UPDATED weights vs the updated oracle
