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


# ---------------------------------------------------------------------------------------------- the whole nerfacto step
def _setup_full(n_rays=160, counts=(64, 32, 16), seed=3):
    from tests.helpers import field_from_oracle, ring_cameras
    m = R.make_model(seed, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=14,
                     num_proposal_samples=counts[:2], num_nerf_samples=counts[2])
    m.train()
    fld = field_from_oracle(m, with_proposals=True)
    c2w, intr = ring_cameras(4, 32, 24)
    rays = [R.generate_rays(c2w[v], *intr[v].tolist(), 32, 24) for v in range(4)]
    g = torch.Generator().manual_seed(seed)
    pick = torch.randperm(4 * 32 * 24, generator=g)[:n_rays]
    o = torch.cat([r.origins for r in rays])[pick].contiguous()
    d = torch.cat([r.directions for r in rays])[pick].contiguous()
    target = torch.rand(n_rays, 3, generator=g)
    jitter = torch.rand(3, n_rays, generator=g)
    cams = torch.randint(0, 30, (n_rays,), generator=g)
    return m, fld, o, d, target, jitter, cams


@pytest.mark.gpu
@pytest.mark.parametrize("jittered", [True, False])
def test_training_sampler_matches_oracle(jittered):
    """ProposalNetworkSampler while training (stratified initial bins, jittered PDF re-sampling on the same draws) and in
    its eval form: bin edges of all three levels and both proposal weight sets."""
    m, fld, o, d, _, jitter, _ = _setup_full(counts=(256, 96, 48))
    with torch.no_grad():
        ref = R.forward_train(m, o, d, jitter if jittered else None)
    smp = T.train_sample(fld, o.cuda(), d.cuda(), (256, 96, 48), m.near, m.far, jitter.cuda() if jittered else None)
    for l in range(3):
        sp_ref = R.sdist_of(ref["samples_list"][l])
        assert rel_l2(smp.spacing[l], sp_ref) < 1e-5, l
        assert float((smp.spacing[l].cpu() - sp_ref).abs().max()) < 2e-4, l
    assert torch.equal(smp.spacing[0].cpu(), R.sdist_of(ref["samples_list"][0]))        # level 0 is arithmetic on the draws only
    for l in range(2):
        assert rel_l2(smp.weights[l], ref["weights_list"][l][..., 0]) < 1e-4, l


@pytest.mark.gpu
@pytest.mark.parametrize("per_image", [True, False])
def test_full_step_losses_and_gradients_match_autograd(per_image):
    """get_outputs while training + get_loss_dict (signerf.py:41-68 without LPIPS / normals): rgb, interlevel and distortion
    losses and the gradient of their sum with respect to EVERY trained tensor - main field, both proposal networks, the
    appearance embedding - against torch autograd through the oracle on the same draws."""
    m, fld, o, d, target, jitter, cams = _setup_full()
    tr = T.NerfactoTrainer(fld, embedding=m.field.embedding_appearance.weight if per_image else None, counts=(64, 32, 16),
                           near=m.near, far=m.far)
    ours = tr.forward_backward(o.cuda(), d.cuda(), target.cuda(), jitter.cuda(), cams.cuda() if per_image else None)
    torch.cuda.synchronize()
    # The oracle differentiates on OUR bins (the samplers carry no gradient and are compared in the test above): an fp32
    # rounding difference of a bin edge would move samples across cells of the finest hash levels.
    smp = T.train_sample(fld, o.cuda(), d.cuda(), (64, 32, 16), m.near, m.far, jitter.cuda())
    fixed = [R.samples_from_edges(smp.spacing[l].cpu(), smp.euclid[l].cpu()) for l in range(3)]
    out = R.forward_train(m, o, d, jitter, cams if per_image else None, fixed_samples=fixed)
    ld = R.signerf_loss_dict(out, target)
    sum(ld.values()).backward()
    for k in ("rgb_loss", "interlevel_loss", "distortion_loss"):
        a, b = float(ours[k]), float(ld[k].detach())
        assert abs(a - b) < 1e-4 * max(abs(b), 1e-3), (k, a, b)
    f = m.field
    app_mean = f.embedding_appearance.weight.detach().mean(dim=0)
    g = T.nerfstudio_gradients(tr.grad_table, tr.grad_mlp, app_mean)
    ref = {"field.mlp_base.encoding.hash_table": f.encoding.hash_table.grad}
    for i, l in enumerate(f.mlp_base.layers):
        ref[f"field.mlp_base.mlp.layers.{i}.weight"], ref[f"field.mlp_base.mlp.layers.{i}.bias"] = l.weight.grad, l.bias.grad
    for i, l in enumerate(f.mlp_head.layers):
        ref[f"field.mlp_head.layers.{i}.weight"], ref[f"field.mlp_head.layers.{i}.bias"] = l.weight.grad, l.bias.grad
    if per_image:                     # the appearance columns / bias / table have their own buffers in this mode
        h0 = f.mlp_head.layers[0]
        g["field.mlp_head.layers.0.weight"] = torch.cat([g["field.mlp_head.layers.0.weight"][:, :31], tr.grad_w_app], dim=1)
        g["field.mlp_head.layers.0.bias"] = tr.grad_b_head0
        g["field.embedding_appearance.weight"] = tr.grad_embedding
        ref["field.embedding_appearance.weight"] = f.embedding_appearance.weight.grad
        assert float(h0.weight.grad[:, 31:].abs().max()) > 0
    errs = {k: rel_l2(g[k], ref[k]) for k in ref}
    for l, p in enumerate(m.proposal_networks):
        v = tr.grad_prop_mlps[l]
        errs[f"prop{l}.table"] = rel_l2(tr.grad_prop_tables[l], p.encoding.hash_table.grad)
        errs[f"prop{l}.w0"] = rel_l2(v[:160].view(16, 10), p.mlp.layers[0].weight.grad)
        errs[f"prop{l}.b0"] = rel_l2(v[160:176], p.mlp.layers[0].bias.grad)
        errs[f"prop{l}.w1"] = rel_l2(v[176:192].view(1, 16), p.mlp.layers[1].weight.grad)
        errs[f"prop{l}.b1"] = rel_l2(v[192:193], p.mlp.layers[1].bias.grad)
        assert float(p.encoding.hash_table.grad.abs().max()) > 0
    print("gradient rel-L2:", {k.replace("field.", ""): f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs


@pytest.mark.gpu
def test_full_training_steps_reduce_the_losses_and_move_every_group():
    """Five whole steps (sampler, three losses, both backward passes, Adam on fields + proposal_networks + embedding):
    the summed loss goes down, every parameter group moves, and the eval cascade renders with the updated networks."""
    from signerf_b200 import ops
    m, fld, o, d, target, jitter, cams = _setup_full(n_rays=256)
    tr = T.NerfactoTrainer(fld, embedding=m.field.embedding_appearance.weight, counts=(64, 32, 16), near=m.near, far=m.far)
    before = [t.clone() for t in (tr.table, tr.mlp, tr.prop_tables[0], tr.prop_mlps[1], tr.embedding, tr.w_app)]
    oc, dc, tc, jc, cc = o.cuda(), d.cuda(), target.cuda(), jitter.cuda(), cams.cuda()
    opts = ops.RenderOptions(mode="cascade", num_samples=16, num_prop_samples=(64, 32))
    rgb0 = ops.render_rays(fld, oc, dc, opts)[0].clone()
    totals, rows = [], []
    for _ in range(12):
        out = tr.train_step(oc, dc, tc, jc, cc)
        rows.append({k: round(float(v), 5) for k, v in out.items()})
        totals.append(sum(float(v) for v in out.values()))
    print("loss dicts:", rows[0], rows[-1], [round(t, 4) for t in totals])
    assert min(totals[-3:]) < totals[0], totals
    after = (tr.table, tr.mlp, tr.prop_tables[0], tr.prop_mlps[1], tr.embedding, tr.w_app)
    assert all(not torch.equal(a, b) for a, b in zip(after, before))
    tr.refresh_renderer()
    rgb1 = ops.render_rays(fld, oc, dc, opts)[0]
    assert not torch.equal(rgb0, rgb1) and bool(torch.isfinite(rgb1).all())


@pytest.mark.gpu
def test_training_entry_points_reject_bad_arguments_and_accept_empty_batches():
    """Error behaviour through the C ABI: a field without proposal networks cannot run the training sampler, workspaces
    that are too small are refused, empty batches are no-ops."""
    import ctypes as C
    from signerf_b200 import _lib
    from tests.helpers import field_from_oracle
    m = R.make_model(0, dense=True, log2_hashmap_size=12)
    bare = field_from_oracle(m, with_proposals=False)
    o = torch.zeros(4, 3, device="cuda")
    d = torch.tensor([[0.0, 0.0, 1.0]], device="cuda").repeat(4, 1)
    with pytest.raises(_lib.SgnError, match="2 proposal networks"):
        T.train_sample(bare, o, d, (16, 8, 4))
    with pytest.raises(ValueError):
        T.NerfactoTrainer(bare)
    full = field_from_oracle(m, with_proposals=True)
    lib = _lib.load()
    smp = T.train_sample(full, o, d, (16, 8, 4), m.near, m.far, torch.rand(3, 4, device="cuda"))
    assert all(bool(torch.isfinite(t).all()) for t in smp.spacing + smp.euclid + smp.sigma + smp.weights)
    assert all(bool((t[:, 1:] >= t[:, :-1]).all()) for t in smp.spacing)                 # bin edges ascend along every ray
    # empty batch: every entry point returns OK without touching memory
    e = torch.zeros(0, 3, device="cuda")
    empty = T.train_sample(full, e, e, (16, 8, 4))
    assert empty.spacing[0].shape == (0, 17)
    rgb, acc, saved = T.train_forward(full, e, e, torch.linspace(0.05, 10.0, 5, device="cuda"))
    assert rgb.shape == (0, 3)
    # too-small workspace
    ws = torch.empty(16, dtype=torch.uint8, device="cuda")
    gw = torch.zeros(4, 16, device="cuda")
    rc = lib.sgn_prop_backward(full.handle, 0, o.data_ptr(), d.data_ptr(), 4, 16, smp.euclid[0].data_ptr(), smp.sigma[0].data_ptr(),
                               gw.data_ptr(), gw.data_ptr(), gw.data_ptr(), ws.data_ptr(), 16, None)
    assert rc != 0 and b"workspace" in lib.sgn_last_error()
    rc = lib.sgn_prop_backward(full.handle, 5, o.data_ptr(), d.data_ptr(), 4, 16, smp.euclid[0].data_ptr(), smp.sigma[0].data_ptr(),
                               gw.data_ptr(), gw.data_ptr(), gw.data_ptr(), ws.data_ptr(), 16, None)
    assert rc != 0 and b"no such proposal network" in lib.sgn_last_error()


@pytest.mark.gpu
def test_trained_parameters_round_trip_to_nerfstudio_names():
    """After fine-tuning, `state_dict()` carries every trained tensor under nerfstudio's names: a renderer rebuilt from it
    renders what the trained field renders, and `load_into` writes the same values into a reference-style module."""
    import signerf_b200.plugin as P
    from signerf_b200 import ops
    m, fld, o, d, target, jitter, cams = _setup_full(n_rays=128)
    tr = T.NerfactoTrainer(fld, embedding=m.field.embedding_appearance.weight, counts=(64, 32, 16), near=m.near, far=m.far)
    for _ in range(3):
        tr.train_step(o.cuda(), d.cuda(), target.cuda(), jitter.cuda(), cams.cuda())
    tr.refresh_renderer()
    sd = tr.state_dict()
    assert sd["field.mlp_head.layers.0.weight"].shape == (64, 63) and sd["proposal_networks.1.mlp_base.mlp.layers.1.weight"].shape == (1, 16)
    opts = ops.RenderOptions(mode="cascade", num_samples=16, num_prop_samples=(64, 32), near_plane=m.near, far_plane=m.far)
    rebuilt = P.FusedNerfactoGraph.from_state_dict(sd, average_init_density=m.field.average_init_density, render_opts=opts)
    a = ops.render_rays(fld, o.cuda(), d.cuda(), opts)
    b = ops.render_rays(rebuilt.field, o.cuda(), d.cuda(), opts)
    assert rel_l2(a[0], b[0]) < 1e-5 and rel_l2(a[1], b[1]) < 1e-5
    from tests.test_plugin_parity import _ContractOnlyModel       # an nn.Module exposing the oracle's live parameters under nerfstudio's names
    model = _ContractOnlyModel(m)
    before = m.field.encoding.hash_table.detach().clone()
    written = tr.load_into(model)
    assert "field.mlp_base.encoding.hash_table" in written and "proposal_networks.0.mlp_base.encoding.hash_table" in written
    assert not torch.equal(before, m.field.encoding.hash_table) and torch.equal(m.field.encoding.hash_table.cpu(), tr.table.cpu())
    assert torch.equal(m.field.mlp_head.layers[0].weight[:, 31:].cpu(), tr.w_app.cpu())
    assert torch.equal(m.proposal_networks[1].mlp.layers[0].weight.detach().cpu().flatten(), tr.prop_mlps[1][:160].cpu())
    assert torch.equal(m.field.embedding_appearance.weight.detach().cpu(), tr.embedding.cpu())
    # ... and the oracle on the written-back weights renders what the fine-tuned field renders (eval cascade, fp16 MLP path)
    m.eval()
    with torch.no_grad():
        ref = R.render_rays(m, o, d, "cascade")
    assert rel_l2(a[0], ref["rgb"]) < 2e-3


@pytest.mark.gpu
def test_fused_training_step_behind_the_models_own_parameters_and_optimizer():
    """FusedTrainingStep: the CUDA step's gradients land in `.grad` of the MODEL's tensors (nerfstudio names, used in place),
    equal to torch autograd through the oracle, scale with the upstream gradient (GradScaler), and torch.optim.Adam on the
    model's parameters - the reference's own optimizer - is what updates the field the kernels read."""
    import copy
    from tests.test_plugin_parity import _ContractOnlyModel
    m_cpu = R.make_model(3, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=14,
                         num_proposal_samples=(64, 32), num_nerf_samples=16)
    m_cpu.train()
    m_gpu = copy.deepcopy(m_cpu).cuda()
    model = _ContractOnlyModel(m_gpu)
    step = T.FusedTrainingStep(model, counts=(64, 32, 16), near=m_cpu.near, far=m_cpu.far,
                               average_init_density=m_cpu.field.average_init_density)
    assert step.field.grid.table.data_ptr() == m_gpu.field.encoding.hash_table.data_ptr()          # in place, no copy
    _, _, o, d, target, jitter, cams = _setup_full(n_rays=128)
    oc, dc, tc, jc, cc = o.cuda(), d.cuda(), target.cuda(), jitter.cuda(), cams.cuda()
    ld = step.loss_dict(oc, dc, tc, cc, jc)
    (3.0 * sum(ld.values())).backward()                                 # an upstream factor, as a GradScaler applies
    smp = T.train_sample(step.field, oc, dc, (64, 32, 16), m_cpu.near, m_cpu.far, jc)
    fixed = [R.samples_from_edges(smp.spacing[l].cpu(), smp.euclid[l].cpu()) for l in range(3)]
    ref = R.signerf_loss_dict(R.forward_train(m_cpu, o, d, jitter, cams, fixed_samples=fixed), target)
    (3.0 * sum(ref.values())).backward()
    for k in ref:
        assert abs(float(ld[k]) - float(ref[k].detach())) < 1e-4 * max(abs(float(ref[k].detach())), 1e-3), k
    errs = {n: rel_l2(pg.grad, pc.grad) for (n, pg), (_, pc) in zip(m_gpu.named_parameters(), m_cpu.named_parameters())
            if pc.grad is not None}
    print("grad rel-L2 through FusedTrainingStep:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert len(errs) >= 19 and max(errs.values()) < 1e-3, errs
    # the model's own optimizer drives the training; the kernels see its updates (tables in place, MLPs re-uploaded)
    opt = torch.optim.Adam(m_gpu.parameters(), lr=1e-2, eps=1e-15)
    totals = []
    for _ in range(12):
        opt.zero_grad()
        ld = step.loss_dict(oc, dc, tc, cc, jc)
        total = sum(ld.values())
        total.backward()
        opt.step()
        totals.append(float(total))
    assert min(totals[-3:]) < totals[0], totals


@pytest.mark.gpu
def test_pipeline_get_train_loss_dict_routes_through_the_fused_step():
    """plugin.SIGNeRFPipeline.enable_fused_training(): `get_train_loss_dict(step)` keeps VanillaPipeline's contract
    (batches from datamanager.next_train, the reference's loss-dict keys, a `distortion` metric) and leaves the gradients
    in the model's parameters."""
    import copy
    import signerf_b200.plugin as P
    from tests.test_plugin_parity import _ContractOnlyModel
    m = R.make_model(3, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=14)
    m_gpu = copy.deepcopy(m).cuda().train()
    model = _ContractOnlyModel(m_gpu)
    model.config.num_proposal_samples_per_ray, model.config.num_nerf_samples_per_ray = (64, 32), 16
    _, _, o, d, target, _, cams = _setup_full(n_rays=64)

    class _Bundle:
        origins, directions, camera_indices = o.cuda(), d.cuda(), cams.cuda()[:, None]

    class _DM:
        def next_train(self, step):
            return _Bundle(), {"image": target}

    pipe = P.SIGNeRFPipeline.__new__(P.SIGNeRFPipeline)
    pipe.datamanager, pipe._model = _DM(), model
    pipe.enable_fused_training()
    assert pipe._fused_step.trainer.counts == (64, 32, 16)
    outputs, loss_dict, metrics = pipe.get_train_loss_dict(step=0)
    assert set(loss_dict) == {"rgb_loss", "interlevel_loss", "distortion_loss"} and "distortion" in metrics
    # use_lpips (the reference's default): the model's own LPIPS module joins as a torch term on 8 x 8 patches of the batch
    model.config.use_lpips, model.config.patch_size, model.config.lpips_loss_mult = True, 8, 0.5
    seen = {}

    def fake_lpips(a, b):
        seen["shapes"] = (tuple(a.shape), tuple(b.shape), float(a.min()) >= -1.0, float(b.max()) <= 1.0)
        return ((a - b) ** 2).mean()

    model.lpips = fake_lpips
    _, with_lpips, _ = pipe.get_train_loss_dict(step=1)
    assert set(with_lpips) == {"rgb_loss", "interlevel_loss", "distortion_loss", "lpips_loss"} and float(with_lpips["lpips_loss"]) > 0
    assert seen["shapes"] == ((1, 3, 8, 8), (1, 3, 8, 8), True, True)
    model.config.use_lpips = False
    sum(loss_dict.values()).backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for n, p in m_gpu.named_parameters() if "scalings" not in n)
    assert float(m_gpu.field.encoding.hash_table.grad.abs().max()) > 0 and float(m_gpu.proposal_networks[1].encoding.hash_table.grad.abs().max()) > 0


@pytest.mark.gpu
def test_predict_normals_losses_and_gradients_match_autograd():
    """predict_normals=True (signerf_config.py:33): analytic normals (closed-form gradient of the density logit through the
    hash grid and the base MLP, no graph), predicted normals, orientation / pred-normal losses (signerf.py:69-80) and the
    gradients they add - into the prediction MLP + head and, through the geo features, the base MLP and the hash table."""
    from tests.helpers import field_from_oracle, ring_cameras
    m = R.make_model(3, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=14, num_proposal_samples=(64, 32),
                     num_nerf_samples=16, predict_normals=True)
    m.train()
    fld = field_from_oracle(m, with_proposals=True)
    _, _, o, d, target, jitter, cams = _setup_full(n_rays=128)
    pn = {k: v for k, v in m.state_dict().items() if "pred_normals" in k}
    tr = T.NerfactoTrainer(fld, embedding=m.field.embedding_appearance.weight, counts=(64, 32, 16), near=m.near, far=m.far,
                           pred_normals=pn, pred_normal_loss_mult=1.0, orientation_loss_mult=1.0)     # weighted up: visible in the sums
    oc, dc = o.cuda(), d.cuda()
    ours = tr.forward_backward(oc, dc, target.cuda(), jitter.cuda(), cams.cuda())
    torch.cuda.synchronize()
    smp = T.train_sample(fld, oc, dc, (64, 32, 16), m.near, m.far, jitter.cuda())
    fixed = [R.samples_from_edges(smp.spacing[l].cpu(), smp.euclid[l].cpu()) for l in range(3)]
    out = R.forward_train(m, o, d, jitter, cams, fixed_samples=fixed)
    ld = R.signerf_loss_dict(out, target, orientation_loss_mult=1.0, pred_normal_loss_mult=1.0)
    sum(ld.values()).backward()
    normals, pred = T.normals_forward(fld, tr.pn, oc, dc, smp.euclid[2])
    # analytic normals are a normalised gradient: compare where the gradient is not degenerate
    cosn = (normals.cpu() * out["normals"]).sum(-1)
    print(f"analytic normals: median cos {float(cosn.median()):.6f}, fraction with cos < 0.999: {float((cosn < 0.999).float().mean()):.4f}; "
          f"pred normals rel-L2 {rel_l2(pred, out['pred_normals'].detach()):.1e}")
    assert float((cosn < 0.999).float().mean()) < 0.01 and rel_l2(pred, out["pred_normals"].detach()) < 1e-4
    for k in ("rgb_loss", "interlevel_loss", "distortion_loss", "orientation_loss", "pred_normal_loss"):
        a, b = float(ours[k]), float(ld[k].detach())
        assert abs(a - b) < 2e-3 * max(abs(b), 1e-3), (k, a, b)
    g = T.pn_block_views(tr.grad_pn)
    f = m.field
    errs = {"pn.w0": rel_l2(g["w0"], f.mlp_pred_normals.layers[0].weight.grad), "pn.b0": rel_l2(g["b0"], f.mlp_pred_normals.layers[0].bias.grad),
            "pn.w1": rel_l2(g["w1"], f.mlp_pred_normals.layers[1].weight.grad), "pn.w2": rel_l2(g["w2"], f.mlp_pred_normals.layers[2].weight.grad),
            "pn.b2": rel_l2(g["b2"], f.mlp_pred_normals.layers[2].bias.grad), "pn.wh": rel_l2(g["wh"], f.field_head_pred_normals.weight.grad),
            "pn.bh": rel_l2(g["bh"][:3], f.field_head_pred_normals.bias.grad),
            "table": rel_l2(tr.grad_table, f.encoding.hash_table.grad),
            "w_base1": rel_l2(T.mlp_block_views(tr.grad_mlp)["w_base1"], f.mlp_base.layers[1].weight.grad),
            "w_base0": rel_l2(T.mlp_block_views(tr.grad_mlp)["w_base0"], f.mlp_base.layers[0].weight.grad)}
    print("gradient rel-L2 with predict_normals:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert max(errs.values()) < 2e-3, errs
    # the three separate entry points (forward / losses / backward) give what the fused pass gives
    w_final = T.weights_from_density(smp.euclid[2], T.train_forward(fld, oc, dc, smp.euclid[2],
                                     T.appearance_bias(tr.w_app, tr.b_head0, tr.embedding, cams.cuda().int()))[2][0])
    l_or, l_pn = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    g_pred = T.normal_losses(w_final, normals, pred, dc, l_or, l_pn, 1.0, 1.0)
    gpn2 = torch.zeros_like(tr.grad_pn)
    T.normals_backward(fld, tr.pn, oc, dc, smp.euclid[2], g_pred, gpn2)
    assert rel_l2(gpn2, tr.grad_pn) < 1e-5 and abs(float(l_pn) - float(ours["pred_normal_loss"])) < 1e-5 * max(1.0, abs(float(l_pn)))
    assert abs(float(l_or) - float(ours["orientation_loss"])) < 1e-5 * max(1.0, abs(float(l_or)))
    before = tr.pn.clone()
    tr.optimizer_step()
    assert not torch.equal(before, tr.pn) and "field.field_head_pred_normals.net.weight" in tr.state_dict()


@pytest.mark.gpu
def test_fused_training_step_with_predict_normals_on_the_models_parameters():
    """The reference's default model (`predict_normals=True`): FusedTrainingStep picks up `field.mlp_pred_normals.*` /
    `field.field_head_pred_normals.net.*`, returns the five loss terms of signerf.py:62-80 and leaves gradients in those
    tensors as well."""
    import copy
    from tests.test_plugin_parity import _ContractOnlyModel
    m_cpu = R.make_model(3, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=14, num_proposal_samples=(64, 32),
                         num_nerf_samples=16, predict_normals=True)
    m_cpu.train()
    m_gpu = copy.deepcopy(m_cpu).cuda()

    class _Model(_ContractOnlyModel):
        def state_dict(self, *a, **k):
            sd = super().state_dict()
            for i, l in enumerate(self.field.mlp_pred_normals.layers):
                sd[f"field.mlp_pred_normals.layers.{i}.weight"], sd[f"field.mlp_pred_normals.layers.{i}.bias"] = l.weight, l.bias
            sd["field.field_head_pred_normals.net.weight"] = self.field.field_head_pred_normals.weight
            sd["field.field_head_pred_normals.net.bias"] = self.field.field_head_pred_normals.bias
            return sd

    step = T.FusedTrainingStep(_Model(m_gpu), counts=(64, 32, 16), near=m_cpu.near, far=m_cpu.far,
                               average_init_density=m_cpu.field.average_init_density)
    assert step.trainer.pn is not None
    _, _, o, d, target, jitter, cams = _setup_full(n_rays=128)
    oc, dc, jc = o.cuda(), d.cuda(), jitter.cuda()
    ld = step.loss_dict(oc, dc, target.cuda(), cams.cuda(), jc)
    assert set(ld) == {"rgb_loss", "interlevel_loss", "distortion_loss", "orientation_loss", "pred_normal_loss"}
    sum(ld.values()).backward()
    smp = T.train_sample(step.field, oc, dc, (64, 32, 16), m_cpu.near, m_cpu.far, jc)
    fixed = [R.samples_from_edges(smp.spacing[l].cpu(), smp.euclid[l].cpu()) for l in range(3)]
    ref = R.signerf_loss_dict(R.forward_train(m_cpu, o, d, jitter, cams, fixed_samples=fixed), target)
    sum(ref.values()).backward()
    for k in ref:
        assert abs(float(ld[k]) - float(ref[k].detach())) < 2e-3 * max(abs(float(ref[k].detach())), 1e-4), (k, float(ld[k]), float(ref[k]))
    errs = {n: rel_l2(pg.grad, pc.grad) for (n, pg), (_, pc) in zip(m_gpu.named_parameters(), m_cpu.named_parameters())
            if pc.grad is not None}
    assert "field.mlp_pred_normals.layers.1.weight" in errs and "field.field_head_pred_normals.weight" in errs
    assert max(errs.values()) < 1e-3, errs


def test_proposal_anneal_and_update_schedule_follow_nerfacto():
    """CPU: `proposal_anneal` = bias(step / 1000, 10) and the update schedule of ProposalNetworkSampler (steps < 10 always,
    then whenever steps_since_update exceeds lerp(step / 5000, 1, 5))."""
    assert T.proposal_anneal(0) == 0.0 and T.proposal_anneal(1000) == 1.0 and T.proposal_anneal(5000) == 1.0
    assert abs(T.proposal_anneal(100) - 10 * 0.1 / (9 * 0.1 + 1)) < 1e-12
    sched = T.ProposalUpdateSchedule()
    early = [sched(s) for s in range(12)]
    assert all(early[:10])                                   # step < 10
    late = T.ProposalUpdateSchedule()
    hits = [late(s) for s in range(6000, 6030)]
    assert sum(hits) == 5 and hits[5] and not hits[0]        # past the warm-up: every sixth step (since > 5)


@pytest.mark.gpu
def test_annealed_sampler_and_frozen_proposals_match_oracle():
    """anneal < 1 (the first 1 000 steps after the reference's step-count reset): weights ** anneal in front of both PDF
    re-samplings; update_proposals False: no gradient into the proposal networks, the main field's unchanged."""
    m, fld, o, d, target, jitter, cams = _setup_full()
    anneal = T.proposal_anneal(137)
    with torch.no_grad():
        ref = R.forward_train(m, o, d, jitter, anneal=anneal)
    smp = T.train_sample(fld, o.cuda(), d.cuda(), (64, 32, 16), m.near, m.far, jitter.cuda(), anneal)
    for l in range(3):          # powf (a few ulp) against torch.pow in front of the cdf: a little looser than the plain sampler's 1e-5
        assert rel_l2(smp.spacing[l], R.sdist_of(ref["samples_list"][l])) < 5e-5, l
    plain = T.train_sample(fld, o.cuda(), d.cuda(), (64, 32, 16), m.near, m.far, jitter.cuda())
    assert rel_l2(plain.spacing[2], smp.spacing[2]) > 1e-3                                   # the exponent really acts
    tr = T.NerfactoTrainer(fld, counts=(64, 32, 16), near=m.near, far=m.far)
    out = tr.forward_backward(o.cuda(), d.cuda(), target.cuda(), jitter.cuda(), anneal=anneal, update_proposals=False)
    assert float(out["interlevel_loss"]) > 0
    assert all(float(g.abs().max()) == 0.0 for g in tr.grad_prop_tables + tr.grad_prop_mlps) and float(tr.grad_table.abs().max()) > 0
