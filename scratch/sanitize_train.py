"""Small-shape pass over the round-2 kernels for compute-sanitizer (memcheck / racecheck): the training step (sampler,
losses, main + proposal backward with the warp-merged scatter-add, appearance embedding, Adam), the cascade renderer, the
split attention and the prompt encoders' entry points."""
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import nerfacto_ref as R
from signerf_b200 import nn_ops as K, ops, train as T
from tests.helpers import field_from_oracle, ring_cameras
m = R.make_model(3, dense=True, table_scale=0.5, density_gain=20.0, log2_hashmap_size=12, num_proposal_samples=(24, 12), num_nerf_samples=8,
                 proposal_net_args=({"hidden_dim": 16, "log2_hashmap_size": 12, "num_levels": 5, "max_res": 128},
                                    {"hidden_dim": 16, "log2_hashmap_size": 12, "num_levels": 5, "max_res": 256}))
fld = field_from_oracle(m, with_proposals=True)
c2w, intr = ring_cameras(2, 16, 12)
o, d, _, _ = ops.generate_rays(c2w.cuda(), intr.cuda(), 12, 16)
o, d = o.reshape(-1, 3)[:77].contiguous(), d.reshape(-1, 3)[:77].contiguous()      # ragged: not a multiple of the warp
g = torch.Generator().manual_seed(0)
from signerf_b200 import synthetic
tr = T.NerfactoTrainer(fld, embedding=torch.randn(5, 32, generator=g), counts=(24, 12, 8), near=m.near, far=m.far,
                       pred_normals=synthetic.random_pred_normals())
cams = torch.randint(0, 5, (77,), generator=g).cuda()
for _ in range(2):
    out = tr.train_step(o, d, torch.rand(77, 3, generator=g).cuda(), torch.rand(3, 77, generator=g).cuda(), cams)
tr.refresh_renderer()
ft = T.FieldTrainer(fld)
ft.step(o, d, ops.piecewise_bin_edges(8, m.near, m.far).cuda(), torch.rand(77, 3, generator=g).cuda())
rgb, depth = ops.render_views(fld, c2w.cuda(), intr.cuda(), 12, 16, ops.RenderOptions(mode="cascade", num_samples=8, num_prop_samples=(24, 12)))
q = torch.randn(2 * 700, 128, generator=g).half().cuda()
oa = K.attention_f16(q, q, q, 2, 2)                                               # 6 query tiles per head: tail items split over keys
torch.cuda.synchronize()
print("sanitize pass ok", {k: float(v) for k, v in out.items()}, float(rgb.mean()), float(oa.float().abs().mean()))
