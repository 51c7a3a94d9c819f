"""Whole nerfacto fine-tune step (train.NerfactoTrainer) at the reference's batch: 16 384 rays of 32 x 32 patches, 256 -> 96 -> 48
samples; per-phase CUDA-event times and the kernel list of one step."""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import ops, synthetic, train as T, _lib
dev = torch.device("cuda")
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=True)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
c2w, intr = synthetic.camera_ring(16, 512, 512)
o, d, _, _ = ops.generate_rays(c2w.to(dev), intr.to(dev), 512, 512)
samp = T.PatchPixelSampler(T.PatchPixelSamplerConfig(patch_size=32, num_rays_per_batch=N))
idx = samp.sample_method(samp.num_rays_per_batch, 16, 512, 512, device=dev)
cams = idx[:, 0].to(torch.int32).contiguous()
o, d = o[idx[:, 0], idx[:, 1], idx[:, 2]].contiguous(), d[idx[:, 0], idx[:, 1], idx[:, 2]].contiguous()
target = torch.rand(o.shape[0], 3, device=dev)
def random_pred_normals(seed=5):
    g = torch.Generator().manual_seed(seed)
    lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5, (torch.rand(o, generator=g) * 2 - 1) / i ** 0.5)
    p = {}
    for j, (o_, i_) in enumerate(((64, 27), (64, 64), (64, 64))):
        p[f"field.mlp_pred_normals.layers.{j}.weight"], p[f"field.mlp_pred_normals.layers.{j}.bias"] = lin(o_, i_)
    p["field.field_head_pred_normals.net.weight"], p["field.field_head_pred_normals.net.bias"] = lin(3, 64)
    return p
with_normals = len(sys.argv) > 2 and sys.argv[2] == "normals"
tr = T.NerfactoTrainer(fld, embedding=torch.randn(16, 32), pred_normals=random_pred_normals() if with_normals else None)
jit = torch.rand(3, o.shape[0], device=dev)
for _ in range(3):
    tr.train_step(o, d, target, jit, cams)
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
# phases
e = [ev() for _ in range(12)]
tr.zero_grad(); e[0].record()
smp = T.train_sample(fld, o, d, tr.counts, tr.near, tr.far, jit); e[1].record()
hb = T.appearance_bias(tr.w_app, tr.b_head0, tr.embedding, cams)
rgb, _, saved = T.train_forward(fld, o, d, smp.euclid[2], hb); e[2].record()
wf = T.weights_from_density(smp.euclid[2], saved[0])
l, g = T.rgb_loss(rgb, target)
inter = torch.zeros(1, device=dev); dl = torch.zeros(1, device=dev)
gp = [T.interlevel_loss(smp.spacing[2], wf, smp.spacing[k], smp.weights[k], inter) for k in range(2)]
gw = T.distortion_loss(smp.spacing[2], wf, dl); e[3].record()
ghb = torch.zeros_like(hb)
T.train_backward(fld, o, d, smp.euclid[2], saved, g, tr.grad_table, tr.grad_mlp, gw, hb, ghb); e[4].record()
T.appearance_bias_backward(tr.w_app, tr.embedding, cams, ghb, tr.grad_w_app, tr.grad_b_head0, tr.grad_embedding)
for k in range(2):
    T.prop_backward(fld, k, o, d, smp.euclid[k], smp.sigma[k], gp[k], tr.grad_prop_tables[k], tr.grad_prop_mlps[k])
e[5].record()
tr._per_image = True
tr.optimizer_step(); e[6].record()
torch.cuda.synchronize()
names = ["sample", "forward", "losses", "backward main", "backward prop+app", "adam"]
print(f"rays {o.shape[0]}: " + ", ".join(f"{n} {e[i].elapsed_time(e[i + 1]):.3f} ms" for i, n in enumerate(names)), f"| sum {e[0].elapsed_time(e[6]):.3f} ms")
t0 = time.perf_counter(); a, b = ev(), ev(); a.record()
for _ in range(20):
    tr.train_step(o, d, target, jit, cams)
b.record(); torch.cuda.synchronize()
print(f"predict_normals {with_normals}: 20 steps: {a.elapsed_time(b) / 20:.3f} ms / step on the device, {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms wall")
