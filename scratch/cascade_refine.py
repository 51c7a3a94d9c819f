"""cost / effect of the fp32 re-march of near-tie median depths in cascade mode"""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import _lib, ops, synthetic
dev = torch.device("cuda")
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=True)
c2w, intr = synthetic.camera_ring(16, 512, 512)
copts = ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96))
f32 = ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96), mlp_mode=ops.MLP_FP32)
cw, it = c2w[:4].to(dev), intr[:4].to(dev)
ref = ops.render_views(fld, cw, it, 512, 512, f32)[1].clone()
def run(tag):
    for _ in range(2): d = ops.render_views(fld, cw, it, 512, 512, copts)[1]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): d = ops.render_views(fld, cw, it, 512, 512, copts)[1]
    b.record(); torch.cuda.synchronize()
    diff = (d - ref).abs() / ref.abs().clamp_min(1e-6)
    print(f"{tag}: {a.elapsed_time(b)/5:.2f} ms / 4 views; depth vs fp32 path: rel-L2 {float((d-ref).norm()/ref.norm()):.2e}, rays off by > 1e-6: {float((diff > 1e-6).float().mean()):.4%}, "
          f"> 1e-3: {float((diff > 1e-3).float().mean()):.4%}, max rel {float(diff.max()):.2e}", flush=True)
run("refine on, delta 1e-3")
_lib.set_option("render_refine_delta_ppm", 100); run("refine on, delta 1e-4")
_lib.set_option("render_refine_depth", 0); run("refine off")
