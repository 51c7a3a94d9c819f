"""C2' cascade (256 -> 96 -> 48) at benchmark size: 16 views 512^2, time + checksum of the outputs"""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libsignerf_b200.so", sys.argv[1])
from signerf_b200 import ops, synthetic
dev = torch.device("cuda")
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=True)
c2w, intr = synthetic.camera_ring(16, 512, 512)
copts = ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96))
for _ in range(2):
    rgb, depth = ops.render_views(fld, c2w.to(dev), intr.to(dev), 512, 512, copts)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    rgb, depth = ops.render_views(fld, c2w.to(dev), intr.to(dev), 512, 512, copts)
b.record(); torch.cuda.synchronize()
cs = int(rgb.view(torch.int32).to(torch.int64).sum()) ^ int(depth.view(torch.int32).to(torch.int64).sum())
print(f"cascade 16 x 512^2: {a.elapsed_time(b) / 3:.1f} ms, checksum {cs}")
# ragged image (tiles overhang the edge) + explicit ray bundle
r2, d2 = ops.render_views(fld, c2w[:2].to(dev), intr[:2].to(dev), 37, 45, copts)
print("ragged checksum", int(r2.view(torch.int32).to(torch.int64).sum()) ^ int(d2.view(torch.int32).to(torch.int64).sum()))
