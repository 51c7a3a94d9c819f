import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import ops, synthetic
dev = torch.device("cuda")
V = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=True)
c2w, intr = synthetic.camera_ring(16, 512, 512)
copts = ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96))
for _ in range(2):
    ops.render_views(fld, c2w[:V].to(dev), intr[:V].to(dev), 512, 512, copts)
torch.cuda.synchronize()
