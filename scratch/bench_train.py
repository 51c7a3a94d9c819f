"""Fine-tune step timing at the reference's batch: train_num_rays_per_batch 16384 -> (16384 // 1024) * 1024 rays of 32 x 32 patches, 48 samples."""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import ops, synthetic, train as T
dev = torch.device("cuda")
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=False)
N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 48
c2w, intr = synthetic.camera_ring(16, 512, 512)
o, d, _, _ = ops.generate_rays(c2w.to(dev), intr.to(dev), 512, 512)
samp = T.PatchPixelSampler(T.PatchPixelSamplerConfig(patch_size=32, num_rays_per_batch=N))
idx = samp.sample_method(samp.num_rays_per_batch, 16, 512, 512, device=dev)
o, d = o[idx[:, 0], idx[:, 1], idx[:, 2]].contiguous(), d[idx[:, 0], idx[:, 1], idx[:, 2]].contiguous()
bins = ops.piecewise_bin_edges(S, 0.05, 1000.0).to(dev)
target = torch.rand(o.shape[0], 3, device=dev)
tr = T.FieldTrainer(fld)
for _ in range(3):
    loss = tr.step(o, d, bins, target)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
tr.zero_grad(); evs[0].record()
rgb, acc, saved = T.train_forward(fld, o, d, bins); evs[1].record()
l, g = T.rgb_loss(rgb, target); evs[2].record()
tr.backward(o, d, bins, saved, g); evs[3].record()
tr.optimizer_step(); evs[4].record()
torch.cuda.synchronize()
names = ["forward", "loss", "backward", "adam"]
print(f"rays {o.shape[0]} x {S} samples: " + ", ".join(f"{n} {evs[i].elapsed_time(evs[i + 1]):.3f} ms" for i, n in enumerate(names)),
      f"| total {evs[0].elapsed_time(evs[4]):.3f} ms, loss {float(l):.4f}")
