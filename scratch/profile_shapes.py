import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, unet, _lib
for a in sys.argv[1:]:
    k, v = a.split("="); _lib.set_option(k, int(v))
nn_ops.PROFILE_SHAPES = True
b = unet.BenchUNet("cuda", (2048, 2048), use_graph=False)
b._run(); torch.cuda.synchronize()
prof = b.profile_eager()
tot = sum(v["ms"] for v in prof.values())
print(f"total {tot:.2f} ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / v["ms"] / 1e9 if v["flops"] else 0
    print(f"{v['ms']:8.3f} ms {v['calls']:4d} calls {v['ms']/v['calls']*1e3:8.1f} us/call {tf:7.0f} TF/s  {k}")
