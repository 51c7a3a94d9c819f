import sys, torch
sys.path.insert(0, "/root/repo")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import sdxl_ref as R
from signerf_b200 import unet as U
from tests.helpers import rel_l2
from tests.test_unet_parity import _inputs
cfg = R.tiny_config()
ref_unet, ref_ctrl = R.make_models(cfg, seed=0, device="cuda")
net = U.SDXLDenoiserB200(U.UNetConfig(**cfg.__dict__), ref_unet.state_dict(), ref_ctrl.state_dict(), "cuda")
x, t, ctx, y, hint = _inputs(cfg)
tr, t1, t2 = {}, {}, {}
with torch.no_grad():
    ref = ref_unet(x, t, ctx, y, taps=tr)
o1 = net.unet.forward(x, t, ctx, y, taps=t1)
o2 = net.unet.forward(x, t, ctx, y, taps=t2)
for k in tr:
    print(f"{k:20s} err {rel_l2(t1[k], tr[k]):.2e}  run-to-run {rel_l2(t1[k], t2[k]):.2e}  |ref| {float(tr[k].abs().mean()):.3f}")
print("out", rel_l2(o1, ref), rel_l2(o1, o2))
