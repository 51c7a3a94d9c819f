"""rows whose maximum keeps growing over the key tiles (the lazy-rescale / speculative fix-up paths), vs torch fp32"""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
B, heads, T = 1, 2, 2048
C = heads * 64
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn(B * T, C, device="cuda", generator=g)
k = torch.randn(B * T, C, device="cuda", generator=g)
v = torch.randn(B * T, C, device="cuda", generator=g)
ramp = (1.0 + 5.0 * torch.arange(T, device="cuda") / T)[:, None]      # later keys score higher
k = k * ramp
for gain in (1.0, 3.0):
    qh, kh, vh = (q * gain).half(), k.half(), v.half()
    qf = qh.float().view(B, T, heads, 64).transpose(1, 2); kf = kh.float().view(B, T, heads, 64).transpose(1, 2); vf = vh.float().view(B, T, heads, 64).transpose(1, 2)
    ref = torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, -1) @ vf
    ref = ref.transpose(1, 2).reshape(B * T, C)
    for var in [int(a) for a in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3")]:
        _lib.set_option("attn_variant", var)
        out = nn_ops.attention_f16(qh, kh, vh, B, heads)
        err = ((out.float() - ref).norm() / ref.norm()).item()
        print(f"gain {gain} variant {var}: rel-L2 {err:.2e} finite {bool(torch.isfinite(out).all())}")
