"""Small-shape pass over every tensor-core kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops as K
g = torch.Generator().manual_seed(0)
a = torch.randn(300, 192, generator=g).half().cuda(); w = torch.randn(320, 192, generator=g).half().cuda()
b = torch.randn(320, generator=g).cuda(); r = torch.randn(300, 320, generator=g).cuda()
o = K.gemm_f16(a, w, b, residual=r)
assert torch.allclose(o, a.float() @ w.float().t() + b + r, atol=1e-2)
a1 = torch.randn(100, 64, generator=g).half().cuda(); w1 = torch.randn(48, 64, generator=g).half().cuda()
o1 = K.gemm_f16(a1, w1, None, out_f16=True)            # single-CTA shape (one M tile)
assert torch.allclose(o1.float(), a1.float() @ w1.float().t(), atol=5e-2)
x = torch.randn(2, 12, 20, 64, generator=g).half().cuda(); wc = torch.randn(96, 576, generator=g).half().cuda() * 0.05
oc = K.conv3x3_f16(x, wc, None)
qkv = torch.randn(2 * 200, 3 * 128, generator=g).half().cuda()
oa = K.attention_f16(qkv[:, :128], qkv[:, 128:256], qkv[:, 256:], 2, 2)
q = torch.randn(2 * 300, 128, generator=g).half().cuda(); kv = torch.randn(2 * 77, 256, generator=g).half().cuda()
ob = K.attention_f16(q, kv[:, :128], kv[:, 128:], 2, 2)
xg = torch.randn(2 * 64, 128, generator=g).cuda()
K.group_norm_f16(xg, 2, 64, 32, 1e-5, torch.ones(128).cuda(), torch.zeros(128).cuda(), True)
K.layer_norm_f16(xg, torch.ones(128).cuda(), torch.zeros(128).cuda())
torch.cuda.synchronize()
print("sanitize pass ok", float(oc.abs().mean()), float(oa.float().abs().mean()), float(ob.float().abs().mean()))
