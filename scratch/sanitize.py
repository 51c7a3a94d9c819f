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
# ---- kernels added late in round 1: coalesced GEMM epilogue modes, short-context attention, small-channel mma conv,
# VAE / inpaint operators, mesh rasteriser, hi / lo split producers
import numpy as np
from signerf_b200 import vae as V, inpaint as P, ops
from oracle import mesh_ref as M
o16 = K.gemm_f16(a, w, b, out_f16=True)
og = K.gemm_f16(a, w, b, geglu=True)
ors = K.gemm_f16(a, w, b, residual=r.clone(), out=r.clone())
qs = torch.randn(2 * 70, 128, generator=g).half().cuda(); kvs = torch.randn(2 * 5, 256, generator=g).half().cuda()
oxs = K.attention_f16(qs, kvs[:, :128], kvs[:, 128:], 2, 2)
xs = torch.randn(1, 21, 19, 16, generator=g).cuda(); ws = (torch.randn(32, 144, generator=g) * 0.1).half().cuda()
K.conv3x3_small_tc(xs, ws, torch.zeros(32).cuda(), stride=2, act_silu=True)
K.conv3x3_small_tc(xs, ws[:16].contiguous(), None)
xv = torch.randn(2 * 9 * 10, 8, generator=g).cuda()
V.im2col3x3_s2_asym_f16(xv, 2, 9, 10); V.im2col3x3_s2_asym_f16(xv, 2, 9, 10, split=True)
V.split_f16(xv); V.upsample2x_split_f16(xv, 2, 9, 10)
V.softmax_rows_f16(torch.randn(5, 128, generator=g).cuda(), 0.1)
K.group_norm_f16(xg, 2, 64, 32, 1e-6, torch.ones(128).cuda(), torch.zeros(128).cuda(), True, split=True)
m8 = (torch.rand(40, 56, generator=g) < 0.3).to(torch.uint8).cuda() * 255
bl = P.a1111_mask_blur(m8, 4)
P.pil_resize_bicubic_u8(bl, (5, 7)); P.overlay_mask_u8(bl)
gen8 = torch.randint(0, 256, (40, 56, 3), generator=g, dtype=torch.uint8).cuda()
P.overlay_composite(gen8, gen8.flip(0).contiguous(), P.overlay_mask_u8(bl))
mv, mf = M.uv_sphere(1.0, 6, 8)
c2w = torch.eye(4)[None, :3].clone(); c2w[0, 2, 3] = 3.0
ops.rasterize_depth(torch.from_numpy(mv).cuda(), torch.from_numpy(mf).cuda(), M.object_pose([0, 0, 0], [0, 0, 0], [0.05] * 3),
                    c2w.cuda(), torch.tensor([[32.0, 32.0, 16.0, 16.0]]).cuda(), 32, 32)
torch.cuda.synchronize()
print("sanitize pass 2 ok")
