import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops as K
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
M, N, Kd = 8192, 1280, 1280
a = torch.randn(M, Kd, device="cuda").half(); w = torch.randn(N, Kd, device="cuda").half()
b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda")
for _ in range(3):
    flush.zero_()
    K.gemm_f16(a, w, b, residual=r, out=r)
    o = K.gemm_f16(a, w, None, out_f16=True)
torch.cuda.synchronize()
