import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops as K
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (M, N, Kd) in ((8192, 1280, 1280), (8192, 1280, 5120), (32768, 640, 640), (32768, 640, 2560)):
    a = torch.randn(M, Kd, device="cuda").half(); w = torch.randn(N, Kd, device="cuda").half()
    b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda")
    K.gemm_f16(a, w, b, residual=r, out=r); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): K.gemm_f16(a, w, b, residual=r, out=r)
    ms = timeit(lambda: g.replay(), n=5) / 20
    print(f"residual in place fp32: M{M} N{N} K{Kd}: {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.0f} TF/s")
