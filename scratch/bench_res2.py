import sys, os, torch
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
shapes = ((8192, 1280, 1280), (32768, 640, 640))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, N, Kd) in shapes:
    a = torch.randn(M, Kd, device="cuda").half(); w = torch.randn(N, Kd, device="cuda").half()
    b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda")
    for mode in ("res", "f32", "f16"):
        for bn in [int(x) for x in os.environ.get("BNS", "0").split(",")]:
            if bn: os.environ["SGN_GEMM_BN"] = str(bn)
            else: os.environ.pop("SGN_GEMM_BN", None)
            if mode == "res": fn = lambda: K.gemm_f16(a, w, b, residual=r, out=r)
            elif mode == "f32": fn = lambda: K.gemm_f16(a, w, b, out=r)
            else:
                o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
                fn = lambda: K.gemm_f16(a, w, b, out_f16=True, out=o16)
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10): fn()
            ms = timeit(lambda: g.replay(), n=5) / 10
            # cold: L2 flushed before each call
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for s, e in ev:
                flush.zero_(); s.record(); fn(); e.record()
            torch.cuda.synchronize()
            cold = min(s.elapsed_time(e) for s, e in ev)
            print(f"{mode} M{M} N{N} K{Kd} bn={bn}: warm {ms*1e3:.1f} us {2*M*N*Kd/ms/1e9:.0f} TF/s | cold {cold*1e3:.1f} us", flush=True)
