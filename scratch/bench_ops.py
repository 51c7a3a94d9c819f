import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

print("== GEMM")
for (M, N, K) in [(32768, 640, 640), (32768, 1920, 640), (32768, 5120, 640), (32768, 640, 2560),
                  (8192, 1280, 1280), (8192, 3840, 1280), (8192, 10240, 1280), (8192, 1280, 5120), (131072, 320, 320)]:
    a = torch.randn(M, K, device="cuda").half(); w = torch.randn(N, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ms = timeit(lambda: nn_ops.gemm_f16(a, w, None, out_f16=True, out=out))
    ms_t = timeit(lambda: torch.matmul(a, w.t()))
    print(f"M{M} N{N} K{K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF/s | cublas {ms_t:.3f} ms {2*M*N*K/ms_t/1e9:.0f} TF/s")
print("== conv3x3")
for (B, H, W, C, N) in [(2, 256, 256, 320, 320), (2, 128, 128, 640, 640), (2, 64, 64, 1280, 1280), (2, 64, 64, 2560, 1280), (2, 128, 128, 1920, 640)]:
    x = torch.randn(B, H, W, C, device="cuda").half(); w = torch.randn(N, 9 * C, device="cuda").half()
    out = torch.empty(B * H * W, N, device="cuda")
    ms = timeit(lambda: nn_ops.conv3x3_f16(x, w, None, out=out))
    print(f"B{B} {H}x{W} C{C}->N{N}: {ms:.3f} ms {2*B*H*W*N*9*C/ms/1e9:.0f} TF/s")
print("== attention")
for (B, heads, T, Tkv) in [(2, 10, 16384, 16384), (2, 20, 4096, 4096), (2, 10, 16384, 77), (2, 20, 4096, 77)]:
    C = heads * 64
    q = torch.randn(B * T, C, device="cuda").half(); k = torch.randn(B * Tkv, C, device="cuda").half(); v = torch.randn(B * Tkv, C, device="cuda").half()
    out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
    ms = timeit(lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out), n=5)
    print(f"B{B} h{heads} T{T} Tkv{Tkv}: {ms:.3f} ms {4*B*heads*T*Tkv*64/ms/1e9:.0f} TF/s")
