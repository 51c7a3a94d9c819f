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
B, heads, T = 2, 10, 16384
C = heads * 64
q = torch.randn(B * T, C, device="cuda").half(); k = torch.randn(B * T, C, device="cuda").half(); v = torch.randn(B * T, C, device="cuda").half()
out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
ms = timeit(lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out), n=5)
print(f"T{T}: {ms:.3f} ms {4*B*heads*T*T*64/ms/1e9:.0f} TF/s")
