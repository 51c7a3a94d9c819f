import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for B, heads, T in ((2, 20, 4096), (2, 10, 16384)):
    C = heads * 64
    q = torch.randn(B * T, C, device="cuda").half(); kv = torch.randn(B * 77, 2 * C, device="cuda").half()
    out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
    for short in (0, 1):
        _lib.set_option("attn_short_kv", short)
        ms = timeit(lambda: nn_ops.attention_f16(q, kv[:, :C], kv[:, C:], B, heads, out=out))
        print(f"cross T{T} h{heads} short_kv={short}: {ms*1e3:.1f} us", flush=True)
