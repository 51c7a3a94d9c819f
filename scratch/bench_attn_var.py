import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
def timeit(fn, n=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(16))
shapes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
idles = [0]
for B, heads, T in ((2, 10, 16384), (2, 20, 4096)):
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B * T, C, device="cuda", generator=g).half() for _ in range(3))
    ref = None
    for var, shape in [(v_, s_) for s_ in shapes for v_ in variants]:
        for idle in idles:
            _lib.set_option("attn_variant", var); _lib.set_option("attn_idle_ns", idle); _lib.set_option("attn_shape", shape)
            out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
            ms = timeit(lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out))
            if ref is None: ref = out.clone()
            print(f"T{T} h{heads} shape {shape} variant {var}: {ms:.3f} ms {4*B*heads*T*T*64/ms/1e9:.0f} TF/s  maxdiff {(out.float()-ref.float()).abs().max().item():.2e}", flush=True)
