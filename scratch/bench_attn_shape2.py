"""attn_shape 2 (one softmax group per CTA, two CTAs per SM) against the default: correctness vs torch fp32, burst and sustained TF/s"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
def ref(q, k, v, B, heads):
    T, C = q.shape[0] // B, q.shape[1]
    f = lambda t: t.float().view(B, T, heads, 64).transpose(1, 2)
    return torch.nn.functional.scaled_dot_product_attention(f(q), f(k), f(v)).transpose(1, 2).reshape(B * T, C)
g = torch.Generator(device="cuda").manual_seed(0)
for sh in (0, 2):
    _lib.set_option("attn_shape", sh)
    for B, heads, T in ((1, 3, 1000), (2, 5, 2304), (1, 2, 777)):
        q, k, v = (torch.randn(B * T, heads * 64, device="cuda", generator=g).half() for _ in range(3))
        o = nn_ops.attention_f16(q, k, v, B, heads)
        r = ref(q, k, v, B, heads)
        print(f"shape {sh} B{B} h{heads} T{T}: rel-L2 {float((o.float() - r).norm() / r.norm()):.2e}", flush=True)
for B, heads, T in ((2, 10, 16384), (2, 20, 4096)):
    C = heads * 64
    q, k, v = (torch.randn(B * T, C, device="cuda", generator=g).half() for _ in range(3))
    out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
    for sh in (0, 2):
        for var in (3, 2, 1, 0):
            _lib.set_option("attn_shape", sh); _lib.set_option("attn_variant", var)
            fn = lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out)
            for _ in range(5): fn()
            torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): fn()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print(f"T{T} shape {sh} variant {var}: {ms:.3f} ms {4*B*heads*T*T*64/ms/1e9:.0f} TF/s", flush=True)
            time.sleep(0.3)
