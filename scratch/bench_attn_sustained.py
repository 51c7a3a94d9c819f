"""attention variants in SUSTAINED mode: each configuration loops ~2.5 s so that the GPU sits at its power cap (the state the
kernel runs in inside the benchmark step); TF/s over the last 1.5 s, SM clock and board power sampled from nvidia-smi"""
import subprocess, sys, time, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
def smi():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout
    return o.strip().replace("\n", " ")
cfgs = [(sh, var, sp) for sh in (0, 1) for var in (0, 3) for sp in (0, 1)]
for B, heads, T in ((2, 10, 16384), (2, 20, 4096)):
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B * T, C, device="cuda", generator=g).half() for _ in range(3))
    out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
    for sh, var, sp in cfgs:
        _lib.set_option("attn_shape", sh); _lib.set_option("attn_variant", var); _lib.set_option("attn_split", sp)
        fn = lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out)
        fn(); torch.cuda.synchronize()
        t0 = time.time()
        while time.time() - t0 < 1.0:
            for _ in range(20): fn()
            torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        n = 0; a.record(); t1 = time.time(); mid = None
        while time.time() - t1 < 1.5:
            for _ in range(20): fn()
            n += 20
            if mid is None and time.time() - t1 > 0.7: mid = smi()
            torch.cuda.synchronize()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"T{T} shape {sh} variant {var} split {sp}: {ms:.3f} ms {4*B*heads*T*T*64/ms/1e9:.0f} TF/s sustained  [sm MHz, W: {mid}]", flush=True)
