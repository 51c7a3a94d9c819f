"""timing probes of k_attention_tc: the same launch through libraries built with -DSGN_ATTN_PROBE=<mask> (results are wrong,
the instruction mix is what is measured): 1 = no scale / shift FFMA2, 2 = no row-sum FADD2, 4 = 1/8 of the max work"""
import subprocess, sys, time, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import _lib
which = sys.argv[1]
if which != "0":
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libsignerf_b200.so", f"libsignerf_b200_probe{which}.so")
from signerf_b200 import nn_ops
for B, heads, T in ((2, 10, 16384), (2, 20, 4096)):
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B * T, C, device="cuda", generator=g).half() for _ in range(3))
    out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
    fn = lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
        time.sleep(0.2)
    print(f"probe {which} T{T}: {best:.3f} ms {4*B*heads*T*T*64/best/1e9:.0f} TF/s (burst)", flush=True)
