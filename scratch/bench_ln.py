import sys, os, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops as K
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for M, C in ((8192, 1280), (32768, 640)):
    x = torch.randn(M, C, device="cuda"); g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    ref = torch.nn.functional.layer_norm(x, (C,), g, b)
    out = K.layer_norm_f16(x, g, b)
    err = (out.float() - ref).abs().max().item()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10): K.layer_norm_f16(x, g, b)
    warm = timeit(lambda: gr.replay()) / 10
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s, e in ev:
        flush.zero_(); s.record(); K.layer_norm_f16(x, g, b); e.record()
    torch.cuda.synchronize()
    cold = min(s.elapsed_time(e) for s, e in ev)
    print(f"LN M{M} C{C} rows/warp={os.environ.get('SGN_LN_ROWS_PER_WARP','4')}: warm {warm*1e3:.1f} us cold {cold*1e3:.1f} us maxerr {err:.1e}", flush=True)
