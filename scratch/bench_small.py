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
print("== direct conv (hint stack + conv_in)")
x = torch.rand(1, 3, 2048, 2048, device="cuda")
cin, nchw, H = 3, True, 2048
for (cout, stride) in ((16, 1), (16, 1), (32, 2), (32, 1), (96, 2), (96, 1), (256, 2)):
    w = torch.randn(cout, 3, 3, cin, device="cuda"); b = torch.randn(cout, device="cuda")
    last = cout == 256
    ms = timeit(lambda: K.conv3x3_direct(x, nchw, w, b, stride=stride, act_silu=True, out_f16=last))
    Ho = (H - 1) // stride + 1
    print(f"  {cin:3d}->{cout:3d} s{stride} @{H}: {ms:.3f} ms  {2*Ho*Ho*9*cin*cout/ms/1e9:.1f} TF/s  out {Ho*Ho*cout*4/1e6:.0f} MB")
    x = K.conv3x3_direct(x, nchw, w, b, stride=stride, act_silu=True, out_f16=last); nchw = False; cin = cout; H = Ho
xin = torch.randn(2, 4, 256, 256, device="cuda"); w = torch.randn(320, 3, 3, 4, device="cuda"); b = torch.randn(320, device="cuda")
print(f"  conv_in 4->320 @256 x2: {timeit(lambda: K.conv3x3_direct(xin, True, w, b)):.3f} ms")
print("== group norm")
for (B, HW, C) in ((2, 65536, 320), (2, 65536, 640), (2, 65536, 960), (2, 16384, 640), (2, 16384, 1280), (2, 16384, 1920), (2, 4096, 1280), (2, 4096, 2560)):
    t = torch.randn(B * HW, C, device="cuda"); g = torch.ones(C, device="cuda"); bb = torch.zeros(C, device="cuda")
    ms = timeit(lambda: K.group_norm_f16(t, B, HW, 32, 1e-5, g, bb, True))
    gb = t.numel() * (4 + 4 + 2) / 1e9
    print(f"  B{B} HW{HW} C{C}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s")
print("== layer norm")
for (M, C) in ((32768, 640), (8192, 1280)):
    t = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); bb = torch.zeros(C, device="cuda")
    ms = timeit(lambda: K.layer_norm_f16(t, g, bb), n=20)
    print(f"  M{M} C{C}: {ms*1e3:.1f} us  {t.numel()*6/ms/1e6:.0f} GB/s")
print("== tiny gemm latency")
for (M, N, Kd) in ((128, 256, 64), (154, 2560, 2048), (154, 1280, 2048), (8192, 1280, 1280), (8192, 1280, 64)):
    a = torch.randn(M, Kd, device="cuda").half(); w = torch.randn(N, Kd, device="cuda").half()
    o = torch.empty(M, N, device="cuda", dtype=torch.float16)
    g = torch.cuda.CUDAGraph()
    K.gemm_f16(a, w, None, out_f16=True, out=o); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(20): K.gemm_f16(a, w, None, out_f16=True, out=o)
    ms = timeit(lambda: g.replay(), n=5) / 20
    print(f"  M{M} N{N} K{Kd}: {ms*1e3:.1f} us per launch inside a graph")
