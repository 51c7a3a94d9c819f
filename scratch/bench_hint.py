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
x = torch.rand(1, 3, 2048, 2048, device="cuda")
cin, nchw, H = 3, True, 2048
for (cout, stride) in ((16, 1), (16, 1), (32, 2), (32, 1), (96, 2), (96, 1), (256, 2)):
    w = torch.randn(cout, 3, 3, cin, device="cuda"); b = torch.randn(cout, device="cuda")
    kp = (9 * cin + 7) // 8 * 8
    wk = torch.randn(cout, 2 * kp, device="cuda").half()
    t_dir = timeit(lambda: K.conv3x3_direct(x, nchw, w, b, stride=stride, act_silu=True))
    t_col = timeit(lambda: K.im2col3x3_split_f16(x, nchw, stride))
    col, ho, wo, _ = K.im2col3x3_split_f16(x, nchw, stride)
    t_gemm = timeit(lambda: K.gemm_f16(col, wk, b, act_silu=True))
    print(f"{cin:3d}->{cout:3d} s{stride} @{H}: direct {t_dir:.3f} ms | im2col {t_col:.3f} + gemm {t_gemm:.3f} = {t_col+t_gemm:.3f} ms  (col {col.numel()*2/1e6:.0f} MB)")
    x = K.conv3x3_direct(x, nchw, w, b, stride=stride, act_silu=True); nchw = False; cin = cout; H = ho
    del col
