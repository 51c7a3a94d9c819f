"""clock64 trace of CTA 0 of k_attention_tc (make -C signerf_b200/csrc trace): per softmax group (first warp, lane 0) and the
issuer.  usage: attn_trace.py <attn_variant flags> [T heads]"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import _lib
_lib.LIB_PATH = "/root/repo/signerf_b200/libsignerf_b200_trace.so"
from signerf_b200 import nn_ops
var = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.set_option("attn_variant", var)
shape = int(sys.argv[4]) if len(sys.argv) > 4 else 1
_lib.set_option("attn_shape", shape)
kvt = 96 if shape == 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 10
B = 2
Cc = heads * 64
q, k, v = (torch.randn(B * T, Cc, device="cuda").half() for _ in range(3))
for _ in range(3): nn_ops.attention_f16(q, k, v, B, heads)
torch.cuda.synchronize()
buf = (C.c_longlong * (32 * 256))()
lib = _lib.load(); lib.sgn_debug_attn_trace.argtypes = [C.c_void_p]; lib.sgn_debug_attn_trace(buf)
t = np.array(buf).reshape(32, 256).astype(np.float64)
n = min(T // kvt, 256)
lo, hi = n // 6, n - n // 6 - 1
sl, sl1 = slice(lo, hi), slice(lo + 1, hi + 1)
print(f"shape {shape} variant {var} T {T}: period A {np.mean(np.diff(t[0, sl])):.0f}  B {np.mean(np.diff(t[16, sl])):.0f}")
for x, g in ((0, "A"), (1, "B")):
    o = 16 * x
    d = lambda a, b: np.mean(t[o + a, sl] - t[o + b, sl])
    has_o = t[o + 6, lo] > 0
    print(f"  group {g}: wait s_full {d(1, 0):.0f} | ld {d(2, 1):.0f} | max/resc {d(3, 2):.0f} | turn wait {d(4, 3):.0f} | "
          + (f"exp to first st {d(5, 4):.0f} | wait o_full {d(6, 5):.0f} | rest+st {d(7, 6):.0f}" if has_o else f"exp all {d(7, 4):.0f}")
          + f" | turn held {d(7, 4):.0f}")
    print(f"           QK(j+1) issue after s_free(j): {np.mean(t[o + 8, sl1] - t[o + 2, sl]):.0f}   s_full(j+1) seen after QK issue: {np.mean(t[o + 1, sl1] - t[o + 8, sl1]):.0f}"
          f"   PV(j) issue after p_full(j): {np.mean(t[o + 9, sl] - t[o + 7, sl]):.0f}")
if shape == 0: print(f"  A turn-got -> B turn-got {np.mean(t[20, sl] - t[4, sl]):.0f}; B turn-got -> A next {np.mean(t[4, sl1] - t[20, sl]):.0f}")
if shape == 1:
    base = t[0, lo]
    for j in range(lo, lo + 4):
        print(f"  tile {j}: " + " | ".join(f"g{x}: top {t[16*x+0, j]-base:.0f} s_full {t[16*x+1, j]-base:.0f} ld {t[16*x+2, j]-base:.0f} p_full {t[16*x+7, j]-base:.0f} PVissue {t[16*x+9, j]-base:.0f} issued {t[16*x+10, j]-base:.0f}" for x in (0, 1)))
