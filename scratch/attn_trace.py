import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import _lib
_lib.LIB_PATH = "/root/repo/signerf_b200/libsignerf_b200_trace.so"
from signerf_b200 import nn_ops
_lib.set_option("attn_variant", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
B, heads, T = 2, 10, 16384
Cc = heads * 64
q, k, v = (torch.randn(B * T, Cc, device="cuda").half() for _ in range(3))
for _ in range(3): nn_ops.attention_f16(q, k, v, B, heads)
torch.cuda.synchronize()
buf = (C.c_longlong * (16 * 256))()
lib = _lib.load(); lib.sgn_debug_attn_trace.argtypes = [C.c_void_p]; lib.sgn_debug_attn_trace(buf)
t = np.array(buf).reshape(16, 256)
names = ["top", "s_full ok", "ld done+s_free", "pre-turn", "turn got", "pre o_full", "o_full ok", "p_full arrived", "issue QK", "issue PV"]
base = t[0, 40]
for j in range(40, 46):
    print(f"tile {j}: " + "  ".join(f"{names[e]}={t[e, j] - base}" for e in range(10)))
d = lambda a, b: np.mean(t[a, 20:120] - t[b, 20:120])
print("mean: period", np.mean(np.diff(t[0, 20:120])), "wait s_full", d(1, 0), "ld", d(2, 1), "max..", d(3, 2), "turn wait", d(4, 3), "exp half", d(5, 4), "wait o_full", d(6, 5), "exp2+st", d(7, 6))
print("QK(j+1) issue after s_free(j):", np.mean(t[8, 21:121] - t[2, 20:120]), " s_full(j+1) seen after QK issue:", np.mean(t[1, 21:121] - t[8, 21:121]))
print("PV(j) issue after p_full(j):", np.mean(t[9, 20:120] - t[7, 20:120]), " o_full(j) seen (in j+1) after PV issue:", np.mean(t[6, 21:121] - t[9, 20:120]))
