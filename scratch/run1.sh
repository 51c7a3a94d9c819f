cd /root/repo
python scratch/attn_split_check.py 2>&1 | tail -8
python scratch/attn_growth.py 0,1,4,12 2>&1 | tail -8
python scratch/bench_attn_var.py 0,1,4,5,8,12,13 2>&1 | tail -40
python - <<'PY'
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
exec(open("scratch/bench_attn_var.py").read().split("variants =")[0])
for sp in (0, 1):
    _lib.set_option("attn_split", sp)
    for B, heads, T in ((2, 10, 16384), (2, 20, 4096)):
        C = heads * 64
        q, k, v = (torch.randn(B * T, C, device="cuda").half() for _ in range(3))
        out = torch.empty(B * T, C, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: nn_ops.attention_f16(q, k, v, B, heads, out=out), n=20)
        print(f"split {sp} T{T}: {ms:.3f} ms {4*B*heads*T*T*64/ms/1e9:.0f} TF/s")
PY
for v in 0 1; do python scratch/attn_trace.py $v 2>&1 | tail -8; done
