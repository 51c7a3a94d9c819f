"""kernel-to-kernel gap inside a CUDA-graph replay on this box: 1 000 dependent tiny launches (sgn_axpy_f32 on 4 floats)"""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops as K
x = torch.zeros(4, device="cuda"); y = torch.ones(4, device="cuda")
K.axpy_f32(y, 1.0, x); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(1000):
        K.axpy_f32(y, 1.0, x)
for _ in range(3): g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): g.replay()
b.record(); torch.cuda.synchronize()
print(f"graph replay: {a.elapsed_time(b) / 10 / 1000 * 1e3:.2f} us per dependent tiny kernel")
a.record()
for _ in range(10000): K.axpy_f32(y, 1.0, x)
b.record(); torch.cuda.synchronize()
print(f"eager: {a.elapsed_time(b) / 10000 * 1e3:.2f} us per launch")
