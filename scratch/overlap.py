"""Feasibility: does a 1-CTA/SM renderer co-run with the CTA-pair GEMM (5 stages) on two streams?"""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import _lib, ops, synthetic, nn_ops as K
dev = torch.device("cuda")
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=False)
c2w, intr = synthetic.camera_ring(16, 512, 512)
c2w, intr = c2w.to(dev), intr.to(dev)
ropts = ops.RenderOptions(mode="flat", num_samples=128)
a = torch.randn(8192, 1280, device=dev).half(); w = torch.randn(10240, 1280, device=dev).half()
o = torch.empty(8192, 10240, device=dev, dtype=torch.float16)
a2 = torch.randn(8192, 1280, device=dev).half(); w2 = torch.randn(1280, 1280, device=dev).half()
o2 = torch.empty(8192, 1280, device=dev, dtype=torch.float16)
NG = 300
def gemms():
    for _ in range(NG):
        K.gemm_f16(a, w, None, out_f16=True, out=o)
        K.gemm_f16(a2, w2, None, out_f16=True, out=o2)
def render():
    return ops.render_views(fld, c2w, intr, 512, 512, ropts)
def t(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for stages in (6, 5):
    _lib.set_option("gemm_pair_stages", stages)
    gemms(); print(f"stages {stages}: {NG} GEMM pairs alone {t(gemms):.1f} ms")
for ctas in (4, 2, 1):
    _lib.set_option("render_ctas_per_sm", ctas)
    render(); print(f"render alone, {ctas} CTA/SM: {t(render):.1f} ms")
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
for ctas in (1, 2):
    _lib.set_option("render_ctas_per_sm", ctas); _lib.set_option("gemm_pair_stages", 5)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); eA = torch.cuda.Event(enable_timing=True); eB = torch.cuda.Event(enable_timing=True)
    e0.record()
    sA.wait_event(e0); sB.wait_event(e0)
    with torch.cuda.stream(sA):
        render(); eA.record()
    with torch.cuda.stream(sB):
        gemms(); eB.record()
    torch.cuda.synchronize()
    print(f"concurrent ({ctas} render CTA/SM, 5-stage GEMM): render done at {e0.elapsed_time(eA):.1f} ms, GEMMs done at {e0.elapsed_time(eB):.1f} ms")
