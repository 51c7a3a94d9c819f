"""Where does the plugin-level e2e unit spend its time?  CUDA events + host clock around each stage."""
import sys, time, torch
sys.path.insert(0, "/root/repo")
import signerf_b200.plugin as P
from signerf_b200 import ops, synthetic
from signerf_b200 import unet as unet_mod

dev = torch.device("cuda")
H = W = 512
fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=False)
ropts = ops.RenderOptions(mode="flat", num_samples=128)
graph = P.FusedNerfactoGraph(fld, ropts)
c2w, _ = synthetic.camera_ring(16, W, H)
c2w = c2w.contiguous().pin_memory()
unet = unet_mod.BenchUNet(dev, sheet_hw=(2048, 2048), seed=0) if "--unet" in sys.argv else None

class D(P.Diffuser):
    def diffuse(self, o, r, m=None, c=None):
        if unet is not None:
            D.latent = unet.step(o, m, c)
        return o

gcfg = P.DatasetGeneratorConfig(rows=4, cols=4, width=W, height=H, downscale_factor=1, fx=float(W), fy=float(W), cx=W / 2.0, cy=H / 2.0)
gen = gcfg.setup(original_transform_matrix=torch.eye(4)[:3], original_scale_factor=1.0, transform_poses_to_original_space=lambda x: x, device=dev)
gen.diffuser = D(gcfg.diffuser, dev)

def timed(name, fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:40s} gpu {a.elapsed_time(b) / n:9.3f} ms   host-enqueue {(t1 - t0) / n * 1e3:9.3f} ms   wall {(t2 - t0) / n * 1e3:9.3f} ms", flush=True)

cams = P.CameraBatch(c2w, float(W), float(W), W / 2.0, H / 2.0, W, H)
c2w_d, intr_d = P.base.c2w_intr(cams, dev)
timed("ops.render_views 16", lambda: ops.render_views(fld, c2w_d, intr_d, H, W, ropts))
timed("ops.render_views 15", lambda: ops.render_views(fld, c2w_d[:15].contiguous(), intr_d[:15].contiguous(), H, W, ropts))
timed("ops.render_views 1", lambda: ops.render_views(fld, c2w_d[15:].contiguous(), intr_d[15:].contiguous(), H, W, ropts))
timed("graph.render_cameras 15", lambda: graph.render_cameras(cams[:15]))
timed("gen.render_views 15", lambda: gen.render_views(graph, cams[:15]))
timed("gen.render_camera 1", lambda: gen.render_camera(graph, cams[15]))
rgb, depth = ops.render_views(fld, c2w_d, intr_d, H, W, ropts)
timed("ops.mask_condition 16", lambda: ops.mask_condition(c2w_d, intr_d, depth, gen._mask_options()))
timed("ops.mask_condition 1", lambda: ops.mask_condition(c2w_d[:1].contiguous(), intr_d[:1].contiguous(), depth[:1].contiguous(), gen._mask_options()))
timed("gen.generate_reference_sheet", lambda: gen.generate_reference_sheet(graph, cams[:15], W, H))
lay = gen._layout(W, H)
sheet = torch.rand(lay.height, lay.width, 3, device=dev)
timed("15 x sheet_cut", lambda: [ops.sheet_cut(sheet, lay, i, H, W) for i in range(15)])
