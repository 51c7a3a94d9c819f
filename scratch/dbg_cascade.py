import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import nerfacto_ref as R
from signerf_b200 import ops
from tests.helpers import field_from_oracle, ring_cameras
# rebuild the same models as the fixture
fx = {}
for name, kw in {"varied": dict(dense=True, table_scale=0.5, density_gain=20.0), "sparse": dict(table_scale=0.5, density_gain=40.0)}.items():
    m = R.make_model(0, **kw)
    fx[name] = (m, field_from_oracle(m))
for name in ("varied", "sparse"):
    m, f = fx[name]
    c2w, intr = ring_cameras(2, 48, 32)
    for mode in (ops.MLP_FP32, ops.MLP_FP16_MMA):
        opts = ops.RenderOptions(mode="cascade", num_samples=48, num_prop_samples=(256, 96), mlp_mode=mode)
        rgb, depth, acc = ops.render_views(f, c2w.cuda(), intr.cuda(), 32, 48, opts, want_acc=True)
        for v in range(2):
            ref = R.render_view(m, c2w[v], *intr[v].tolist(), 48, 32, "cascade")
            d = depth[v].cpu().double().flatten(); r = ref["depth"].double().flatten()
            rel = ((d - r).abs() / r.abs())
            q = torch.quantile(rel, torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], dtype=torch.double))
            print(name, mode, v, "rel quantiles", [f"{x:.2e}" for x in q.tolist()], "relL2", float((d-r).norm()/r.norm()),
                  "frac>1e-4", float((rel > 1e-4).double().mean()), "frac>1e-3", float((rel > 1e-3).double().mean()))
