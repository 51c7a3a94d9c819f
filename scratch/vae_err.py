import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import vae_ref as VR
from signerf_b200 import vae as V
from tests.helpers import rel_l2
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for name, cfg in (("tiny", VR.tiny_vae_config()), ("sdxl", VR.VAEConfig())):
    ref = VR.make_vae(cfg, seed=3, device="cuda")
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).cuda()
    z = ref.encode(x)
    for exact in (True, False):
        net = V.VAEB200(V.VAEConfig(**cfg.__dict__), ref.state_dict(), "cuda", exact=exact)
        print(name, "exact" if exact else "single", "encode", f"{rel_l2(net.encode(x), z):.2e}", "decode", f"{rel_l2(net.decode(z), ref.decode(z)):.2e}", flush=True)
