"""Single-kernel drivers for `ncu --set full` captures (one GPU, a handful of launches).
    python profiles/run_kernel.py attention | gemm | conv | render"""
import sys
import torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops

which = sys.argv[1]
if len(sys.argv) > 2:   # run_kernel.py attention <attn_variant>
    from signerf_b200 import _lib
    _lib.set_option("attn_variant", int(sys.argv[2]))
if which == "attention":      # self-attention of the 640-channel level at the 2048^2 sheet
    B, heads, T = 2, 10, 16384
    qkv = torch.randn(B * T, 3 * heads * 64, device="cuda").half()
    c = heads * 64
    for _ in range(3):
        nn_ops.attention_f16(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], B, heads)
elif which == "gemm":         # GEGLU projection of the 1280-channel level
    a = torch.randn(8192, 1280, device="cuda").half()
    w = torch.randn(10240, 1280, device="cuda").half()
    for _ in range(3):
        nn_ops.gemm_f16(a, w, None, out_f16=True)
elif which == "conv":
    x = torch.randn(2, 128, 128, 640, device="cuda").half()
    w = torch.randn(640, 9 * 640, device="cuda").half()
    for _ in range(3):
        nn_ops.conv3x3_f16(x, w, None)
torch.cuda.synchronize()
if which == "render":
    from signerf_b200 import ops, synthetic
    dev = torch.device("cuda")
    fld = synthetic.random_field(seed=0, device=dev, dense=True, with_proposals=False)
    c2w, intr = synthetic.camera_ring(16, 512, 512)
    for _ in range(2):
        ops.render_views(fld, c2w.to(dev), intr.to(dev), 512, 512, ops.RenderOptions(mode="flat", num_samples=128))
    torch.cuda.synchronize()
if which == "gemm_mid":       # attention out-projection of the 1280-channel level: the most frequent GEMM shape
    a = torch.randn(8192, 1280, device="cuda").half()
    w = torch.randn(1280, 1280, device="cuda").half()
    r = torch.randn(8192, 1280, device="cuda")
    b = torch.randn(1280, device="cuda")
    for _ in range(3):
        nn_ops.gemm_f16(a, w, b, residual=r, out=r)
    torch.cuda.synchronize()
if which == "cross_attention":   # attn2 of the 1280-channel level: 4096 queries x 77 prompt tokens, 20 heads
    B, heads, T = 2, 20, 4096
    c = heads * 64
    q = torch.randn(B * T, c, device="cuda").half()
    kv = torch.randn(B * 77, 2 * c, device="cuda").half()
    for _ in range(3):
        nn_ops.attention_f16(q, kv[:, :c], kv[:, c:], B, heads)
    torch.cuda.synchronize()
if which == "hint_conv":         # ControlNet input_hint_block 16 -> 16 at the 2048^2 sheet, fp32-exact on mma.sync
    x = torch.randn(1, 2048, 2048, 16, device="cuda")
    w = torch.randn(16, 9 * 16, device="cuda").half()
    b = torch.randn(16, device="cuda")
    for _ in range(3):
        nn_ops.conv3x3_small_tc(x, w, b, act_silu=True)
    torch.cuda.synchronize()
