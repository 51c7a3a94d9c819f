"""GPU parity at the shapes bench.py actually times (BASELINE config C3: 2048^2 sheet, latent 256^2, CFG batch 2).

The per-kernel parity files cover every code path on small shapes; this file repeats the checks on the launches that
dominate the UNet step (profiles/r1_unet_step_by_shape.txt), because the tile schedule depends on the shape:
  * k_attention_tc at 2 x 10 heads x 16 384^2 and 2 x 20 heads x 4 096^2 (128 / 32 key tiles per row: the lazy rescale
    and the fp16 P accumulate over the whole row) against a fp32 softmax computed head by head (1 GB score slab);
  * k_gemm_tc with the in-place fp32 residual on the tail-split schedule (8192x1280x{1280,5120}, 32768x640x2560), the
    narrow-tile schedule (32768x640x640) and the GEGLU epilogue on the split-2 schedule (8192x10240x1280); the plan the
    library reports for the shape (sgn_gemm_plan) is asserted, so the test fails if the schedule under test changes;
  * the 3x3 implicit-GEMM conv at the sheet's three resolutions.
Tolerances are the north star's 1e-3 relative L2 for fp16-rounded outputs, 2e-5 where the output stays fp32."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from signerf_b200 import _lib, nn_ops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _plan(M, N, K, res=0):
    p = (C.c_int * 6)()
    assert _lib.load().sgn_gemm_plan(M, N, K, res, p) == 0
    return dict(zip(("bn", "cluster", "tiles", "split", "items", "units"), p))


def _randn(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


@pytest.mark.parametrize("B,heads,T", [(2, 10, 16384), (2, 20, 4096)])
def test_self_attention_at_benchmark_shapes(B, heads, T):
    Cc = heads * 64
    qkv = _randn((B * T, 3 * Cc), T, 1.5, torch.float16)             # fused projection layout, as the UNet calls it
    q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
    out = nn_ops.attention_f16(q, k, v, B, heads)
    torch.cuda.synchronize()
    num = torch.zeros((), dtype=torch.float64, device="cuda")
    den = torch.zeros((), dtype=torch.float64, device="cuda")
    worst = 0.0
    for b in range(B):
        rows = slice(b * T, (b + 1) * T)
        for h in range(heads):
            cols = slice(h * 64, (h + 1) * 64)
            s = (q[rows, cols].float() @ k[rows, cols].float().t()) * 0.125        # [T, T] fp32
            ref = torch.softmax(s, dim=-1) @ v[rows, cols].float()
            d = out[rows, cols].float() - ref
            num += d.double().pow(2).sum()
            den += ref.double().pow(2).sum()
            worst = max(worst, float(d.abs().max()))
            del s, ref, d
    err = float((num / den).sqrt())
    print(f"attention {B}x{heads}x{T}^2: rel-L2 {err:.2e}, max abs {worst:.2e}")
    assert err < 1e-3
    assert worst < 2e-2


@pytest.mark.parametrize("M,N,K,want_split", [(8192, 1280, 1280, True), (8192, 1280, 5120, True), (32768, 640, 2560, True),
                                              (32768, 640, 640, False)])
def test_in_place_residual_gemm_at_benchmark_shapes(M, N, K, want_split):
    plan = _plan(M, N, K, 1)
    assert plan["cluster"] == 2
    assert (plan["split"] > 1) == want_split, plan
    a = _randn((M, K), 1, 1.0, torch.float16)
    w = _randn((N, K), 2, K ** -0.5, torch.float16)
    bias = _randn((N,), 3)
    res = _randn((M, N), 4)
    ref = a.float() @ w.float().t() + bias + res
    nn_ops.gemm_f16(a, w, bias, residual=res, out=res)               # the residual stream is updated in place
    err = rel_l2(res, ref)
    print(f"gemm {M}x{N}x{K} res in place (plan {plan}): rel-L2 {err:.2e}")
    assert err < 2e-5
    o16 = nn_ops.gemm_f16(a, w, bias, out_f16=True)
    assert rel_l2(o16, a.float() @ w.float().t() + bias) < 1e-3


@pytest.mark.parametrize("M,C", [(8192, 1280), (32768, 640)])
def test_geglu_gemm_at_benchmark_shapes(M, C):
    N = 8 * C
    plan = _plan(M, N, C)
    if (M, C) == (8192, 1280):
        assert plan["split"] == 2, plan                                # the split-2 tail schedule of the 1280-channel FF
    a = _randn((M, C), 1, 1.0, torch.float16)
    w = _randn((N, C), 2, C ** -0.5, torch.float16)                    # sgm GEGLU proj: first half value, second half gate
    b = _randn((N,), 3)
    proj = a.float() @ w.float().t() + b
    ref = proj[:, :N // 2] * F.gelu(proj[:, N // 2:])
    del proj
    wi = torch.stack([w[:N // 2], w[N // 2:]], 1).reshape(N, C).contiguous()   # rows interleaved (value_j, gate_j)
    bi = torch.stack([b[:N // 2], b[N // 2:]], 1).reshape(-1).contiguous()
    out = nn_ops.gemm_f16(a, wi, bi, geglu=True)
    err = rel_l2(out, ref)
    print(f"geglu {M}x{N}x{C} (plan {plan}): rel-L2 {err:.2e}")
    assert out.shape == (M, N // 2) and err < 1e-3


@pytest.mark.parametrize("M,N,K", [(8192, 3840, 1280), (32768, 1920, 640)])
def test_qkv_projection_at_benchmark_shapes(M, N, K):
    a = _randn((M, K), 1, 1.0, torch.float16)
    w = _randn((N, K), 2, K ** -0.5, torch.float16)
    out = nn_ops.gemm_f16(a, w, None, out_f16=True)
    assert rel_l2(out, a.float() @ w.float().t()) < 1e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 1280, 1280), (2, 128, 128, 640, 640), (2, 256, 256, 320, 320),
                                            (2, 64, 64, 2560, 1280), (2, 256, 256, 320, 4)])
def test_conv3x3_at_benchmark_shapes(B, H, W, Cin, Cout):
    x = _randn((B, Cin, H, W), 1, 1.0, torch.float16)
    w = _randn((Cout, Cin, 3, 3), 2, (9 * Cin) ** -0.5, torch.float16)
    bias = _randn((Cout,), 3)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1)              # fp32 NCHW (TF32 is off: tests/conftest.py)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_packed = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    if Cout % 16 == 0:
        emb = _randn((B, Cout), 4)
        out = nn_ops.conv3x3_f16(x_nhwc, w_packed, bias, rowbias=emb)
        ref_e = (ref + emb[:, :, None, None]).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
        assert rel_l2(out, ref_e) < 2e-5
    else:                                                              # the UNet's 4-channel output conv stores NCHW
        out = nn_ops.conv3x3_f16(x_nhwc, w_packed, bias, nchw=True)
        assert out.shape == ref.shape and rel_l2(out, ref) < 2e-5
