"""GPU parity of K5 (tcgen05 GEMM / implicit-GEMM conv) through the C ABI against plain fp32 torch on the same
fp16-rounded operands.  The kernel accumulates fp16 x fp16 products exactly in fp32, so the only difference to the
fp32 reference is summation order: tolerance 2e-5 relative L2 (fp16 outputs: 1e-3 = one fp16 rounding)."""
import pytest
import torch
import torch.nn.functional as F

from signerf_b200 import nn_ops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 320, 320), (1000, 1920, 640), (77 * 2, 2560, 2048),
                                   (4096, 640, 2560), (129, 16, 72), (2, 1280, 320), (512, 640, 128), (2048, 1280, 256),
                                   (8192, 640, 64), (300, 1008, 64)])
def test_gemm_matches_fp32_reference(M, N, K):
    a = _rand((M, K), 1).half().cuda()
    w = _rand((N, K), 2, K ** -0.5).half().cuda()
    bias = _rand((N,), 3).cuda()
    out = nn_ops.gemm_f16(a, w, bias)
    ref = a.float() @ w.float().t() + bias
    assert rel_l2(out, ref) < 2e-5
    # residual in place + fp16 output + per-batch row bias
    res = _rand((M, N), 4).cuda()
    res0 = res.clone()
    nn_ops.gemm_f16(a, w, bias, residual=res, out=res)
    assert rel_l2(res, ref + res0) < 2e-5
    rows = (M + 1) // 2
    rb = _rand((2, N), 5).cuda()
    o16 = nn_ops.gemm_f16(a, w, None, rowbias=rb, rows_per_batch=rows, out_f16=True)
    ref_rb = a.float() @ w.float().t() + rb.repeat_interleave(rows, 0)[:M]
    assert o16.dtype == torch.float16 and rel_l2(o16, ref_rb) < 1e-3


def test_gemm_geglu_epilogue():
    M, C = 640, 320
    a = _rand((M, C), 1).half().cuda()
    w = _rand((8 * C, C), 2, C ** -0.5).half().cuda()       # sgm GEGLU proj: [2*inner, C], first half value, second gate
    b = _rand((8 * C,), 3).cuda()
    proj = a.float() @ w.float().t() + b
    val, gate = proj.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    wi = torch.stack([w[:4 * C], w[4 * C:]], 1).reshape(8 * C, C).contiguous()   # interleave (value_j, gate_j)
    bi = torch.stack([b[:4 * C], b[4 * C:]], 1).reshape(-1).contiguous()
    out = nn_ops.gemm_f16(a, wi, bi, geglu=True)
    assert out.shape == (M, 4 * C) and rel_l2(out, ref) < 1e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 64), (2, 16, 16, 128, 320), (2, 12, 20, 64, 96),
                                            (1, 5, 7, 64, 32), (2, 32, 32, 320, 4)])
def test_conv3x3_implicit_gemm_matches_torch(B, H, W, Cin, Cout):
    x = _rand((B, Cin, H, W), 1).half()
    w = _rand((Cout, Cin, 3, 3), 2, (9 * Cin) ** -0.5).half()
    bias = _rand((Cout,), 3)
    ref = F.conv2d(x.float().cuda(), w.float().cuda(), bias.cuda(), padding=1)              # NCHW fp32
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    w_packed = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous().cuda()
    emb = _rand((B, Cout), 4).cuda()
    if Cout % 16 == 0:
        out = nn_ops.conv3x3_f16(x_nhwc, w_packed, bias.cuda(), rowbias=emb)
        ref_e = (ref + emb[:, :, None, None]).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
        assert rel_l2(out, ref_e) < 2e-5
    out_nchw = nn_ops.conv3x3_f16(x_nhwc, w_packed, bias.cuda(), nchw=True)
    assert out_nchw.shape == ref.shape and rel_l2(out_nchw, ref) < 2e-5
