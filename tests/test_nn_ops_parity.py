"""GPU parity of K7-K9 (norms, data movement, direct convs, embedding path, sampler update) against plain fp32 torch
through the C ABI.  fp16 outputs are compared at 1e-3 relative L2 (one fp16 rounding = 4.9e-4 max relative), fp32
outputs at 1e-5."""
import math

import pytest
import torch
import torch.nn.functional as F

from signerf_b200 import nn_ops as K
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _r(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).cuda()


@pytest.mark.parametrize("B,H,W,C,silu,eps", [(2, 16, 16, 320, True, 1e-5), (1, 9, 7, 64, False, 1e-6), (2, 32, 32, 960, True, 1e-5)])
def test_group_norm(B, H, W, C, silu, eps):
    x = _r((B, C, H, W), 1, 2.0, 0.7)
    g, b = _r((C,), 2, 0.5, 1.0), _r((C,), 3)
    ref = F.group_norm(x, 32, g, b, eps)
    if silu:
        ref = F.silu(ref)
    x_nhwc = x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    out = K.group_norm_f16(x_nhwc, B, H * W, 32, eps, g, b, silu)
    assert torch.equal(out, K.group_norm_f16(x_nhwc, B, H * W, 32, eps, g, b, silu))      # deterministic reduction
    assert rel_l2(out.view(B, H, W, C).permute(0, 3, 1, 2), ref) < 1e-3


@pytest.mark.parametrize("M,C", [(100, 128), (4096, 640), (333, 1280), (50, 192)])
def test_layer_norm(M, C):
    x = _r((M, C), 1, 3.0, -0.5)
    g, b = _r((C,), 2, 0.5, 1.0), _r((C,), 3)
    assert rel_l2(K.layer_norm_f16(x, g, b), F.layer_norm(x, (C,), g, b, 1e-5)) < 1e-3


def test_cast_upsample_concat_axpy():
    B, H, W, C = 2, 6, 5, 64
    x = _r((B * H * W, C), 1)
    assert torch.equal(K.cast_f16(x), x.half())
    up = K.upsample2x_f16(x, B, H, W)
    ref = F.interpolate(x.view(B, H, W, C).permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up, ref.half())
    b, b2 = _r((B * H * W, 128), 2), _r((B * H * W, 128), 3)
    assert torch.equal(K.concat_f32(x, b), torch.cat([x, b], 1))
    assert torch.allclose(K.concat_f32(x, b, b2, 0.8), torch.cat([x, b + 0.8 * b2], 1), atol=1e-6)
    y = b.clone()
    K.axpy_f32(b2, 0.8, y)
    assert torch.allclose(y, b + 0.8 * b2, atol=1e-6)


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 64), (1, 9, 7, 128)])
def test_downsample_conv_via_im2col(B, H, W, C):
    x = _r((B, C, H, W), 1)
    w = _r((C, C, 3, 3), 2, (9 * C) ** -0.5)
    bias = _r((C,), 3)
    x_nhwc = x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    col, ho, wo = K.im2col3x3_s2_f16(x_nhwc, B, H, W)
    out = K.gemm_f16(col, w.permute(0, 2, 3, 1).reshape(C, 9 * C).half().contiguous(), bias)
    ref = F.conv2d(x.half().float(), w.half().float(), bias, stride=2, padding=1)
    assert (ho, wo) == tuple(ref.shape[2:])
    assert rel_l2(out.view(B, ho, wo, C).permute(0, 3, 1, 2), ref) < 2e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,nchw", [(2, 32, 32, 4, 320, 1, True), (1, 64, 48, 3, 16, 1, True),
                                                        (1, 33, 31, 16, 32, 2, False), (2, 16, 16, 96, 256, 2, False),
                                                        (1, 20, 20, 32, 32, 1, False)])
def test_direct_conv(B, H, W, Cin, Cout, stride, nchw):
    x = _r((B, Cin, H, W), 1)
    w = _r((Cout, Cin, 3, 3), 2, (9 * Cin) ** -0.5)
    bias = _r((Cout,), 3)
    ref = F.silu(F.conv2d(x, w, bias, stride=stride, padding=1))
    xin = x if nchw else x.permute(0, 2, 3, 1).contiguous()
    out = K.conv3x3_direct(xin, nchw, w.permute(0, 2, 3, 1).contiguous(), bias, stride=stride, act_silu=True)
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 1e-5
    # residual shared by the CFG pair (one residual image for B outputs) + fp16 output
    res = _r((1,) + tuple(out.shape[1:]), 4)
    out2 = K.conv3x3_direct(xin, nchw, w.permute(0, 2, 3, 1).contiguous(), bias, residual=res, stride=stride, out_f16=True)
    ref2 = F.conv2d(x, w, bias, stride=stride, padding=1) + res.permute(0, 3, 1, 2)
    assert out2.dtype == torch.float16 and rel_l2(out2.permute(0, 3, 1, 2), ref2) < 1e-3


def test_embedding_path():
    t = torch.tensor([801.25, 3.5], device="cuda")
    emb = K.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None] * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert float((emb - ref).abs().max()) < 2e-4      # fp32 sin/cos of arguments up to ~800
    x, w, b, r = _r((2, 320), 1), _r((1280, 320), 2, 320 ** -0.5), _r((1280,), 3), _r((2, 1280), 4)
    out = K.linear_small(x, w, b, residual=r, silu_in=True, silu_out=True)
    assert rel_l2(out, F.silu(F.linear(F.silu(x), w, b) + r)) < 1e-5


def test_cfg_euler_ancestral_step():
    B, C, H, W = 2, 4, 16, 24
    x, init, noise = _r((B, C, H, W), 1, 5.0), _r((B, C, H, W), 2), _r((B, C, H, W), 3)
    eps = _r((2 * B, C, H, W), 4)
    mask = (torch.rand(B, 1, H, W, device="cuda") > 0.5).float()
    sigma, down, up, cfg = 7.3, 5.1, 2.2, 7.0
    xn, den = K.cfg_euler_step(x, eps, init, mask, noise, cfg, sigma, down, up)
    e = eps[B:] + cfg * (eps[:B] - eps[B:])
    d_ref = init * mask + (1 - mask) * (x - sigma * e)
    x_ref = x + (x - d_ref) / sigma * (down - sigma) + noise * up
    assert torch.allclose(den, d_ref, rtol=1e-5, atol=1e-5) and torch.allclose(xn, x_ref, rtol=1e-5, atol=1e-4)
    x2 = K.scale_cat2(x, 0.25)
    assert torch.equal(x2, torch.cat([x, x]) * 0.25)


def test_sheet_to_conditioning():
    Hs, Ws = 64, 96
    cond, mask = torch.rand(Hs, Ws, 1, device="cuda"), (torch.rand(Hs, Ws, 1, device="cuda") > 0.4).float()
    hint, lat = torch.empty(1, 3, Hs, Ws, device="cuda"), torch.empty(1, 1, Hs // 8, Ws // 8, device="cuda")
    K.make_hint_and_latent_mask(cond, mask, hint, lat)
    q = (cond[..., 0] * 255).to(torch.uint8).float() / 255
    assert torch.equal(hint[0], q.expand(3, Hs, Ws))
    ref = 1 - torch.round(F.avg_pool2d(mask[..., 0][None, None], 8))
    assert torch.equal(lat, ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,nchw", [(1, 64, 48, 3, 16, 1, True), (1, 33, 31, 16, 32, 2, False),
                                                        (2, 16, 16, 96, 256, 2, False), (1, 20, 20, 32, 32, 1, False)])
def test_small_channel_conv_on_tensor_cores_is_fp32_exact(B, H, W, Cin, Cout, stride, nchw):
    """hi/lo split im2col + [W | W] GEMM + SiLU epilogue == fp32 conv (weights fp16-representable)."""
    x = _r((B, Cin, H, W), 1)
    w = _r((Cout, Cin, 3, 3), 2, (9 * Cin) ** -0.5).half().float()
    bias = _r((Cout,), 3)
    ref = F.silu(F.conv2d(x, w, bias, stride=stride, padding=1))
    xin = x if nchw else x.permute(0, 2, 3, 1).contiguous()
    col, ho, wo, kp = K.im2col3x3_split_f16(xin, nchw, stride)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, -1)
    wp = torch.zeros(Cout, kp, device="cuda")
    wp[:, :wk.shape[1]] = wk
    out = K.gemm_f16(col, torch.cat([wp, wp], 1).half().contiguous(), bias, act_silu=True)
    assert (ho, wo) == tuple(ref.shape[2:])
    assert rel_l2(out.view(B, ho, wo, Cout).permute(0, 3, 1, 2), ref) < 2e-6


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride", [(1, 64, 48, 16, 16, 1), (2, 33, 31, 16, 32, 1), (1, 50, 70, 32, 32, 1),
                                                   (1, 16, 16, 32, 32, 1), (1, 5, 3, 16, 16, 1), (1, 64, 96, 16, 32, 2),
                                                   (2, 33, 31, 16, 32, 2)])
@pytest.mark.parametrize("out_f16", [False, True])
def test_small_tc_conv_is_fp32_exact(B, H, W, Cin, Cout, stride, out_f16):
    """sgn_conv3x3_small_tc: hi/lo split in shared memory + mma.sync == fp32 conv + SiLU (weights fp16-representable),
    ragged tiles included."""
    x = _r((B, Cin, H, W), 11) * 3
    w = _r((Cout, Cin, 3, 3), 12, (9 * Cin) ** -0.5).half().float()
    bias = _r((Cout,), 13)
    ref = F.silu(F.conv2d(x, w, bias, stride=stride, padding=1))
    w16 = w.permute(0, 2, 3, 1).reshape(Cout, -1).half().contiguous()
    out = K.conv3x3_small_tc(x.permute(0, 2, 3, 1).contiguous(), w16, bias, stride=stride, act_silu=True, out_f16=out_f16)
    assert out.dtype == (torch.float16 if out_f16 else torch.float32)
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < (6e-4 if out_f16 else 2e-6)
    lin = K.conv3x3_small_tc(x.permute(0, 2, 3, 1).contiguous(), w16, None, stride=stride)
    assert rel_l2(lin.permute(0, 3, 1, 2), F.conv2d(x, w, None, stride=stride, padding=1)) < 2e-6


def test_sdxl_vector_conditioning():
    """y = [pooled | fourier256(orig_h, orig_w, crop_t, crop_l, target_h, target_w)] (sgm ConcatTimestepEmbedderND)."""
    import math
    from signerf_b200 import conditioning as Cn
    pooled = _r((2, 1280), 21)
    y = Cn.sdxl_vector(pooled, 2048, 1536, (0, 0))
    assert tuple(y.shape) == (2, 2816) and torch.equal(y[:, :1280], pooled)
    freqs = torch.exp(-math.log(10000.0) * torch.arange(128, dtype=torch.float32) / 128).cuda()
    ref = []
    for v in (2048.0, 1536.0, 0.0, 0.0, 2048.0, 1536.0):
        a = v * freqs
        ref.append(torch.cat([torch.cos(a), torch.sin(a)]))
    ref = torch.cat(ref)
    assert torch.allclose(y[0, 1280:], ref, atol=2e-4) and torch.equal(y[0, 1280:], y[1, 1280:])
