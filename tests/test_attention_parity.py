"""GPU parity of K6 (tcgen05 flash attention) against fp32 torch softmax(QK^T/8)V on the same fp16 inputs.
P is rounded to fp16 before the PV product and the output is fp16: tolerance 1e-3 relative L2."""
import pytest
import torch

from signerf_b200 import nn_ops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _ref(q, k, v, B, heads):
    Tq, Tkv = q.shape[0] // B, k.shape[0] // B
    qf = q.float().view(B, Tq, heads, 64).transpose(1, 2)
    kf = k.float().view(B, Tkv, heads, 64).transpose(1, 2)
    vf = v.float().view(B, Tkv, heads, 64).transpose(1, 2)
    p = torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, dim=-1)
    return (p @ vf).transpose(1, 2).reshape(B * Tq, heads * 64)


@pytest.mark.parametrize("B,heads,Tq,Tkv", [(1, 1, 128, 128), (2, 2, 256, 256), (2, 3, 64, 64), (1, 2, 200, 333),
                                            (2, 10, 1024, 1024), (2, 4, 256, 77), (1, 20, 4096, 4096), (2, 3, 200, 77), (1, 2, 70, 5),
                                            (2, 5, 1000, 80), (1, 1, 64, 64), (2, 20, 4096, 77)])
def test_attention_matches_fp32_reference(B, heads, Tq, Tkv):
    g = torch.Generator().manual_seed(Tq * 7 + Tkv)
    C = heads * 64
    self_attn = Tq == Tkv
    if self_attn:   # fused QKV layout [B*T, 3C]: column slices
        qkv = (torch.randn(B * Tq, 3 * C, generator=g) * 1.5).half().cuda()
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:           # cross attention: q [B*Tq, C], fused KV [B*Tkv, 2C]
        q = (torch.randn(B * Tq, C, generator=g) * 1.5).half().cuda()
        kv = (torch.randn(B * Tkv, 2 * C, generator=g) * 1.5).half().cuda()
        k, v = kv[:, :C], kv[:, C:]
    out = nn_ops.attention_f16(q, k, v, B, heads)
    ref = _ref(q, k, v, B, heads)
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < 1e-3
    assert float((out.float() - ref).abs().max()) < 2e-2


@pytest.fixture
def attn_options():
    """restores the library's attention options after a test that changes them"""
    from signerf_b200 import _lib
    yield _lib.set_option
    for name, value in (("attn_shape", 0), ("attn_split", 1), ("attn_variant", 3)):
        _lib.set_option(name, value)


@pytest.mark.parametrize("shape", [0, 1])
@pytest.mark.parametrize("B,heads,T", [(2, 20, 4096), (1, 3, 1000), (1, 2, 777), (2, 5, 2304), (1, 1, 130)])
def test_key_range_split_of_tail_items(attn_options, shape, B, heads, T):
    """Work items of an under-filled last wave are cut into key ranges and combined in-kernel (sgn_attention_f16_ws):
    same result as the unsplit launch to fp16 rounding and within tolerance of fp32, for both CTA shapes, ragged T
    included.  (2, 20, 4096) is the benchmark's 4.32-wave launch."""
    g = torch.Generator().manual_seed(T + shape)
    C = heads * 64
    q, k, v = ((torch.randn(B * T, C, generator=g) * 1.5).half().cuda() for _ in range(3))
    attn_options("attn_shape", shape)
    attn_options("attn_split", 0)
    whole = nn_ops.attention_f16(q, k, v, B, heads)
    attn_options("attn_split", 1)
    split = nn_ops.attention_f16(q, k, v, B, heads)
    again = nn_ops.attention_f16(q, k, v, B, heads)
    ref = _ref(q, k, v, B, heads)
    assert torch.equal(split, again), "the combine must not depend on which CTA arrives last"
    assert rel_l2(split, ref) < 1e-3 and rel_l2(whole, ref) < 1e-3
    assert float((split.float() - whole.float()).abs().max()) < 2e-3


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [0, 1])
def test_row_maximum_growing_over_the_key_tiles(attn_options, shape, variant):
    """Later keys score higher and higher (x6 ramp), so the running maximum of every row outgrows the lazy-rescale
    threshold (2^8) many times: the O / l rescale path of every issuer variant and CTA shape against fp32."""
    B, heads, T = 1, 2, 2048
    C = heads * 64
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B * T, C, generator=g) * 3.0
    k = torch.randn(B * T, C, generator=g) * (1.0 + 5.0 * torch.arange(T) / T)[:, None]
    v = torch.randn(B * T, C, generator=g)
    q, k, v = q.half().cuda(), k.half().cuda(), v.half().cuda()
    attn_options("attn_shape", shape)
    attn_options("attn_variant", variant)
    out = nn_ops.attention_f16(q, k, v, B, heads)
    assert torch.isfinite(out).all()
    assert rel_l2(out, _ref(q, k, v, B, heads)) < 1e-3
