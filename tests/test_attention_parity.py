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
