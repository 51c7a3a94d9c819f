"""GPU: the two SDXL text encoders (signerf_b200/text_encoder.py) against HuggingFace transformers' own CLIPTextModel /
CLIPTextModelWithProjection - the implementation A1111 / diffusers run - on random fp16-representable weights: CLIP-L's real
width / heads with 3 layers, bigG's real width / heads / gelu with 3 layers, and both at full depth on the GPU oracle.
Tolerance 1e-3 relative L2 (fp16 operands, fp32 accumulate / residual stream)."""
import pytest
import torch

from tests.helpers import rel_l2

transformers = pytest.importorskip("transformers")
pytestmark = pytest.mark.gpu


def _hf(cfg, seed, projection):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    hc = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                        max_position_embeddings=77, hidden_act=cfg.hidden_act, projection_dim=cfg.projection_dim or 512, eos_token_id=2,
                        layer_norm_eps=cfg.layer_norm_eps)
    torch.manual_seed(seed)
    m = (CLIPTextModelWithProjection if projection else CLIPTextModel)(hc).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(3.0 if p.dim() > 1 else 1.0)          # default init is tiny: make attention / MLP matter
            p.copy_(p.half().float())
    return m


@pytest.mark.parametrize("which,layers", [("clip_l", 3), ("bigg", 3), ("clip_l", 12), ("bigg", 32)])
def test_text_encoder_matches_transformers(which, layers):
    from signerf_b200 import text_encoder as TE
    cfg = TE.CLIPTextConfig.clip_l() if which == "clip_l" else TE.CLIPTextConfig.open_clip_bigg()
    cfg.num_hidden_layers, cfg.vocab_size = layers, 4096
    ref = _hf(cfg, 0, cfg.projection_dim is not None).cuda()
    enc = TE.CLIPTextEncoderB200(cfg, ref.state_dict(), "cuda")
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 4000, (2, 77), generator=g)
    ids[0, 20:] = 0
    ids[0, 19] = 4095                                    # EOT = the largest id (argmax pooling), then padding
    ids[1, 76] = 4095
    with torch.no_grad():
        o = ref(ids.cuda(), output_hidden_states=True)
    out = enc.forward(ids)
    e_pen = rel_l2(out["penultimate"], o.hidden_states[-2])
    e_last = rel_l2(out["last"], o.last_hidden_state)
    pooled_ref = o.text_embeds if cfg.projection_dim is not None else o.pooler_output
    e_pool = rel_l2(out["pooled"], pooled_ref)
    print(f"{which} x{layers}: penultimate {e_pen:.1e} last {e_last:.1e} pooled {e_pool:.1e}")
    assert e_pen < 1e-3 and e_last < 1e-3 and e_pool < 1e-3


def test_sdxl_prompt_conditioning_shapes():
    from signerf_b200 import text_encoder as TE
    cl, cg = TE.CLIPTextConfig.clip_l(), TE.CLIPTextConfig.open_clip_bigg()
    cl.num_hidden_layers = cg.num_hidden_layers = 2
    cl.vocab_size = cg.vocab_size = 512
    a = TE.CLIPTextEncoderB200(cl, _hf(cl, 0, False).state_dict(), "cuda")
    b = TE.CLIPTextEncoderB200(cg, _hf(cg, 1, True).state_dict(), "cuda")
    ids = torch.randint(1, 500, (2, 77))
    ctx, y = TE.sdxl_prompt_conditioning(a, b, ids, ids, 2048, 2048)
    assert tuple(ctx.shape) == (2, 77, 2048) and tuple(y.shape) == (2, 2816) and ctx.dtype == torch.float32
    assert bool(torch.isfinite(ctx).all()) and bool(torch.isfinite(y).all())
