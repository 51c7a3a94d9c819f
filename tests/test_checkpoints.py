"""CPU: weight-file formats either side of the diffusion path (signerf_b200/checkpoints.py).  No weights exist offline, so the
source formats are rebuilt from their architectures: the diffusers ControlNetModel parameter tree is enumerated from the
SDXL ControlNet's block structure and must land, name for name and shape for shape, on the cldm tree the engine loads
(`param_schema(cfg, controlnet=True)`); the open_clip text tower is checked against `transformers`' own module."""
import pytest
import torch

from signerf_b200 import checkpoints as CK
from signerf_b200 import unet as U


def _diffusers_controlnet_schema(cfg: U.UNetConfig):
    """Parameter names / shapes of diffusers' ControlNetModel for SDXL (DownBlock2D, CrossAttnDownBlock2D x 2,
    UNetMidBlock2DCrossAttn; use_linear_projection; conditioning_embedding_out_channels (16, 32, 96, 256))."""
    mc, ted, ctx, adm = cfg.model_channels, 4 * cfg.model_channels, cfg.context_dim, cfg.adm_in_channels
    chans = [mc * m for m in cfg.channel_mult]
    depth = [d if (1 << i) in cfg.attention_resolutions else 0 for i, d in enumerate(cfg.transformer_depth)]   # DownBlock2D at level 0
    s = {}

    def lin(n, o, i, bias=True):
        s[n + ".weight"] = (o, i)
        if bias:
            s[n + ".bias"] = (o,)

    def conv(n, o, i, k=3):
        s[n + ".weight"], s[n + ".bias"] = (o, i, k, k), (o,)

    def norm(n, c):
        s[n + ".weight"], s[n + ".bias"] = (c,), (c,)

    def resnet(n, cin, cout):
        norm(n + ".norm1", cin), conv(n + ".conv1", cout, cin), lin(n + ".time_emb_proj", cout, ted)
        norm(n + ".norm2", cout), conv(n + ".conv2", cout, cout)
        if cin != cout:
            conv(n + ".conv_shortcut", cout, cin, 1)

    def attn(n, c, d):
        norm(n + ".norm", c), lin(n + ".proj_in", c, c), lin(n + ".proj_out", c, c)
        for t in range(d):
            b = f"{n}.transformer_blocks.{t}"
            for i in (1, 2, 3):
                norm(f"{b}.norm{i}", c)
            for a, kv in (("attn1", c), ("attn2", ctx)):
                lin(f"{b}.{a}.to_q", c, c, False), lin(f"{b}.{a}.to_k", c, kv, False), lin(f"{b}.{a}.to_v", c, kv, False)
                lin(f"{b}.{a}.to_out.0", c, c)
            lin(f"{b}.ff.net.0.proj", 8 * c, c), lin(f"{b}.ff.net.2", c, 4 * c)

    conv("conv_in", mc, 4)
    lin("time_embedding.linear_1", ted, mc), lin("time_embedding.linear_2", ted, ted)
    lin("add_embedding.linear_1", ted, adm), lin("add_embedding.linear_2", ted, ted)
    emb = (16, 32, 96, 256)
    conv("controlnet_cond_embedding.conv_in", emb[0], cfg.hint_channels)
    for i in range(3):
        conv(f"controlnet_cond_embedding.blocks.{2 * i}", emb[i], emb[i])
        conv(f"controlnet_cond_embedding.blocks.{2 * i + 1}", emb[i + 1], emb[i])
    conv("controlnet_cond_embedding.conv_out", mc, emb[3])
    zero = 0
    conv(f"controlnet_down_blocks.{zero}", mc, mc, 1)
    cin = mc
    for i, c in enumerate(chans):
        for j in range(cfg.num_res_blocks):
            resnet(f"down_blocks.{i}.resnets.{j}", cin, c)
            if depth[i]:
                attn(f"down_blocks.{i}.attentions.{j}", c, depth[i])
            cin = c
            zero += 1
            conv(f"controlnet_down_blocks.{zero}", c, c, 1)
        if i + 1 < len(chans):
            conv(f"down_blocks.{i}.downsamplers.0.conv", c, c)
            zero += 1
            conv(f"controlnet_down_blocks.{zero}", c, c, 1)
    top = chans[-1]
    resnet("mid_block.resnets.0", top, top), attn("mid_block.attentions.0", top, cfg.transformer_depth[-1]), resnet("mid_block.resnets.1", top, top)
    conv("controlnet_mid_block", top, top, 1)
    return s


def test_diffusers_controlnet_lands_on_the_cldm_parameter_tree():
    cfg = U.UNetConfig()
    src = _diffusers_controlnet_schema(cfg)
    sd = {k: torch.empty(shp, device="meta") for k, shp in src.items()}
    assert CK.is_diffusers_controlnet(sd)
    out = CK.controlnet_from_diffusers(sd)
    want = U.param_schema(cfg, controlnet=True)
    assert set(out) == set(want), (sorted(set(out) - set(want))[:5], sorted(set(want) - set(out))[:5])
    bad = {k: (tuple(out[k].shape), tuple(want[k])) for k in want if tuple(out[k].shape) != tuple(want[k])}
    assert not bad, dict(list(bad.items())[:5])
    assert len(out) == len(src) == 844
    with pytest.raises(KeyError, match="unexpected diffusers"):
        CK.controlnet_from_diffusers({"up_blocks.0.resnets.0.conv1.weight": torch.empty(1)})


def test_open_clip_text_tower_loads_into_transformers_and_computes_the_same():
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    hcfg = CLIPTextConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, vocab_size=100,
                          max_position_embeddings=77, projection_dim=48, hidden_act="gelu")
    torch.manual_seed(0)
    hf = CLIPTextModelWithProjection(hcfg).eval()
    ref = {k: v for k, v in hf.state_dict().items() if "position_ids" not in k}
    # the same weights the way open_clip names and stores them (written out by hand, independently of the converter)
    oc = {"token_embedding.weight": ref["text_model.embeddings.token_embedding.weight"],
          "positional_embedding": ref["text_model.embeddings.position_embedding.weight"],
          "ln_final.weight": ref["text_model.final_layer_norm.weight"], "ln_final.bias": ref["text_model.final_layer_norm.bias"],
          "text_projection": ref["text_projection.weight"].t().contiguous(), "logit_scale": torch.tensor(4.6)}
    for i in range(2):
        p, q = f"text_model.encoder.layers.{i}.", f"transformer.resblocks.{i}."
        oc[q + "attn.in_proj_weight"] = torch.cat([ref[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        oc[q + "attn.in_proj_bias"] = torch.cat([ref[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                     ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            oc[q + a + ".weight"], oc[q + a + ".bias"] = ref[p + b + ".weight"], ref[p + b + ".bias"]
    conv = CK.open_clip_text_to_hf(oc)
    assert set(conv) == set(ref) and all(torch.equal(conv[k], ref[k]) for k in ref)
    fresh = CLIPTextModelWithProjection(hcfg).eval()
    fresh.load_state_dict(conv, strict=False)
    ids = torch.randint(0, 99, (2, 77))
    ids[:, -1] = 99                                                       # EOT = the largest id (argmax pooling)
    with torch.no_grad():
        a, b = hf(ids), fresh(ids)
    assert torch.equal(a.text_embeds, b.text_embeds) and torch.equal(a.last_hidden_state, b.last_hidden_state)
    with pytest.raises(KeyError):
        CK.open_clip_text_to_hf({"visual.conv1.weight": torch.zeros(1)})


def test_single_file_sdxl_checkpoint_splits_by_prefix_and_round_trips_through_safetensors(tmp_path):
    from safetensors.torch import save_file
    sd = {CK.UNET_PREFIX + "input_blocks.0.0.weight": torch.randn(4, 4, 3, 3), CK.UNET_PREFIX + "out.2.bias": torch.randn(4),
          CK.VAE_PREFIX + "encoder.conv_in.weight": torch.randn(2, 3, 3, 3),
          CK.CLIP_L_PREFIX + "text_model.final_layer_norm.weight": torch.randn(8),
          CK.CLIP_L_PREFIX + "text_model.embeddings.position_ids": torch.arange(77)[None],
          CK.CLIP_G_PREFIX + "ln_final.weight": torch.randn(8), CK.CLIP_G_PREFIX + "text_projection": torch.randn(8, 6),
          "model_ema.decay": torch.tensor(0.999)}
    path = tmp_path / "sdxl.safetensors"
    save_file({k: v.contiguous() for k, v in sd.items()}, str(path))
    parts = CK.load_sdxl(path)
    assert set(parts) == {"unet", "vae", "clip_l", "clip_g"}
    assert set(parts["unet"]) == {"input_blocks.0.0.weight", "out.2.bias"} and set(parts["vae"]) == {"encoder.conv_in.weight"}
    assert set(parts["clip_l"]) == {"text_model.final_layer_norm.weight"}
    assert parts["clip_g"]["text_projection.weight"].shape == (6, 8)
    assert torch.equal(parts["clip_g"]["text_projection.weight"], sd[CK.CLIP_G_PREFIX + "text_projection"].t())
    with pytest.raises(KeyError, match="no unet"):
        CK.split_sdxl_checkpoint({"first_stage_model.x": torch.zeros(1)})
    # a ControlNet file in A1111 naming (control_model. prefix, cldm tree) passes through
    cn = tmp_path / "cn.safetensors"
    save_file({"control_model.zero_convs.0.0.weight": torch.zeros(2, 2, 1, 1)}, str(cn))
    assert set(CK.load_controlnet(cn)) == {"zero_convs.0.0.weight"}
