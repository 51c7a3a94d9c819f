"""CPU: host logic of signerf_b200/vae.py — the independent parameter schema against the oracle's state_dict."""
import pytest
import torch

from oracle import vae_ref as VR
from signerf_b200 import vae as V


@pytest.mark.parametrize("cfg", [VR.VAEConfig(), VR.tiny_vae_config()])
def test_vae_schema_matches_oracle_state_dict(cfg):
    schema = V.vae_param_schema(V.VAEConfig(**cfg.__dict__))
    sd = VR.AutoencoderKL(cfg).state_dict()
    assert list(schema.keys()) != [] and set(schema.keys()) == set(sd.keys())
    for n, shp in schema.items():
        assert tuple(sd[n].shape) == tuple(shp), n


def test_sdxl_vae_parameter_count():
    import math
    assert sum(math.prod(s) for s in V.vae_param_schema(V.VAEConfig()).values()) == 83_653_863


def test_vae_rejects_cpu_and_bad_weights():
    cfg = VR.tiny_vae_config()
    sd = VR.make_vae(cfg).state_dict()
    bad = dict(sd)
    bad.pop("encoder.conv_in.weight")
    with pytest.raises(KeyError):
        V.VAEB200(V.VAEConfig(**cfg.__dict__), bad, "cpu")
    bad = dict(sd)
    bad["decoder.conv_out.weight"] = torch.zeros(3, 7, 3, 3)
    with pytest.raises(ValueError):
        V.VAEB200(V.VAEConfig(**cfg.__dict__), bad, "cpu")
