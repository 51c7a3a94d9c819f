"""CPU tests of the denoiser's host logic: parameter schema vs the oracle's modules (upstream naming), sampler
schedule scalars, argument validation.  No GPU compute."""
import math

import pytest
import torch

from oracle import sdxl_ref as R
from signerf_b200 import unet as U


@pytest.mark.parametrize("cfg", [R.tiny_config(), R.tiny_config(transformer_depth=(1, 2, 3), num_res_blocks=1)])
def test_schema_matches_oracle_state_dict(cfg):
    ref_unet, ref_ctrl = R.make_models(cfg)
    ucfg = U.UNetConfig(**cfg.__dict__)
    for ref, ctrl in ((ref_unet, False), (ref_ctrl, True)):
        sd, schema = ref.state_dict(), U.param_schema(ucfg, ctrl)
        assert list(schema) and set(schema) == set(sd)
        assert all(tuple(sd[k].shape) == tuple(v) for k, v in schema.items())


def test_sdxl_base_parameter_count():
    n_unet = sum(math.prod(s) for s in U.param_schema(U.UNetConfig()).values())
    n_ctrl = sum(math.prod(s) for s in U.param_schema(U.UNetConfig(), True).values())
    assert n_unet == 2_567_463_684          # sd_xl_base_1.0 UNet: 2.57 B parameters
    assert n_ctrl == 1_251_014_160 or abs(n_ctrl - 1.251e9) < 1e6


def test_sampler_scalars_match_oracle():
    sig = U.img2img_sigmas(20, 0.9)
    ref = R.img2img_schedule(20, 0.9)
    assert len(sig) == 20 and torch.allclose(torch.tensor(sig), ref, rtol=1e-6)
    table = R.sdxl_sigmas()
    for s in (sig[0], sig[5], sig[-2], 0.5, 14.0):
        assert abs(U.sigma_to_t(s) - float(R.sigma_to_t(table, torch.tensor([s]))[0])) < 1e-3
    for a, b in ((sig[0], sig[1]), (sig[-3], sig[-2]), (sig[-2], 0.0)):
        assert U.ancestral_step(a, b) == pytest.approx(R.ancestral_step(a, b))


def test_random_weights_follow_schema_and_are_deterministic():
    ucfg = U.UNetConfig(**R.tiny_config().__dict__)
    schema = U.param_schema(ucfg)
    w1, w2 = U.RandomWeights(schema, 0, "cpu"), U.RandomWeights(schema, 0, "cpu")
    for n in list(schema)[:40]:
        assert tuple(w1[n].shape) == tuple(schema[n]) and torch.equal(w1[n], w2[n])
    assert torch.equal(w1["out.0.weight"], torch.ones(ucfg.model_channels))


def test_missing_or_misshapen_weights_are_rejected():
    ucfg = U.UNetConfig(**R.tiny_config().__dict__)
    sd = dict(U.RandomWeights(U.param_schema(ucfg), 0, "cpu").items())
    bad = dict(sd)
    bad.pop("out.2.bias")
    with pytest.raises(KeyError):
        U.SDXLUNetB200(ucfg, bad, "cpu")
    bad = dict(sd)
    bad["out.2.bias"] = torch.zeros(5)
    with pytest.raises(ValueError):
        U.SDXLUNetB200(ucfg, bad, "cpu")


def test_sigma_table_known_answers():
    """Published constants of the SD / SDXL scaled-linear schedule (beta 0.00085 .. 0.012, 1000 steps) as k-diffusion's
    DiscreteSchedule reports them in A1111: sigma_min 0.0292, sigma_max 14.6146; ancestral step: up^2 + down^2 = to^2."""
    for table in (U.sdxl_sigmas(), R.sdxl_sigmas()):
        assert len(table) == 1000
        assert float(table[0]) == pytest.approx(0.029167, abs=2e-5) and float(table[-1]) == pytest.approx(14.614642, abs=2e-4)
        assert bool((table[1:] > table[:-1]).all())
    sig = U.img2img_sigmas(20, 0.9)
    assert sig[-1] == 0.0 and all(a > b for a, b in zip(sig[:-1], sig[1:])) and sig[0] < 14.614642
    for a, b in zip(sig[:-2], sig[1:-1]):
        down, up = U.ancestral_step(a, b)
        assert down ** 2 + up ** 2 == pytest.approx(b ** 2, rel=1e-9) and 0 < up <= b
    assert U.ancestral_step(sig[-2], 0.0) == (0.0, 0.0)
    assert U.sigma_to_t(float(U.sdxl_sigmas()[500])) == pytest.approx(500.0, abs=1e-3)
