"""Known-answer tests that pin the ORACLE (the reference has no tests or golden vectors for this
path -- SURVEY.md section 4 -- so these are the independent pins: parameter counts, schedule closed forms,
algebraic identities, and the committed golden fixtures)."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from audioldm_with_lora_b200 import synthetic
from audioldm_with_lora_b200.arch import CONFIGS, UNetConfig, attention_paths, unet_param_shapes
from audioldm_with_lora_b200.lora import parse_lora_state_dict
from oracle import pipeline_ref, unet_ref, vae_ref
from oracle.ddim_ref import DDIMRef, PNDMRef

GOLD = Path(__file__).resolve().parent / "golden"
TINY = UNetConfig("tiny", (64, 128, 192, 256))
TINY_SPEC = unet_ref.UNetSpec(block_out_channels=TINY.block_out_channels, time_proj_dim=64)


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_param_counts_match_published_model_sizes():
    # 739.1 M is the AudioLDM paper's published size of the L UNet; S is 185.0 M
    assert abs(unet_ref.count_params(unet_ref.ARCH_S) / 1e6 - 185.0) < 0.05
    assert abs(unet_ref.count_params(unet_ref.ARCH_L) / 1e6 - 739.1) < 0.05


@pytest.mark.parametrize("arch", ["S", "L"])
def test_two_independent_enumerations_agree(arch):
    a = unet_ref.param_shapes(unet_ref.ARCHS[arch])
    b = unet_param_shapes(CONFIGS[arch])
    assert a == b


def test_attention_modules_and_lora_param_counts():
    paths = attention_paths(CONFIGS["S"])
    assert len(paths) == 32 and paths == unet_ref.attention_module_names(unet_ref.ARCH_S) or sorted(paths) == sorted(unet_ref.attention_module_names(unet_ref.ARCH_S))
    n = lambda sd: sum(v.numel() for v in sd.values())
    assert n(synthetic.random_lora_state_dict(CONFIGS["S"], 8)) == 901_120
    assert n(synthetic.random_lora_state_dict(CONFIGS["S"], 2, targets=("to_q", "to_v"))) == 112_640
    assert n(synthetic.random_lora_state_dict(CONFIGS["S"], 16)) == 1_802_240
    assert n(synthetic.random_lora_state_dict(CONFIGS["L"], 32)) == 7_208_960


def test_ddim_timesteps_and_schedule_closed_forms():
    s = DDIMRef()
    assert s.set_timesteps(200).tolist() == list(range(996, 0, -5))
    assert s.set_timesteps(10).tolist() == list(range(901, 0, -100))
    assert s.set_timesteps(50).tolist() == list(range(981, 0, -20))
    betas = torch.linspace(0.0015 ** 0.5, 0.0195 ** 0.5, 1000, dtype=torch.float64) ** 2
    ac = torch.cumprod(1 - betas, 0)
    assert torch.allclose(s.alphas_cumprod.double(), ac, rtol=1e-5)
    assert float(s.final_alpha_cumprod) == float(s.alphas_cumprod[0]) and abs(float(s.alphas_cumprod[0]) - 0.9985) < 1e-6
    g = np.load(GOLD / "ddim_schedule.npz")
    assert np.array_equal(g["t200"], s.set_timesteps(200).numpy())
    assert np.array_equal(g["alphas_cumprod"], s.alphas_cumprod.numpy())
    assert np.array_equal(g["plms10"], PNDMRef().set_timesteps(10).numpy())


def test_ddim_step_is_affine_and_invertible_at_eta0():
    s = DDIMRef(); s.set_timesteps(10)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 8, 5, 16, generator=g), torch.randn(2, 8, 5, 16, generator=g)
    t = 501
    xp = s.step(e, t, x)
    a_t, a_p = float(s.alphas_cumprod[t]), float(s.alphas_cumprod[t - 100])
    c1 = math.sqrt(a_p / a_t)
    c2 = math.sqrt(1 - a_p) - math.sqrt(a_p * (1 - a_t) / a_t)
    assert torch.allclose(xp, c1 * x + c2 * e, atol=1e-5)
    assert torch.allclose((xp - c2 * e) / c1, x, atol=1e-5)              # eta = 0: deterministic inverse
    xl = s.step(e, 1, x)                                                 # last step uses final_alpha_cumprod
    a_t, a_p = float(s.alphas_cumprod[1]), float(s.alphas_cumprod[0])
    assert torch.allclose(xl, math.sqrt(a_p / a_t) * x + (math.sqrt(1 - a_p) - math.sqrt(a_p * (1 - a_t) / a_t)) * e, atol=1e-5)


def test_add_noise_matches_definition():
    s = DDIMRef()
    g = torch.Generator().manual_seed(1)
    x0, n = torch.randn(3, 8, 4, 16, generator=g), torch.randn(3, 8, 4, 16, generator=g)
    t = torch.tensor([0, 500, 999])
    out = s.add_noise(x0, n, t)
    for i in range(3):
        a = s.alphas_cumprod[t[i]]
        assert torch.allclose(out[i], a.sqrt() * x0[i] + (1 - a).sqrt() * n[i], atol=1e-6)


def test_timestep_embedding_is_cos_then_sin():
    e = unet_ref.timestep_embedding(torch.tensor([0.0, 3.0]), 128)
    assert torch.allclose(e[0, :64], torch.ones(64)) and torch.allclose(e[0, 64:], torch.zeros(64))
    assert abs(e[1, 0].item() - math.cos(3.0)) < 1e-6 and abs(e[1, 64].item() - math.sin(3.0)) < 1e-6


@pytest.fixture(scope="module")
def tiny():
    sd = synthetic.random_unet_state_dict(TINY, seed=0)
    ad = parse_lora_state_dict(synthetic.random_lora_state_dict(TINY, 4, fmt="peft"))
    x = synthetic.initial_latents(2, 12)
    pos, neg = synthetic.clap_embeddings(2)
    return sd, ad, x, pos, neg


def test_lora_identities(tiny):
    sd, ad, x, pos, _ = tiny
    with torch.no_grad():
        base = unet_ref.unet_forward(sd, TINY_SPEC, x, 500, pos)
        # B = 0 (peft init) == base model
        zero = unet_ref.LoraSet({k: (e.A, torch.zeros_like(e.B), e.alpha) for k, e in ad.items()})
        assert torch.equal(unet_ref.unet_forward(sd, TINY_SPEC, x, 500, pos, lora=zero), base)
        # unmerged == merged (W + s B A) up to fp32 rounding, and differs from base
        lora = unet_ref.LoraSet({k: (e.A, e.B, 2 * e.alpha) for k, e in ad.items()})
        un = unet_ref.unet_forward(sd, TINY_SPEC, x, 500, pos, lora=lora)
        merged = dict(sd)
        for k, e in ad.items():
            merged[k + ".weight"] = sd[k + ".weight"] + 2.0 * (e.B @ e.A)
        me = unet_ref.unet_forward(merged, TINY_SPEC, x, 500, pos)
        assert rel(un, me) < 1e-5 and rel(un, base) > 1e-4
        # runtime scale multiplies the branch: scale 0 == base
        off = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()}, scale=0.0)
        assert torch.equal(unet_ref.unet_forward(sd, TINY_SPEC, x, 500, pos, lora=off), base)


def test_cfg_guidance_one_is_conditional_only(tiny):
    sd, _, x, pos, neg = tiny
    with torch.no_grad():
        a = pipeline_ref.denoise_loop(sd, TINY_SPEC, pos, neg, x.clone(), 3, 1.0)           # g <= 1: no CFG at all
        # g slightly > 1 computes both branches; in the limit it equals the conditional branch
        b = pipeline_ref.denoise_loop(sd, TINY_SPEC, pos, neg, x.clone(), 3, 1.0 + 1e-7)
    assert rel(b, a) < 1e-5


def test_latent_height_rule():
    assert pipeline_ref.latent_height(10.0) == 250 and pipeline_ref.latent_height(5.0) == 125
    assert pipeline_ref.latent_height(5.12) == 128 and pipeline_ref.latent_height(30.0) == 750
    assert pipeline_ref.latent_height(4.0) == 100 and pipeline_ref.latent_height(0.05) == 2     # 5 -> 8 -> 2 (round up to x4)


def test_batch_and_timestep_vector_consistency(tiny):
    sd, ad, x, pos, _ = tiny
    lora = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})
    with torch.no_grad():
        both = unet_ref.unet_forward(sd, TINY_SPEC, x, torch.tensor([17, 903]), pos, lora=lora)
        one = unet_ref.unet_forward(sd, TINY_SPEC, x[1:], 903, pos[1:], lora=lora)
    assert rel(both[1:], one) < 1e-5


def test_vae_decoder_shapes():
    sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_decoder_param_shapes(), seed=7)
    with torch.no_grad():
        mel = vae_ref.vae_decode(sd, torch.randn(1, 8, 4, 16))
    assert tuple(mel.shape) == (1, 1, 16, 64)


@pytest.mark.slow
def test_oracle_reproduces_committed_golden_vectors():
    """Regression pin of the restatement (fixtures made by tests/golden/make_golden.py)."""
    cfg = CONFIGS["S"]
    sd = synthetic.random_unet_state_dict(cfg, seed=0)
    ad = parse_lora_state_dict(synthetic.random_lora_state_dict(cfg, 8, fmt="peft"))
    lora = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})
    g = np.load(GOLD / "unet_s_r8_b2_h24.npz")
    x = synthetic.initial_latents(2, 24)
    pos, _ = synthetic.clap_embeddings(2)
    with torch.no_grad():
        eps = unet_ref.unet_forward(sd, unet_ref.ARCH_S, x, 501, pos, lora=lora)
    assert rel(eps, torch.from_numpy(g["eps"])) < 1e-5
    assert rel(torch.from_numpy(g["eps"]), torch.from_numpy(g["eps_nolora"])) > 1e-4      # the LoRA branch is live
