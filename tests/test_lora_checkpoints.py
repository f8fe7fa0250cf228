"""CPU tests of the LoRA checkpoint writers and the opt-in pre-merge: the files the
reference's tooling writes / reads (accelerate `model.safetensors` with full peft keys -- train_audioldm_lora.py:575,
generate_audio.py:32; diffusers `pytorch_lora_weights.safetensors` -- train:577-579, app.py:11) round-trip through
this package, and W' = W + (alpha / r) * scale * B A reproduces the unmerged forward of the fp32 oracle."""
import pytest
import torch

import audioldm_with_lora_b200 as b2
from audioldm_with_lora_b200 import engine as engine_mod, synthetic
from audioldm_with_lora_b200.arch import UNetConfig
from audioldm_with_lora_b200.lora import _CKPT_FILES, lora_state_dict_as
from oracle import unet_ref

TINY = UNetConfig("tiny", (64, 128, 192, 256))
TINY_SPEC = unet_ref.UNetSpec(block_out_channels=TINY.block_out_channels, time_proj_dim=64)


@pytest.fixture(scope="module")
def weights():
    sd = synthetic.random_unet_state_dict(TINY, seed=0)
    ad = b2.parse_lora_state_dict(synthetic.random_lora_state_dict(TINY, 8, fmt="peft"), alpha=4.0)
    return sd, ad


@pytest.mark.parametrize("fmt", ["peft_full", "peft", "diffusers"])
def test_checkpoint_roundtrip_every_format(tmp_path, weights, fmt):
    _, ad = weights
    path = b2.save_lora_checkpoint(ad, tmp_path / fmt, fmt=fmt)
    assert path.name == _CKPT_FILES[fmt] and path.exists()
    keys = set(lora_state_dict_as(ad, fmt))
    marker = {"peft_full": ".lora_A.default.weight", "peft": ".lora_A.weight", "diffusers": ".lora.down.weight"}[fmt]
    assert sum(marker in k for k in keys) == len(ad) and len(keys) == 2 * len(ad)
    back = b2.load_lora_checkpoint(tmp_path / fmt)              # directory or file, alpha from the metadata
    assert set(back) == set(ad)
    for p, e in ad.items():
        assert torch.equal(back[p].A, e.A) and torch.equal(back[p].B, e.B) and back[p].alpha == e.alpha
    assert b2.load_lora_checkpoint(path, alpha=2.0)[next(iter(ad))].alpha == 2.0


def test_bf16_checkpoint_and_unknown_format(tmp_path, weights):
    _, ad = weights
    path = b2.save_lora_checkpoint(ad, tmp_path, fmt="diffusers", dtype=torch.bfloat16)
    back = b2.load_lora_checkpoint(path)
    p = next(iter(ad))
    assert torch.equal(back[p].A, ad[p].A.bfloat16().float())    # stored in bf16, adapters are fp32 on load
    with pytest.raises(ValueError):
        b2.save_lora_checkpoint(ad, tmp_path, fmt="kohya")


def test_merge_equals_unmerged_projection_and_unmerges(weights):
    sd, ad = weights
    merged = b2.merge_lora_into_state_dict(sd, ad, scale=0.5)
    touched = {p + ".weight" for p in ad}
    assert all(merged[k] is sd[k] for k in sd if k not in touched)           # untouched tensors are shared
    g = torch.Generator().manual_seed(1)
    for p, e in list(ad.items())[:6]:
        w, w2 = sd[p + ".weight"].double(), merged[p + ".weight"].double()
        x = torch.randn(5, w.shape[1], generator=g, dtype=torch.float64)
        want = x @ w.T + 0.5 * (e.alpha / e.A.shape[0]) * (x @ e.A.double().T) @ e.B.double().T
        assert torch.allclose(x @ w2.T, want, rtol=1e-5, atol=1e-5)
    again = b2.merge_lora_into_state_dict(merged, ad, scale=0.5, sign=-1.0)
    assert all(torch.allclose(again[k], sd[k], atol=1e-6) for k in touched)
    with pytest.raises(KeyError):
        b2.merge_lora_into_state_dict({}, ad)


def test_merged_model_matches_unmerged_oracle_forward(weights):
    """The adapter-free UNet on merged weights == the unmerged LoRA forward, in the fp32 oracle (exact up to fp32
    rounding), i.e. the merge is the algebra peft's merge_and_unload promises."""
    sd, ad = weights
    lora = unet_ref.LoraSet({k: (e.A, e.B, e.alpha) for k, e in ad.items()})
    x = synthetic.initial_latents(1, 16)
    pos, _ = synthetic.clap_embeddings(1)
    with torch.no_grad():
        unmerged = unet_ref.unet_forward(sd, TINY_SPEC, x, 321, pos, lora=lora)
        merged = unet_ref.unet_forward(b2.merge_lora_into_state_dict(sd, ad), TINY_SPEC, x, 321, pos)
        base = unet_ref.unet_forward(sd, TINY_SPEC, x, 321, pos)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(merged, unmerged) < 1e-5 < rel(base, unmerged)


def test_model_save_attn_procs_and_merged_state_dict(tmp_path, weights, fake_kernels):
    """`unet.save_attn_procs(dir)` -> `other.load_attn_procs(dir)` (app.py:11) and the merged, adapter-free model on the
    engine (kernels emulated on CPU by tests/fake_ops.py) against the unmerged engine forward."""
    sd, ad = weights
    unet = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    unet.load_lora_state_dict(b2.to_peft_state_dict(ad), alpha=4.0)
    f = unet.save_attn_procs(tmp_path)
    assert f.name == "pytorch_lora_weights.safetensors"
    other = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    other.load_attn_procs(str(tmp_path), network_alpha=4.0)
    assert set(other.engine.lora) == set(ad)
    p = next(iter(ad))
    assert torch.equal(other._adapters[p].B, ad[p].B) and other._adapters[p].alpha == 4.0
    # without network_alpha the alpha recorded by save_attn_procs is honoured (alpha != r here: rank 4... alpha 4.0 vs r)
    r = ad[p].A.shape[0]
    unet.load_lora_state_dict(b2.to_peft_state_dict(ad), alpha=2.0 * r)
    unet.save_attn_procs(tmp_path)
    third = b2.UNet2DConditionModel(TINY, sd, device="cpu")
    third.load_attn_procs(str(tmp_path))                                   # directory
    assert third._adapters[p].alpha == 2.0 * r and third.engine.lora[p].scaling == 2.0
    third.load_attn_procs(str(tmp_path / "pytorch_lora_weights.safetensors"))   # file
    assert third._adapters[p].alpha == 2.0 * r
    x = synthetic.initial_latents(1, 16)
    pos, _ = synthetic.clap_embeddings(1)
    unmerged = unet.engine.forward(x, 77, pos)
    plain = engine_mod.UNetEngine(TINY, unet.merged_state_dict(), "cpu").forward(x, 77, pos)
    rel = ((plain.float() - unmerged.float()).norm() / unmerged.float().norm()).item()
    assert rel < 2e-2                       # bf16 rounding of W' vs of W and the rank-r term separately
