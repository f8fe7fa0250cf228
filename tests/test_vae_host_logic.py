"""CPU tests of the VAE host logic: `B200VaeDecoder` / `B200VaeEncoder` with the kernels replaced by
tests/fake_ops.py against the fp32 oracle restatement of diffusers' AutoencoderKL (oracle/vae_ref.py) -- weight packing, the
post_quant_conv / quant_conv folds, the K-segment shortcuts, the GEMM + row-softmax mid-block attention and the
asymmetric-padding stride-2 convolutions of the encoder."""
import pytest
import torch

from audioldm_with_lora_b200 import synthetic
from audioldm_with_lora_b200.vae import B200VaeDecoder, B200VaeEncoder
from oracle import vae_ref


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_decoder_host_logic_matches_oracle(fake_kernels, monkeypatch):
    from audioldm_with_lora_b200 import vae as vae_mod
    sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_decoder_param_shapes(), seed=3, std=0.05)
    dec = B200VaeDecoder(sd, device="cpu")
    g = torch.Generator().manual_seed(0)
    z = torch.randn(2, 8, 4, 16, generator=g)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    got = dec.decode(z)
    ref = vae_ref.vae_decode(sd, z)
    assert got.shape == ref.shape == (2, 1, 16, 64)
    assert rel(got, ref) < 2e-2


def test_encoder_host_logic_matches_oracle(fake_kernels):
    sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_encoder_param_shapes(), seed=4, std=0.05)
    enc = B200VaeEncoder(sd, device="cpu")
    g = torch.Generator().manual_seed(1)
    mel = torch.randn(2, 1, 16, 64, generator=g) * 2.0 - 4.0
    out = enc.encode(mel)
    mean, logvar = vae_ref.vae_encode(sd, mel)
    d = out.latent_dist
    assert d.mean.shape == mean.shape == (2, 8, 4, 16)
    assert rel(d.mean, mean) < 2e-2 and rel(d.logvar, logvar) < 2e-2
    gen = torch.Generator().manual_seed(7)
    s1 = d.sample(gen)
    assert s1.shape == mean.shape and torch.isfinite(s1).all()
    assert torch.equal(d.mode(), d.mean)
    # mean + std * noise with the same generator state
    noise = torch.randn(mean.shape, generator=torch.Generator().manual_seed(7))
    assert torch.allclose(s1, d.mean + d.std * noise)


def test_encoder_odd_length_uses_pad01_geometry(fake_kernels):
    """F.pad (0,1,0,1) + k3 s2: out = (n - 2) // 2 + 1 per level, also for lengths that are not multiples of 4."""
    sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_encoder_param_shapes(), seed=5, std=0.05)
    enc = B200VaeEncoder(sd, device="cpu")
    mel = torch.randn(1, 1, 22, 64, generator=torch.Generator().manual_seed(2))
    mean, _ = vae_ref.vae_encode(sd, mel)
    got = enc.encode(mel).latent_dist.mean
    assert got.shape == mean.shape and rel(got, mean) < 2e-2
