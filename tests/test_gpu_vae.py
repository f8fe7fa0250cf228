"""GPU parity of the VAE decoder on the sm_100a kernels: `B200VaeDecoder.decode` against the fp32
oracle restatement of diffusers' `AutoencoderKL.decode` (oracle/vae_ref.py) on the same random-init weights, and the row
softmax + GEMM formulation of the head_dim-512 mid-block attention against torch SDPA."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2          # bf16 activations through ~30 conv / norm layers against fp32, like the UNet's per-call bar


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_and_lib():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the hot path has no CPU fallback")
    from audioldm_with_lora_b200 import _lib
    _lib.load()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,cols", [(1000, 4000), (37, 72), (5, 8192), (4000, 400)])
def test_softmax_rows(rows, cols):
    from audioldm_with_lora_b200 import ops
    g = torch.Generator().manual_seed(1)
    pad = (cols + 63) // 64 * 64
    s = torch.full((rows, pad), float("nan"), device=DEV)
    s[:, :cols] = (torch.randn(rows, cols, generator=g) * 6).to(DEV)
    p = torch.full((rows, pad), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.softmax_rows(s, rows, cols, pad, p, scale=0.37)
    ref = F.softmax(s[:, :cols].double() * 0.37, dim=-1)
    assert rel(p[:, :cols], ref) < 5e-3
    assert (p[:, cols:] == 0).all()


@pytest.fixture(scope="module")
def vae_pair():
    from audioldm_with_lora_b200 import tail
    from audioldm_with_lora_b200.vae import from_torch_decoder
    vae = tail.random_vae_decoder(7, std=0.05)
    return vae, from_torch_decoder(vae, DEV)


@pytest.mark.parametrize("nb,h", [(2, 25), (1, 250), (3, 63)])
def test_vae_decode_matches_oracle(vae_pair, nb, h):
    from oracle import vae_ref
    vae, dec = vae_pair
    g = torch.Generator().manual_seed(5)
    z = torch.randn(nb, 8, h, 16, generator=g)
    vsd = {k: v.detach().float().to(DEV) for k, v in vae.state_dict().items()}
    with torch.no_grad():
        ref = vae_ref.vae_decode(vsd, z.to(DEV))
    got = dec.decode(z.to(DEV))
    assert got.shape == (nb, 1, 4 * h, 64) and got.dtype == torch.float32 and torch.isfinite(got).all()
    assert rel(got, ref) < TOL
    # deterministic, and a sample's mel does not depend on its batch neighbours
    assert torch.equal(got, dec.decode(z.to(DEV)))
    if nb > 1:
        assert rel(dec.decode(z[:1].to(DEV)), got[:1]) < TOL


def test_pipeline_tail_with_b200_vae_matches_torch_tail():
    """`AudioLDMPipeline` re-hosts a torch VAE decoder on the kernels by default; b200_vae=False keeps the reference path.
    Same latents through both tails (bf16 vocoder in both): waveforms agree to bf16 noise."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import mel, synthetic, tail
    from audioldm_with_lora_b200.vae import B200VaeDecoder
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=DEV)
    voc = tail.build_vocoder(0)
    a = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), vae=tail.random_vae_decoder(7), vocoder=voc)
    b = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), vae=tail.random_vae_decoder(7), vocoder=voc, b200_vae=False)
    assert isinstance(a.vae, B200VaeDecoder) and not isinstance(b.vae, B200VaeDecoder)
    lat = synthetic.initial_latents(2, 64).to(DEV)
    wa, wb = a.latents_to_waveform(lat).float().cpu(), b.latents_to_waveform(lat).float().cpu()
    assert wa.shape == wb.shape and torch.isfinite(wa).all()
    gain = 0.5 / wb.abs().max().clamp_min(1e-30)
    assert mel.logmel_l1(wa * gain, wb * gain) < 0.1
    wa2 = a.latents_to_waveform(lat).float().cpu()          # second call replays the captured tail graph
    assert torch.equal(wa, wa2)


# ------------------------------------------------------------------------------------------ encoder
@pytest.mark.parametrize("nb,h,w,c,co", [(2, 64, 64, 128, 128), (1, 37, 32, 256, 256), (3, 9, 16, 64, 128), (1, 1024, 64, 128, 128)])
def test_conv3x3_s2_pad01_matches_torch(nb, h, w, c, co):
    """Downsample2D(padding=0) of the VAE encoder: F.pad(x, (0, 1, 0, 1)) + conv k3 s2 p0, as strided TMA boxes."""
    from audioldm_with_lora_b200 import ops, packing
    g = torch.Generator().manual_seed(h * 7 + w)
    x = torch.randn(nb, c, h, w, generator=g).to(torch.bfloat16)
    wt = torch.randn(co, c, 3, 3, generator=g) / (9 * c) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    ho, wo = (h - 2) // 2 + 1, (w - 2) // 2 + 1
    bn = ops.choose_tiling(co, ops.num_m_tiles(nb, ho, wo), 9 * c // 64, allow_split=False)[0]
    pw = packing.pack([packing.conv3x3_to_k(wt)], b, bn, 9, c, device=DEV)
    xin = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.full((nb, ho, wo, co), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.conv3x3_s2_pad01(pw, xin, nb, h, w, out)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), wt.to(torch.bfloat16).float(), b, stride=2)
    assert ref.shape == (nb, co, ho, wo) and torch.isfinite(out).all()
    assert rel(out.permute(0, 3, 1, 2), ref) < 5e-3


@pytest.fixture(scope="module")
def enc_pair():
    from audioldm_with_lora_b200 import synthetic
    from audioldm_with_lora_b200.vae import from_torch_encoder
    from oracle import vae_ref
    sd = synthetic.random_state_dict_from_shapes(vae_ref.vae_encoder_param_shapes(), seed=11, std=0.05)
    return sd, from_torch_encoder(sd, DEV)


@pytest.mark.parametrize("nb,t", [(2, 64), (1, 1024), (3, 100)])
def test_vae_encode_matches_oracle(enc_pair, nb, t):
    """`vae.encode(log_mel).latent_dist` (train_audioldm_lora.py:495) on the kernels vs the fp32 oracle on the same weights."""
    from oracle import vae_ref
    sd, enc = enc_pair
    g = torch.Generator().manual_seed(nb * 100 + t)
    mel = (torch.randn(nb, 1, t, 64, generator=g) * 2.0 - 4.0)
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        mean, logvar = vae_ref.vae_encode(dsd, mel.to(DEV))
    d = enc.encode(mel.to(DEV)).latent_dist
    assert d.mean.shape == mean.shape == (nb, 8, t // 4, 16) and d.mean.dtype == torch.float32
    assert torch.isfinite(d.mean).all() and torch.isfinite(d.logvar).all()
    assert rel(d.mean, mean) < TOL and rel(d.logvar, logvar) < TOL
    assert torch.equal(d.mean, enc.encode(mel.to(DEV)).latent_dist.mean)          # deterministic
    lat = d.sample(torch.Generator(device=DEV).manual_seed(3)) * enc.config.scaling_factor
    assert lat.shape == mean.shape and torch.isfinite(lat).all()
