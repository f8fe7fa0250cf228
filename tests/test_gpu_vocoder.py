"""GPU parity of the HiFi-GAN vocoder on the sm_100a kernels: `b200_conv1d` against torch's
conv1d / conv_transpose1d on the same bf16 inputs, and `B200HifiGan` against the reference's own vocoder code
(transformers SpeechT5HifiGan, fp32; /root/reference/script/train/train_audioldm_lora.py:371) on the same weights."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
KTOL = 5e-3         # single kernel vs torch fp32 on the same bf16 inputs (bf16 output rounding)
TOL = 3e-2          # whole vocoder: bf16 activations through ~50 sequential convolutions against fp32


@pytest.fixture(scope="module", autouse=True)
def _need_cuda_and_lib():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device: the hot path has no CPU fallback")
    from audioldm_with_lora_b200 import _lib
    _lib.load()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("c,co,k,d,nb,length,res,act", [
    (64, 64, 3, 1, 2, 300, False, 1.0),
    (128, 128, 7, 3, 3, 517, True, 0.1),
    (512, 512, 11, 5, 2, 1001, True, 0.1),
    (256, 256, 3, 5, 1, 129, False, 0.1),
    (64, 1024, 7, 1, 2, 250, False, 0.1),
])
def test_conv1d_matches_torch(c, co, k, d, nb, length, res, act):
    from audioldm_with_lora_b200 import ops, packing
    from audioldm_with_lora_b200.vocoder import _conv1d_to_k
    g = torch.Generator().manual_seed(k * 100 + d)
    w = torch.randn(co, c, k, generator=g) / (c * k) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    x = torch.randn(nb, length, c, generator=g).to(torch.bfloat16)
    r = (torch.randn(nb, length, co, generator=g) * 0.5).to(torch.bfloat16) if res else None
    m_tiles = nb * ((length + 127) // 128)
    bn = ops.choose_tiling(co, m_tiles, k * c // 64, allow_split=False)[0]
    pw = packing.pack([_conv1d_to_k(w, co)], b, bn, k, c, device=DEV)
    out = torch.full((nb, length, co), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.conv1d(pw, x.to(DEV), nb, length, out, dh0=-d * (k - 1) // 2, dh_step=d, residual=None if r is None else r.to(DEV),
               res_slope=0.1 if res else 1.0, act_slope=act)
    ref = F.conv1d(x.float().transpose(1, 2), w.to(torch.bfloat16).float(), b, dilation=d, padding=d * (k - 1) // 2).transpose(1, 2)
    if res:
        rf = r.float()
        ref = ref + torch.minimum(rf, rf * 10.0)             # the residual is stored post-LeakyReLU(0.1)
    ref = torch.maximum(ref, ref * act)
    assert torch.isfinite(out).all()
    assert rel(out, ref) < KTOL


@pytest.mark.parametrize("ci,co,k,s,nb,length", [(1024, 512, 16, 5, 2, 100), (512, 256, 16, 4, 1, 333), (256, 128, 8, 2, 2, 200),
                                                 (64, 64, 4, 2, 3, 257)])
def test_conv_transpose1d_by_phases_matches_torch(ci, co, k, s, nb, length):
    from audioldm_with_lora_b200 import ops, packing
    from audioldm_with_lora_b200.vocoder import _convT_phase_to_k
    g = torch.Generator().manual_seed(k + s)
    w = torch.randn(ci, co, k, generator=g) / (ci * k / s) ** 0.5
    b = torch.randn(co, generator=g) * 0.1
    x = torch.randn(nb, length, ci, generator=g).to(torch.bfloat16)
    pad = (k - s) // 2
    ref = F.conv_transpose1d(x.float().transpose(1, 2), w.to(torch.bfloat16).float(), b, stride=s, padding=pad).transpose(1, 2)
    l_out = ref.shape[1]
    y = torch.full((nb, l_out, co), float("nan"), dtype=torch.bfloat16, device=DEV)
    xd = x.to(DEV)
    for phi in range(s):
        a, bb = (phi + pad) % s, (phi + pad) // s
        ntaps = (k - a + s - 1) // s
        rows = (l_out - phi + s - 1) // s
        bn = ops.choose_tiling(co, nb * ((rows + 127) // 128), ntaps * ci // 64, allow_split=False)[0]
        pw = packing.pack([_convT_phase_to_k(w, s, a, ntaps, co)], b, bn, ntaps, ci, device=DEV)
        ops.conv1d(pw, xd, nb, length, y.view(-1)[phi * co:], dh0=bb, dh_step=-1, m_rows=rows, out_ld=s * co,
                   out_batch_stride=l_out * co)
    assert torch.isfinite(y).all()
    assert rel(y, ref) < KTOL


def test_lrelu_mean3_and_cast():
    from audioldm_with_lora_b200 import ops
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(3, 1000, 64, generator=g) for _ in range(3)]
    a = [F.leaky_relu(x, 0.1).to(torch.bfloat16).to(DEV) for x in xs]
    y = torch.empty_like(a[0])
    ops.lrelu_mean3(a[0], a[1], a[2], 0.1, 0.01, y)
    back = [torch.minimum(t.float(), t.float() * 10.0) for t in a]
    ref = F.leaky_relu((back[0] + back[1] + back[2]) / 3.0, 0.01)
    assert rel(y, ref) < KTOL
    f = torch.randn(4096, generator=g).to(DEV)
    o = torch.empty(4096, dtype=torch.bfloat16, device=DEV)
    assert torch.equal(ops.f32_to_bf16(f, o), f.to(torch.bfloat16))


@pytest.fixture(scope="module")
def voc_pair():
    from audioldm_with_lora_b200 import tail
    from audioldm_with_lora_b200.vocoder import from_torch_vocoder
    voc = tail.build_vocoder(3)
    with torch.no_grad():                 # random-init weights are tiny (std 0.01): scale up so every layer matters (x8 would saturate the final tanh)
        for p in voc.parameters():
            p.mul_(3.0)
    return voc.to(DEV), from_torch_vocoder(voc, DEV)


@pytest.mark.parametrize("nb,t", [(2, 100), (1, 1000), (3, 37)])
def test_b200_hifigan_matches_transformers(voc_pair, nb, t):
    voc, mine = voc_pair
    g = torch.Generator().manual_seed(nb * 1000 + t)
    mel = (torch.randn(nb, t, 64, generator=g) * 2.0 - 4.0).to(DEV)
    with torch.no_grad():
        ref = voc(mel)
    got = mine(mel)
    assert (ref.abs() > 0.999).float().mean() < 0.1, "reference output saturates the final tanh: the comparison would be vacuous"
    assert got.shape == ref.shape and got.dtype == torch.float32 and torch.isfinite(got).all()
    assert rel(got, ref) < TOL
    assert torch.equal(got, mine(mel))                       # deterministic
    if nb > 1:                                               # a clip does not depend on its batch neighbours
        assert rel(mine(mel[:1].contiguous()), got[:1]) < 1e-6


def test_pipeline_tail_with_b200_vocoder_matches_torch_tail():
    """`AudioLDMPipeline` re-hosts SpeechT5HifiGan on the kernels by default; b200_vocoder=False keeps the reference path.
    Same latents through both tails: waveforms agree to bf16 noise (log-mel L1, the north star's end-to-end metric)."""
    import audioldm_with_lora_b200 as b2
    from audioldm_with_lora_b200 import mel, synthetic, tail
    cfg = b2.CONFIGS["S"]
    unet = b2.UNet2DConditionModel(cfg, synthetic.random_unet_state_dict(cfg, seed=0), device=DEV)
    a = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), vae=tail.random_vae_decoder(7), vocoder=tail.build_vocoder(0))
    b = b2.AudioLDMPipeline(unet, b2.DDIMScheduler(), vae=tail.random_vae_decoder(7), vocoder=tail.build_vocoder(0),
                            b200_vocoder=False)
    assert isinstance(a.vocoder, b2.B200HifiGan) and not isinstance(b.vocoder, b2.B200HifiGan)
    lat = synthetic.initial_latents(2, 64).to(DEV)
    wa, wb = a.latents_to_waveform(lat).float().cpu(), b.latents_to_waveform(lat).float().cpu()
    assert wa.shape == wb.shape and torch.isfinite(wa).all()
    gain = 0.5 / wb.abs().max().clamp_min(1e-30)
    assert mel.logmel_l1(wa * gain, wb * gain) < 0.1
    assert torch.equal(wa, a.latents_to_waveform(lat).float().cpu())       # second call replays the captured tail graph
