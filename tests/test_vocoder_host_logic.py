"""CPU test of the HiFi-GAN host logic: weight packing, the per-phase decomposition of the
transposed convolutions, the post-activation storage convention and the residual-block schedule of `B200HifiGan`, with the
kernels replaced by tests/fake_ops.py, against the reference's OWN vocoder code: transformers' SpeechT5HifiGan in fp32
(/root/reference/script/train/train_audioldm_lora.py:371 loads exactly this class)."""
import pytest
import torch

from audioldm_with_lora_b200 import tail
from audioldm_with_lora_b200.vocoder import B200HifiGan, _convT_phase_to_k


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_transposed_conv_phase_decomposition_is_exact():
    """y[s q + phi] = sum_j x[q + b - j] W[:, :, s j + a]: the identity the per-phase launches rest on, in fp64."""
    torch.manual_seed(0)
    for k, s in [(16, 5), (16, 4), (8, 2), (4, 2)]:
        ci, co, length = 6, 5, 23
        pad = (k - s) // 2
        w = torch.randn(ci, co, k, dtype=torch.float64)
        x = torch.randn(1, ci, length, dtype=torch.float64)
        ref = torch.nn.functional.conv_transpose1d(x, w, stride=s, padding=pad)[0].T        # [L_out, co]
        l_out = ref.shape[0]
        assert l_out == (length - 1) * s - 2 * pad + k          # = s L, or s L + 1 when k - s is odd (the first stage)
        xt = x[0].T                                                                          # [L, ci]
        for phi in range(s):
            a, b = (phi + pad) % s, (phi + pad) // s
            ntaps = (k - a + s - 1) // s
            wk = _convT_phase_to_k(w, s, a, ntaps, 8).view(8, ntaps, 64)[:co, :, :ci]
            rows = (l_out - phi + s - 1) // s
            for q in range(rows):
                acc = torch.zeros(co, dtype=torch.float64)
                for j in range(ntaps):
                    src = q + b - j
                    if 0 <= src < length:
                        acc += wk[:, j] @ xt[src]
                assert torch.allclose(acc, ref[s * q + phi], atol=1e-10)


@pytest.mark.parametrize("nb,t", [(2, 12), (1, 7)])
def test_b200_hifigan_host_logic_matches_transformers(fake_kernels, nb, t):
    voc = tail.build_vocoder(3)
    with torch.no_grad():                 # random-init weights are tiny (std 0.01): scale up so every layer matters (x8 would saturate the final tanh)
        for p in voc.parameters():
            p.mul_(3.0)
    mine = B200HifiGan(voc, device="cpu")
    g = torch.Generator().manual_seed(1)
    mel = torch.randn(nb, t, 64, generator=g) * 2.0 - 4.0
    with torch.no_grad():
        ref = voc(mel)
    got = mine(mel)
    assert (ref.abs() > 0.999).float().mean() < 0.1, "reference output saturates the final tanh: the comparison would be vacuous"
    assert got.shape == ref.shape and got.shape[0] == nb and got.shape[1] >= t * 160 and got.dtype == torch.float32
    assert rel(got, ref) < 3e-2            # bf16 storage between ~50 layers; a wrong tap / phase / slope gives O(1)
    one = mine(mel[0])                     # un-batched surface, like SpeechT5HifiGan.forward
    assert one.shape == ref[0].shape and rel(one, ref[0]) < 3e-2
