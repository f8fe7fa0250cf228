"""CPU tests of the log-mel front-end / metric (audioldm_with_lora_b200/mel.py), the restatement of
/root/reference/script/data/datasets.py:301-354 used to STATE the final-waveform log-mel L1."""
import math

import pytest
import torch

from audioldm_with_lora_b200 import mel


def test_filterbank_matches_torchaudio_slaney():
    """librosa is absent; torchaudio's Slaney-scale, Slaney-normalised bank is the same published construction."""
    ta = pytest.importorskip("torchaudio")
    want = ta.functional.melscale_fbanks(513, 0.0, 8000.0, 64, 16000, norm="slaney", mel_scale="slaney").T
    got = mel.mel_filterbank()
    assert got.shape == (64, 513)
    assert torch.allclose(got, want, atol=1e-6, rtol=1e-4)


def test_filterbank_known_answers():
    fb = mel.mel_filterbank()
    assert (fb >= 0).all() and (fb.sum(1) > 0).all()
    # Slaney normalisation: every triangle has unit area in Hz (bin width 16000 / 1024 Hz), up to the sampling of its
    # corners; the lowest filters are only a few bins wide, so the tolerance is loose there
    area = fb.double().sum(1) * (16000 / 1024)
    assert torch.allclose(area[8:], torch.ones(56, dtype=torch.float64), atol=0.05)
    # mel scale: linear (200/3 Hz per mel) below 1 kHz, 1 kHz == mel 15, 6.4 kHz is 27 log steps above
    assert float(mel.hz_to_mel(1000.0)) == pytest.approx(15.0)
    assert float(mel.hz_to_mel(6400.0)) == pytest.approx(42.0)
    assert float(mel.mel_to_hz(mel.hz_to_mel(3210.0))) == pytest.approx(3210.0)


def test_log_mel_of_a_sine_and_metric_properties():
    sr, n = 16000, 16000
    t = torch.arange(n) / sr
    tone = 0.5 * torch.sin(2 * math.pi * 1000.0 * t)
    lm = mel.log_mel_spectrogram(tone)
    assert lm.shape == (1, 64, n // 160)                       # hop 160, reflect padding: frames = samples // hop
    peak_bin = lm[0].mean(1).argmax().item()
    centre = mel.mel_to_hz(torch.linspace(float(mel.hz_to_mel(0.0)), float(mel.hz_to_mel(8000.0)), 66))[1:-1]
    assert abs(float(centre[peak_bin]) - 1000.0) < 80.0        # energy sits in the filter centred near 1 kHz
    assert float(mel.log_mel_spectrogram(torch.zeros(n)).max()) == pytest.approx(math.log(1e-5))   # clamp floor
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, n, generator=g) * 0.1
    assert mel.logmel_l1(a, a) == 0.0
    assert mel.logmel_l1(a, a * 2) == pytest.approx(math.log(2.0), abs=1e-4)   # gain 2 = +ln 2 in every bin
    small, big = mel.logmel_l1(a, a + 1e-3 * torch.randn(2, n, generator=g)), mel.logmel_l1(a, a.flip(0))
    assert 0 < small < big
    with pytest.raises(ValueError):
        mel.logmel_l1(a, a[:, :100])
