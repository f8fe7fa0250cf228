"""Log-mel front-end of the reference and the log-mel L1 distance between two waveforms.

Restates `AudioDataset.mel_spectrogram_train` -- /root/reference/script/data/datasets.py:301-354 -- with the
parameters that file hard-codes (:74-80): 16 kHz, filter_length = win_length = 1024, hop 160, 64 mel bins,
fmin 0, fmax 8000, `librosa.filters.mel` basis (Slaney mel scale, Slaney area normalisation), reflect padding of
(filter_length - hop) / 2 samples per side, magnitude STFT (center=False), `log(clamp(., 1e-5))` (:19-27).
This is the spectrogram the AudioLDM VAE is trained on and the HiFi-GAN vocoder inverts, so it is the natural
space in which to STATE how far two waveforms are apart (BASELINE.json north_star: "final-waveform log-mel L1
stated"): `logmel_l1(a, b)` = mean |logmel(a) - logmel(b)| in nats.

Host-side torch code (torch.stft); it is a metric / data front-end, not part of the timed denoising path.
librosa is not installed here: the filter bank below restates its published construction and is checked against
`torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney")` in tests/test_mel.py.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

SAMPLING_RATE = 16000
FILTER_LENGTH = 1024
HOP_LENGTH = 160
WIN_LENGTH = 1024
N_MEL = 64
MEL_FMIN = 0.0
MEL_FMAX = 8000.0
CLIP_VAL = 1e-5

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def hz_to_mel(f: torch.Tensor) -> torch.Tensor:
    """Slaney (Auditory Toolbox) mel scale: linear below 1 kHz, logarithmic above (librosa htk=False)."""
    f = torch.as_tensor(f, dtype=torch.float64)
    lin = f / _F_SP
    log = _MIN_LOG_MEL + torch.log(torch.clamp(f, min=1e-10) / _MIN_LOG_HZ) / _LOGSTEP
    return torch.where(f >= _MIN_LOG_HZ, log, lin)


def mel_to_hz(m: torch.Tensor) -> torch.Tensor:
    m = torch.as_tensor(m, dtype=torch.float64)
    lin = m * _F_SP
    log = _MIN_LOG_HZ * torch.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return torch.where(m >= _MIN_LOG_MEL, log, lin)


def mel_filterbank(sr: int = SAMPLING_RATE, n_fft: int = FILTER_LENGTH, n_mels: int = N_MEL, fmin: float = MEL_FMIN,
                   fmax: float = MEL_FMAX) -> torch.Tensor:
    """`librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` -> fp32 [n_mels, n_fft // 2 + 1]: triangular filters with
    corners equally spaced on the Slaney mel scale, each scaled by 2 / (its bandwidth in Hz)."""
    fft_f = torch.linspace(0.0, sr / 2.0, n_fft // 2 + 1, dtype=torch.float64)
    mel_f = mel_to_hz(torch.linspace(float(hz_to_mel(fmin)), float(hz_to_mel(fmax)), n_mels + 2, dtype=torch.float64))
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp(torch.minimum(lower, upper), min=0.0)
    w = w * (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]
    return w.float()


_CACHE: Dict[Tuple[str, int], Tuple[torch.Tensor, torch.Tensor]] = {}


def log_mel_spectrogram(wave: torch.Tensor) -> torch.Tensor:
    """wave fp32 [B, samples] (or [samples]) in [-1, 1] at 16 kHz -> log-mel [B, 64, frames], frames = samples // 160
    (datasets.py:301-354)."""
    y = torch.as_tensor(wave, dtype=torch.float32)
    if y.dim() == 1:
        y = y[None]
    key = (str(y.device), N_MEL)
    if key not in _CACHE:
        _CACHE[key] = (mel_filterbank().to(y.device), torch.hann_window(WIN_LENGTH, device=y.device))
    basis, window = _CACHE[key]
    pad = (FILTER_LENGTH - HOP_LENGTH) // 2
    y = torch.nn.functional.pad(y[:, None], (pad, pad), mode="reflect")[:, 0]
    spec = torch.stft(y, FILTER_LENGTH, hop_length=HOP_LENGTH, win_length=WIN_LENGTH, window=window, center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True).abs()
    return torch.log(torch.clamp(basis @ spec, min=CLIP_VAL))


def logmel_l1(a: torch.Tensor, b: torch.Tensor) -> float:
    """Mean absolute difference of the two waveforms' log-mel spectrograms (nats per time-frequency bin)."""
    a, b = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)
    if a.shape != b.shape:
        raise ValueError(f"logmel_l1: waveform shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    return (log_mel_spectrogram(a) - log_mel_spectrogram(b)).abs().mean().item()
