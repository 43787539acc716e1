"""CPU restatement of the reference log-mel front end.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Follows ``/root/reference/model.py:68-97`` (``ComputeMelSpectrogram``).  Its
arithmetic is delegated to torchaudio (reference pins 2.8.0, ``requirements.txt:1-2``;
this image has 2.11.0 - same code path): ``transforms.MelSpectrogram`` ->
``functional.spectrogram`` (reflect-pad n_fft/2, frame, periodic Hann, rFFT,
|.|^2) -> ``MelScale`` (``spec^T @ melscale_fbanks(htk, norm=None)``).

Two forms:
* ``logmel_direct``  - explicit frames, NumPy rFFT, any float dtype (float64 = "truth").
  It only evaluates the frames the reference keeps, ``t = wpi .. T-wpi-2``, whose support
  ``[t*hop - n_fft/2, t*hop + n_fft/2)`` lies inside the signal, so the reflect padding is
  never read (checked against the reference in tests/test_oracle_vs_reference.py).
* ``logmel_torchaudio`` - the same library calls the reference makes; used as the CPU
  baseline arm because that is what the reference costs on a CPU.
"""
from __future__ import annotations

import math

import numpy as np


def hann_periodic(n: int, dtype=np.float64) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)).astype(dtype)


def htk_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """torchaudio.functional.melscale_fbanks(mel_scale='htk', norm=None) in float64."""
    freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_lo = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_hi = 2595.0 * math.log10(1.0 + f_max / 700.0)
    pts = 700.0 * (10.0 ** (np.linspace(m_lo, m_hi, n_mels + 2) / 2595.0) - 1.0)
    width = np.diff(pts)
    slope = pts[None, :] - freqs[:, None]
    rise = -slope[:, :-2] / width[:-1]
    fall = slope[:, 2:] / width[1:]
    return np.maximum(0.0, np.minimum(rise, fall))


def frame_geometry(n_samples: int, sample_rate: int, win_length: int, time_res: float):
    hop = int(time_res * sample_rate)
    wpi = int((win_length / 2) // hop + 1)                # model.py:79
    t_total = 1 + n_samples // hop                        # center=True
    first, last = wpi, t_total - wpi - 2                  # slice [wpi : -(wpi+1)]  model.py:95-97
    return hop, wpi, first, max(0, last - first + 1)


def logmel_direct(wave: np.ndarray, sample_rate: int, win_length: int, time_res: float, n_mels: int,
                  dtype=np.float64, fb: np.ndarray | None = None, window: np.ndarray | None = None) -> np.ndarray:
    wave = np.atleast_2d(np.asarray(wave)).astype(dtype)
    b, n = wave.shape
    hop, _wpi, first, t_out = frame_geometry(n, sample_rate, win_length, time_res)
    win = hann_periodic(win_length, dtype) if window is None else np.asarray(window, dtype)
    fbm = (htk_fbanks(win_length // 2 + 1, 20.0, float(sample_rate // 2), n_mels, sample_rate)
           if fb is None else np.asarray(fb)).astype(dtype)
    out = np.zeros((b, t_out, n_mels), dtype)
    half = win_length // 2
    for j in range(t_out):
        lo = (first + j) * hop - half
        spec = np.fft.rfft(wave[:, lo: lo + win_length] * win, axis=-1)
        power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2)
        out[:, j, :] = power @ fbm
    out = np.log(out + dtype(1e-10))                                   # model.py:91
    out = np.clip(out, dtype(-23), dtype(12))                          # model.py:92
    return ((out + dtype(23)) / dtype(35)).astype(dtype)               # model.py:93


def logmel_torchaudio(wave, sample_rate: int, win_length: int, time_res: float, n_mels: int):
    """model.py:71-97 call for call (torch CPU); returns a float32 torch tensor."""
    import torch
    import torchaudio.transforms as T
    key = (sample_rate, win_length, time_res, n_mels)
    spec = _CACHE.get(key)
    if spec is None:
        spec = _CACHE[key] = T.MelSpectrogram(sample_rate=sample_rate, n_fft=win_length,
                                              hop_length=int(time_res * sample_rate), n_mels=n_mels,
                                              f_min=20.0, power=2)
    wpi = int((win_length / 2) // int(time_res * sample_rate) + 1)
    x = torch.as_tensor(wave).float()
    mel = spec(x)
    y = torch.clamp(torch.log(mel + 1e-10), -23, 12)
    y = (y + 23) / (12 + 23)
    return y.permute(0, 2, 1)[:, wpi: -(wpi + 1), :]


_CACHE: dict = {}
