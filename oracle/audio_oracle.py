"""CPU restatement of the reference's eval / inference audio front.  TEST INFRASTRUCTURE ONLY - imported by
``tests/`` and the benchmark's CPU leg, never by ``adt_str_b200``.

Follows ``utils/audio_utils.py:17-23`` (``resample`` = ``torchaudio.transforms.Resample``, ``normalize``),
``utils/audio_utils.py:12`` / ``inference.py:86-87`` (channel mean) and ``inference.py:35-48,75-98`` (chunking, order of
the steps).  The resampling arithmetic lives in torchaudio (reference pins 2.8.0, ``requirements.txt:1-2``; this
image has 2.11.0 - same ``functional._get_sinc_resample_kernel`` / ``_apply_sinc_resample_kernel``): ``sinc_kernel``
restates its published algorithm in NumPy float64, ``resample_direct`` applies it as the explicit sum
``y[j*n + p] = sum_k xpad[j*o + k] * kernel[p][k]`` in any dtype.  Pinned against the live reference functions in
tests/test_oracle_vs_reference.py and against tests/golden/audio_front.npz everywhere.
"""
from __future__ import annotations

import math

import numpy as np

from .synth_oracle import chunk_audio  # noqa: F401  (inference.py:35-48)


def sinc_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """float64 ``(kernel (new/gcd, 2*width + orig/gcd), width)`` - Hann-windowed sinc, torchaudio's defaults."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    # torchaudio divides an int64 arange by an int: the phase offsets are float32 numbers
    phase = (np.arange(0, -n, -1, dtype=np.int64).astype(np.float32) / np.float32(n)).astype(np.float64)[:, None]
    t = (phase + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return k * window * (base / o), width


def resample_direct(x: np.ndarray, orig_freq: int, new_freq: int, dtype=np.float64, kernel=None) -> np.ndarray:
    """Explicit polyphase sum over the last axis.  ``kernel``: the float32 bank to use (default: ``sinc_kernel``
    rounded to float32 as torchaudio caches it)."""
    if orig_freq == new_freq:
        return np.asarray(x)
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    k64, width = sinc_kernel(orig_freq, new_freq)
    ker = (k64.astype(np.float32) if kernel is None else np.asarray(kernel).reshape(n, -1)).astype(dtype)
    x = np.asarray(x)
    lead = x.shape[:-1]
    rows = x.reshape(-1, x.shape[-1]).astype(dtype)
    length = rows.shape[1]
    taps = ker.shape[1]
    xpad = np.pad(rows, ((0, 0), (width, width + o)))
    n_groups = (xpad.shape[1] - taps) // o + 1
    win = np.lib.stride_tricks.sliding_window_view(xpad, taps, axis=1)[:, ::o][:, :n_groups]   # (rows, J, taps)
    y = np.einsum("rjk,pk->rjp", win, ker).reshape(rows.shape[0], -1)
    target = -(-n * length // o)
    return y[:, :target].reshape(lead + (target,))


def resample_torchaudio(x, orig_freq: int, new_freq: int):
    """The same library call the reference makes (utils/audio_utils.py:17-19), on CPU."""
    import torch
    import torchaudio.transforms as T
    return T.Resample(orig_freq=orig_freq, new_freq=new_freq)(torch.as_tensor(np.asarray(x, np.float32))).numpy()


def normalize(x: np.ndarray) -> np.ndarray:
    """utils/audio_utils.py:22-23 in float32."""
    x = np.asarray(x, np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return x / np.abs(x).max()


def downmix(x: np.ndarray) -> np.ndarray:
    """``wav.mean(0)`` in float32: channels added in order, divided by the count."""
    x = np.asarray(x, np.float32)
    s = x[0].copy()
    for c in range(1, x.shape[0]):
        s = s + x[c]
    return s / np.float32(x.shape[0])


def long_form_chunks(waveform: np.ndarray, sample_rate: int, target_sr: int, input_sec: float, dtype=np.float32):
    """inference.py:75-98 up to the chunk list: resample all channels, channel mean, chunks of round(input_sec*sr)."""
    w = np.asarray(waveform, np.float32)
    if sample_rate != target_sr:
        w = resample_torchaudio(w, sample_rate, target_sr) if dtype == np.float32 else \
            resample_direct(w, sample_rate, target_sr, dtype)
    if w.shape[0] > 1:
        w = downmix(w)[None] if dtype == np.float32 else w.mean(0, keepdims=True)
    return chunk_audio(np.asarray(w[0], np.float32), int(round(input_sec * target_sr)))
