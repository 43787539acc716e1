"""``ComputeMelSpectrogram`` - drop-in for reference ``model.py:68-97``.

Same constructor ``(sample_rate, win_length, time_res, n_mels)``, same
``forward(wave[B, L]) -> float32 [B, T, n_mels]`` on ``wave.device``, same
``window_pad_idxs`` attribute, and the same two state-dict buffers the reference
gets from torchaudio's ``MelSpectrogram`` - ``compute_spec.spectrogram.window``
and ``compute_spec.mel_scale.fb`` - so checkpoints load with ``strict=True``
(``build_model.py:66``).  The arithmetic runs in one fused sm_100a kernel
(``csrc/logmel.cu``) that reads the window and filterbank *from those buffers*.
There is no CPU implementation: without a B200 the call raises.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib


def htk_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """(n_freqs, n_mels) triangular HTK mel filterbank, norm=None - the float32 op sequence of
    ``torchaudio.functional.melscale_fbanks`` so the buffer is bit-identical to the reference's."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    mel_lo = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    mel_hi = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    pts = 700.0 * (10.0 ** (torch.linspace(mel_lo, mel_hi, n_mels + 2) / 2595.0) - 1.0)
    width = pts[1:] - pts[:-1]
    slope = pts.unsqueeze(0) - freqs.unsqueeze(1)
    rising = (-1.0 * slope[:, :-2]) / width[:-1]
    falling = slope[:, 2:] / width[1:]
    return torch.max(torch.zeros(1), torch.min(rising, falling))


class _Window(nn.Module):
    def __init__(self, n_fft: int):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))  # periodic, as torchaudio's default


class _FilterBank(nn.Module):
    def __init__(self, n_freqs: int, n_mels: int, sample_rate: int, f_min: float):
        super().__init__()
        self.register_buffer("fb", htk_filterbank(n_freqs, f_min, float(sample_rate // 2), n_mels, sample_rate))


class _MelSpectrogramState(nn.Module):
    """Holds the buffers under torchaudio's names: ``spectrogram.window``, ``mel_scale.fb``."""

    def __init__(self, sample_rate: int, n_fft: int, hop_length: int, n_mels: int, f_min: float):
        super().__init__()
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        self.spectrogram = _Window(n_fft)
        self.mel_scale = _FilterBank(n_fft // 2 + 1, n_mels, sample_rate, f_min)


class _NativeMel:
    def __init__(self, n_fft, hop, n_mels, window: torch.Tensor, fb: torch.Tensor, device: torch.device):
        lib = _lib.load()
        w = window.detach().to("cpu", torch.float32).contiguous()
        f = fb.detach().to("cpu", torch.float32).contiguous()
        if w.numel() != n_fft or tuple(f.shape) != (n_fft // 2 + 1, n_mels):
            raise ValueError("window / fb buffers do not match (n_fft, n_mels)")
        h = C.c_void_p()
        _lib.check(lib.adtfe_mel_create(n_fft, hop, n_mels, w.data_ptr(), f.data_ptr(), device.index, C.byref(h)),
                   "adtfe_mel_create")
        self.handle, self.lib = h, lib

    def __del__(self):
        try:
            if self.handle:
                self.lib.adtfe_mel_destroy(self.handle)
        except Exception:
            pass


class ComputeMelSpectrogram(nn.Module):
    def __init__(self, sample_rate, win_length, time_res, n_mels):
        super().__init__()
        hop = int(time_res * sample_rate)
        self.compute_spec = _MelSpectrogramState(sample_rate, win_length, hop, n_mels, f_min=20.0)
        self.window_pad_idxs = int((win_length / 2) // hop + 1)
        self._native = {}

    # -- native handle, rebuilt if the buffers are replaced (load_state_dict, .to())
    def _handle(self, device: torch.device):
        st = self.compute_spec
        w, fb = st.spectrogram.window, st.mel_scale.fb
        key = (device.index, w.data_ptr(), w._version, fb.data_ptr(), fb._version)
        hit = self._native.get(device.index)
        if hit is None or hit[0] != key:
            hit = (key, _NativeMel(st.n_fft, st.hop_length, st.n_mels, w, fb, device))
            self._native[device.index] = hit
        return hit[1]

    def n_frames(self, n_samples: int) -> int:
        """Frames kept for an ``n_samples`` long input (model.py:95-97)."""
        return max(0, 1 + n_samples // self.compute_spec.hop_length - 2 * self.window_pad_idxs - 1)

    def forward(self, wave):
        if wave.dim() != 2:
            raise ValueError(f"wave must be (batch, samples), got {tuple(wave.shape)}")
        if not torch.cuda.is_available():
            raise RuntimeError("ComputeMelSpectrogram needs a CUDA device (sm_100a); there is no CPU path")
        src_device = wave.device
        dev = src_device if src_device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        x = wave.to(dev, torch.float32)              # model.py:88 wave.float()
        if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
            x = x.contiguous()
        b, n = x.shape
        out = torch.empty((b, self.n_frames(n), self.compute_spec.n_mels), dtype=torch.float32, device=dev)
        if out.numel():
            native = self._handle(dev)
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                ld = x.stride(0) if b > 1 else max(n, 1)
                _lib.check(native.lib.adtfe_logmel(native.handle, x.data_ptr(), b, ld, n, out.data_ptr(), stream),
                           "adtfe_logmel")
        return out if src_device == dev else out.to(src_device)
