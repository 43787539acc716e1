"""Audio front of eval / inference on the GPU, with the reference's call signatures.

Mirrors ``utils/audio_utils.py`` of pier-maker92/ADT_STR:

* ``resample(wav_seg, orig_sr, target_sr)``   (:17-19)  = ``torchaudio.transforms.Resample(orig, new)(wav)``
* ``normalize(wav_seg)``                      (:22-23)  = ``wav / wav.abs().max()``
* ``Resample``                                the transform itself (``inference.py:82-84``): same constructor
  arguments, same ``kernel`` buffer and ``width`` attribute; ``forward`` runs ``adtfe_resample``
* ``downmix(waveform)``                       the channel mean of ``utils/audio_utils.py:12`` / ``inference.py:86-87``

Results are returned on the input's device (CPU in -> CPU out, like the reference); the arithmetic always runs
in the CUDA library - there is no CPU path.  ``sinc_resample_kernel`` restates
``torchaudio.functional.functional._get_sinc_resample_kernel`` operation for operation (float64, cast to float32)
and is bit-identical to it (tests/test_audio_front.py).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99,
                         resampling_method: str = "sinc_interp_hann", beta: Optional[float] = None
                         ) -> Tuple[torch.Tensor, int]:
    """``(kernel (new/gcd, 1, 2*width + orig/gcd) float32, width)`` - what ``T.Resample`` caches.

    This restates ``torchaudio.functional.functional._get_sinc_resample_kernel`` operation for operation (the buffer has
    to be bit-identical to the one in existing state dicts).  torchaudio is BSD 2-Clause licensed, Copyright (c) 2017
    Facebook Inc. (Soumith Chintala); redistribution of this derived function keeps that notice."""
    if not (int(orig_freq) == orig_freq and int(new_freq) == new_freq):
        raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
    if resampling_method not in ("sinc_interp_hann", "sinc_interp_kaiser"):
        raise ValueError("Invalid resampling method: {}".format(resampling_method))
    if lowpass_filter_width <= 0:
        raise ValueError("Low pass filter width should be positive.")
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base_freq = min(orig, new)
    base_freq *= rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=None)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    if resampling_method == "sinc_interp_hann":
        window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    else:
        if beta is None:
            beta = 14.769656459379492
        beta_tensor = torch.tensor(float(beta))
        window = torch.i0(beta_tensor * torch.sqrt(1 - (t / lowpass_filter_width) ** 2)) / torch.i0(beta_tensor)
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(dtype=torch.float32), width


def _cuda_device(t: torch.Tensor) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("adt_str_b200.audio_utils needs a CUDA device (sm_100a); there is no CPU path")
    if t.device.type == "cuda":
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


class _NativeResampler:
    def __init__(self, orig_freq: int, new_freq: int, width: int, kernel: torch.Tensor, device: torch.device):
        lib = _lib.load()
        k = kernel.detach().to("cpu", torch.float32).reshape(kernel.shape[0], -1).contiguous()
        gcd = math.gcd(int(orig_freq), int(new_freq))
        if tuple(k.shape) != (int(new_freq) // gcd, 2 * width + int(orig_freq) // gcd):
            raise ValueError("kernel buffer does not match (orig_freq, new_freq, width)")
        h = C.c_void_p()
        _lib.check(lib.adtfe_resampler_create(int(orig_freq), int(new_freq), int(width), k.data_ptr(), device.index,
                                              C.byref(h)), "adtfe_resampler_create")
        self.handle, self.lib = h, lib

    def __del__(self):
        try:
            if self.handle:
                self.lib.adtfe_resampler_destroy(self.handle)
        except Exception:
            pass


class Resample(nn.Module):
    """Drop-in for ``torchaudio.transforms.Resample`` (float32): ``forward(waveform (..., time)) -> (..., time')``."""

    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000, resampling_method: str = "sinc_interp_hann",
                 lowpass_filter_width: int = 6, rolloff: float = 0.99, beta: Optional[float] = None):
        super().__init__()
        self.orig_freq, self.new_freq = orig_freq, new_freq
        self.gcd = math.gcd(int(orig_freq), int(new_freq))
        self.resampling_method, self.lowpass_filter_width, self.rolloff, self.beta = \
            resampling_method, lowpass_filter_width, rolloff, beta
        self._native = {}
        if self.orig_freq != self.new_freq:
            kernel, self.width = sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width, rolloff,
                                                      resampling_method, beta)
            self.register_buffer("kernel", kernel)

    def _handle(self, device: torch.device) -> _NativeResampler:
        k = self.kernel
        key = (k.data_ptr(), k._version)
        hit = self._native.get(device.index)
        if hit is None or hit[0] != key:
            hit = (key, _NativeResampler(self.orig_freq, self.new_freq, self.width, k, device))
            self._native[device.index] = hit
        return hit[1]

    def output_length(self, n_in: int) -> int:
        if self.orig_freq == self.new_freq:
            return n_in
        return -(-(int(self.new_freq) // self.gcd) * n_in // (int(self.orig_freq) // self.gcd))

    def resample_into(self, x: torch.Tensor, out: torch.Tensor, absmax_bits: Optional[torch.Tensor] = None) -> None:
        """``x`` (rows, n) float32 cuda with unit column stride -> ``out`` (rows, >= output_length(n)) on the current
        stream; ``absmax_bits``: zeroed int32[1] that receives the float bits of ``max|out|``."""
        dev = x.device
        native = self._handle(dev)
        with torch.cuda.device(dev):
            _lib.check(native.lib.adtfe_resample(
                native.handle, x.data_ptr(), x.shape[0], x.stride(0) if x.shape[0] > 1 else max(x.shape[1], 1),
                x.shape[1], out.data_ptr(), out.stride(0) if out.shape[0] > 1 else max(out.shape[1], 1),
                absmax_bits.data_ptr() if absmax_bits is not None else None,
                torch.cuda.current_stream(dev).cuda_stream), "adtfe_resample")

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if self.orig_freq == self.new_freq:
            return waveform
        if not waveform.is_floating_point():
            raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
        src = waveform.device
        dev = _cuda_device(waveform)
        shape = waveform.shape
        x = waveform.to(dev, torch.float32).reshape(math.prod(shape[:-1]), shape[-1])
        if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
            x = x.contiguous()
        out = torch.empty((x.shape[0], self.output_length(shape[-1])), dtype=torch.float32, device=dev)
        if out.numel():
            self.resample_into(x, out)
        out = out.view(shape[:-1] + out.shape[-1:])
        return out if src == dev else out.to(src)


def resample(wav_seg: torch.Tensor, orig_sr: int, target_sr: int) -> torch.Tensor:
    """utils/audio_utils.py:17-19."""
    return Resample(orig_freq=orig_sr, new_freq=target_sr)(wav_seg)


def load_and_resample(wav_file: str, target_sr: Optional[int], loader=None) -> torch.Tensor:
    """utils/audio_utils.py:10-15: decode the file (``torchaudio.load``, on the host - or ``loader(path) -> (waveform
    (channels, samples), sample_rate)``), channel mean, resample to ``target_sr`` (``None``: keep the rate).  Mean and
    resampling run on the GPU; the result comes back where the decoded audio was (CPU), like the reference's."""
    if loader is None:
        import torchaudio
        loader = torchaudio.load
    wav_seg, orig_sr = loader(wav_file)
    wav_seg = downmix(wav_seg)
    if target_sr is None:
        return wav_seg
    return resample(wav_seg, int(orig_sr), int(target_sr))


def normalize(wav_seg: torch.Tensor) -> torch.Tensor:
    """utils/audio_utils.py:22-23 - ``wav_seg / wav_seg.abs().max()`` (a new tensor; all-zero input gives NaN)."""
    src = wav_seg.device
    dev = _cuda_device(wav_seg)
    x = wav_seg.to(dev, torch.float32).contiguous()
    if x.data_ptr() == wav_seg.data_ptr():
        x = x.clone()
    if x.numel() == 0:
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0")
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().adtfe_peak_normalise(x.data_ptr(), x.numel(), scratch.data_ptr(), 0,
                                                    torch.cuda.current_stream(dev).cuda_stream), "adtfe_peak_normalise")
    return x if src == dev else x.to(src)


def downmix(waveform: torch.Tensor, keepdim: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``waveform.mean(0)`` of a (channels, samples) signal (utils/audio_utils.py:12, inference.py:86-87).
    ``out``: a contiguous float32 cuda vector of ``samples`` elements to write into."""
    if waveform.dim() != 2:
        raise ValueError(f"waveform must be (channels, samples), got {tuple(waveform.shape)}")
    src = waveform.device
    dev = _cuda_device(waveform)
    x = waveform.to(dev, torch.float32)
    if x.stride(1) != 1:
        x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape[1], dtype=torch.float32, device=dev)
    elif out.device != dev or out.dtype != torch.float32 or out.shape != (x.shape[1],) or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 vector of `samples` elements on the compute device")
    if x.shape[0] == 0:
        return torch.full_like(out, float("nan")).to(src)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().adtfe_downmix(x.data_ptr(), x.shape[0], x.stride(0) if x.shape[0] > 1 else max(x.shape[1], 1),
                                             x.shape[1], out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                   "adtfe_downmix")
    if keepdim:
        out = out.unsqueeze(0)
    return out if src == dev else out.to(src)
