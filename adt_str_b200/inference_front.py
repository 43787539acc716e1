"""Long-form front of ``inference.py`` on the GPU: raw audio -> chunks -> log-mel.

Mirrors the reference's ``inference.py``:

* ``_chunk_audio(wav, chunk_samples)`` (:35-48) -> ``chunk_audio``: the same list of ``(start, chunk)`` pairs; the
  chunks are views of ONE zero-padded copy of the signal, not ``len(chunks)`` concatenations;
* the front of ``main()`` (:75-98): resample to the model's rate when it differs (:82-84), channel mean (:86-87),
  chunks of ``round(input_sec * sr)`` samples, and per chunk the log-mel ``ComputeMelSpectrogram`` the model applies
  to its input (``model.py:248,290,361``) -> ``LongFormFrontEnd``: one resample launch per signal that writes straight
  into the padded chunk matrix and one log-mel launch over all chunks.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .audio_utils import Resample, _cuda_device, downmix
from .mel import ComputeMelSpectrogram


def chunk_audio(wav: torch.Tensor, chunk_samples: int) -> List[Tuple[int, torch.Tensor]]:
    """inference.py:35-48 - ``wav`` (channels, samples) -> ``[(start, chunk (channels, chunk_samples)), ...]``,
    the last chunk zero-padded.  Works on any device; an empty signal gives an empty list."""
    if wav.dim() != 2:
        raise ValueError(f"wav must be (channels, samples), got {tuple(wav.shape)}")
    if chunk_samples <= 0:
        raise ValueError("range() arg 3 must not be zero" if chunk_samples == 0 else "chunk_samples must be positive")
    n = wav.shape[-1]
    k = -(-n // chunk_samples)
    if k * chunk_samples != n:
        wav = torch.nn.functional.pad(wav, (0, k * chunk_samples - n))
    return [(i * chunk_samples, wav[:, i * chunk_samples:(i + 1) * chunk_samples]) for i in range(k)]


class LongFormFrontEnd:
    """``(waveform (channels, samples), sample_rate) -> (chunks (n, chunk_samples), logmel (n, T, n_mels))`` on the
    GPU, with the arithmetic in the order ``inference.py`` applies it: resample every channel, then the mean."""

    def __init__(self, sample_rate: int, input_sec: float, mel: ComputeMelSpectrogram):
        self.sample_rate, self.input_sec, self.mel = int(sample_rate), float(input_sec), mel
        self.chunk_samples = int(round(self.input_sec * self.sample_rate))     # inference.py:80
        self._resamplers = {}

    def resampler(self, orig_sr: int) -> Resample:
        r = self._resamplers.get(int(orig_sr))
        if r is None:
            r = self._resamplers[int(orig_sr)] = Resample(int(orig_sr), self.sample_rate)
        return r

    def chunks(self, waveform: torch.Tensor, sample_rate: Optional[int] = None) -> torch.Tensor:
        """The (n_chunks, chunk_samples) matrix of mono chunks at the model's rate (device memory)."""
        if waveform.dim() != 2:
            raise ValueError(f"waveform must be (channels, samples), got {tuple(waveform.shape)}")
        dev = _cuda_device(waveform)
        x = waveform.to(dev, torch.float32)
        if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
            x = x.contiguous()
        c, n_in = x.shape
        sr = self.sample_rate if sample_rate is None else int(sample_rate)
        if sr != self.sample_rate:
            rs = self.resampler(sr)
            n = rs.output_length(n_in)
        else:
            rs, n = None, n_in
        k = -(-n // self.chunk_samples)
        padded = k * self.chunk_samples
        if c == 1:   # mono: resample straight into the zero-padded chunk matrix
            out = torch.zeros(padded, dtype=torch.float32, device=dev)
            if rs is not None and n:
                rs.resample_into(x, out[:n].unsqueeze(0))
            else:
                out[:n] = x[0]
        else:
            y = x
            if rs is not None:
                y = torch.empty((c, n), dtype=torch.float32, device=dev)
                if n:
                    rs.resample_into(x, y)
            out = torch.zeros(padded, dtype=torch.float32, device=dev)
            if n:
                downmix(y, out=out[:n])
        return out.view(k, self.chunk_samples)

    def __call__(self, waveform: torch.Tensor, sample_rate: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        chunks = self.chunks(waveform, sample_rate)
        if chunks.shape[0] == 0:
            n_mels = self.mel.compute_spec.n_mels
            return chunks, torch.empty((0, self.mel.n_frames(self.chunk_samples), n_mels), device=chunks.device)
        return chunks, self.mel(chunks)
