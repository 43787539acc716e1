"""Fused front end: notes -> (waveform, log-mel) in one library call per batch.

This is the path ``bench.py`` measures: ``SynthDrum`` (render) feeding
``ComputeMelSpectrogram`` (log-mel) the way reference ``train.py:52`` +
``model.py:248`` chain them, but without the waveform ever leaving the GPU.
``run_plan`` takes a plan already resident on the device; ``run_plan_host`` goes
through ``adtfe_frontend_host`` with pinned host buffers on both sides (plan blob
in, log-mel out), i.e. the end-to-end number.
"""
from __future__ import annotations

import ctypes as C
import random as _random
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .mel import ComputeMelSpectrogram
from .planner import RenderPlan
from .synthetiser import PlanBuffers, SynthDrum


class FrontEnd:
    def __init__(self, synth: SynthDrum, mel: ComputeMelSpectrogram):
        self.synth, self.mel = synth, mel

    def _outputs(self, plan: RenderPlan, n_samples: int):
        dev = self.synth.device
        n_mels = self.mel.compute_spec.n_mels
        wav = torch.empty((plan.n_seg, plan.ld_wav), dtype=torch.float32, device=dev)
        if plan.mel_rows is not None:  # several batches: flat (rows, n_mels), split per batch by plan.split
            feat = torch.empty((plan.mel_total_rows, n_mels), dtype=torch.float32, device=dev)
        else:
            feat = torch.empty((plan.n_seg, self.mel.n_frames(n_samples), n_mels), dtype=torch.float32, device=dev)
        return wav, feat

    def run_plan(self, plan: RenderPlan, buffers: Optional[PlanBuffers] = None, wav: Optional[torch.Tensor] = None,
                 feat: Optional[torch.Tensor] = None, upload: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """Render + log-mel for one planned batch on the current stream.  Returns
        ``(wav (B, Lmax), mel (B, T, n_mels))`` on the device."""
        dev = self.synth.device
        n_samples = int(plan.wave_lengths.max()) if plan.n_seg else 0
        if plan.fx is not None and len(plan.fx) and self.synth.fx_backend == "pedalboard":
            # the FX chain runs on the host (the reference's own library): render, FX, then the log-mel per batch
            full = self.synth.render_plan(plan)
            if plan.mel_rows is None:
                return full[:, :n_samples], self.mel(full[:, :n_samples])
            parts = [self.mel(full[int(plan.batch_ptr[b]):int(plan.batch_ptr[b + 1]), :int(plan.batch_samples[b])])
                     for b in range(len(plan.batch_frames))]
            return full[:, :n_samples], torch.cat([p.reshape(-1, p.shape[-1]) for p in parts])
        with torch.cuda.device(dev):
            bank = self.synth.device_bank()
            native = self.mel._handle(dev)
            buf = buffers or self.synth.buffers()
            if upload or not getattr(buf, "_resident", None) is plan:
                shape = buf.pack(plan)
                buf._dplan = buf.upload(shape)
                buf._resident = plan
            if wav is None or feat is None:
                wav, feat = self._outputs(plan, n_samples)
            stream = torch.cuda.current_stream(dev).cuda_stream
            if plan.n_seg:
                _lib.check(bank.lib.adtfe_render_logmel(bank.handle, native.handle, C.byref(buf._dplan), n_samples,
                                                        wav.data_ptr(), feat.data_ptr(), buf.workspace.data_ptr(),
                                                        buf.workspace.numel(), stream), "adtfe_render_logmel")
        return wav[:, :n_samples], feat

    def __call__(self, batch_notes: Sequence, rng=_random) -> Tuple[torch.Tensor, torch.Tensor]:
        """``[notes, ...] -> (wav (B, Lmax) cuda, logmel (B, T, n_mels) cuda)``."""
        return self.run_plan(self.synth.plan(batch_notes, rng))

    def plan_batches(self, batches: Sequence[Sequence], rng=_random, chunk_batches: int = 1) -> RenderPlan:
        return self.synth.plan_batches(batches, self.mel.n_frames, rng, chunk_batches)

    def run_batches(self, batches: Sequence[Sequence], rng=_random):
        """Several batches through one plan: one H2D copy and one launch per kernel for all of them
        (what a prefetching loader hands over).  Returns ``[(wav_b, logmel_b), ...]``, each pair
        what ``__call__`` would return for that batch."""
        plan = self.plan_batches(batches, rng)
        wav, feat = self.run_plan(plan)
        return plan.split(wav, feat)

    def run_plan_host(self, plan: RenderPlan, mel_out_host: torch.Tensor, wav_out_host: Optional[torch.Tensor] = None,
                      buffers: Optional[PlanBuffers] = None, wav: Optional[torch.Tensor] = None,
                      feat: Optional[torch.Tensor] = None, packed: bool = False,
                      copy_stream: Optional[torch.cuda.Stream] = None) -> None:
        """End-to-end entry with HOST buffers: packs the plan into pinned memory (unless ``packed``:
        ``buffers.pack(plan)`` was already called, e.g. by a planning thread), then one
        ``adtfe_frontend_host`` call does H2D(plan) -> render -> log-mel -> D2H(log-mel[, wav]).
        Asynchronous; synchronise the current stream (``copy_stream`` when given: the D2H copies then
        run there, overlapping the next call's kernels) before reading ``mel_out_host``."""
        dev = self.synth.device
        n_samples = int(plan.wave_lengths.max()) if plan.n_seg else 0
        if not mel_out_host.is_pinned() or (wav_out_host is not None and not wav_out_host.is_pinned()):
            raise ValueError("host output buffers must be pinned")
        with torch.cuda.device(dev):
            bank = self.synth.device_bank()
            native = self.mel._handle(dev)
            buf = buffers or self.synth.buffers()
            shape = buf.shape if packed else buf.pack(plan)
            if wav is None or feat is None:
                wav, feat = self._outputs(plan, n_samples)
            need = plan.mel_total_rows * feat.shape[-1] if plan.mel_rows is not None else feat.numel()
            if mel_out_host.numel() < need or (wav_out_host is not None and wav_out_host.numel() < wav.numel()):
                raise ValueError("host output buffer too small")
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(bank.lib.adtfe_frontend_host(
                bank.handle, native.handle, C.byref(shape), n_samples, buf.host.data_ptr(), buf.nbytes,
                buf.dev.data_ptr(), wav.data_ptr(), feat.data_ptr(), buf.workspace.data_ptr(), buf.workspace.numel(),
                mel_out_host.data_ptr(), wav_out_host.data_ptr() if wav_out_host is not None else None, stream,
                copy_stream.cuda_stream if copy_stream is not None else None),
                "adtfe_frontend_host")
            buf._uploaded = torch.cuda.Event()   # the call's H2D copy reads the pinned blob: see PlanBuffers.fence
            buf._uploaded.record()
