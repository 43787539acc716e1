#!/usr/bin/env python3
"""BASELINE.json configs[4]: the long-form inference front - a 10-minute stereo 44.1 kHz signal resampled to the
model's rate, channel mean, chunks of 2.56 s, log-mel of every chunk (inference.py:75-98 + model.py:81-97).

    python tools/bench_longform.py [--seconds 600] [--cpu-seconds 60]

Prints one JSON object: audio-seconds per second with the signal resident in HBM (CUDA events), end to end from
pinned host memory (H2D of the PCM, D2H of the log-mel), the resample kernel's algorithmic bytes per second, and the
reference's CPU path (torchaudio Resample -> mean -> chunks -> ComputeMelSpectrogram arithmetic) on the host cores
over a bounded sample.  ``measure()`` is also called by bench.py for its ``long_form`` key.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

ORIG_SR, SR, INPUT_SEC = 44100, 24000, 2.56


def synth_signal(seconds: float, device, seed: int = 0):
    """Stereo decaying-noise bursts at 2 Hz plus a low tone, |x| < 1 (same recipe as the golden fixtures)."""
    import torch
    n = int(seconds * ORIG_SR)
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(n, device=device, dtype=torch.float32) / ORIG_SR
    env = torch.exp(-8.0 * ((t * 2.0) % 1.0))
    x = 0.4 * torch.randn(2, n, generator=g, device=device) * env + 0.3 * torch.sin(2 * torch.pi * 110.0 * t)
    return x.contiguous()


def cpu_reference(seconds: float):
    """The reference's CPU path on a bounded sample (oracle = the same library calls), all host threads."""
    import torch
    from oracle import audio_oracle, mel_oracle
    x = synth_signal(seconds, "cpu", seed=1).numpy()
    t0 = time.perf_counter()
    chunks = audio_oracle.long_form_chunks(x, ORIG_SR, SR, INPUT_SEC)
    mel = mel_oracle.logmel_torchaudio(chunks, SR, 2048, 0.01, 128)
    dt = time.perf_counter() - t0
    return {"value": seconds / dt, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{seconds:.0f} s of the same signal class, torch intra-op threads = {torch.get_num_threads()}, "
                      f"{dt:.2f} s wall, log-mel {tuple(mel.shape)}"}


def measure(device, seconds: float = 600.0, reps: int = 5, cpu_seconds: float = 60.0):
    import torch
    from adt_str_b200 import ComputeMelSpectrogram, LongFormFrontEnd
    fe = LongFormFrontEnd(SR, INPUT_SEC, ComputeMelSpectrogram(SR, 2048, 0.01, 128))
    x = synth_signal(seconds, device)
    rs = fe.resampler(ORIG_SR)
    n_out = rs.output_length(x.shape[1])
    y = torch.empty((2, n_out), dtype=torch.float32, device=device)

    def timed(fn, reps=reps):
        fn(); fn(); torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize(device)
        return a.elapsed_time(b) / reps

    whole_ms = timed(lambda: fe(x, ORIG_SR))
    resample_ms = timed(lambda: rs.resample_into(x, y))
    chunks, mel = fe(x, ORIG_SR)
    logmel_ms = timed(lambda: fe.mel(chunks))
    # end to end: pinned host PCM in, pinned host log-mel out
    host_in = torch.empty(x.shape, dtype=torch.float32).pin_memory(); host_in.copy_(x.cpu())
    host_out = torch.empty(mel.shape, dtype=torch.float32).pin_memory()

    def e2e():
        d = host_in.to(device, non_blocking=True)
        _, m = fe(d, ORIG_SR)
        host_out.copy_(m, non_blocking=True)
        torch.cuda.synchronize(device)

    e2e(); e2e()
    t0 = time.perf_counter()
    for _ in range(reps):
        e2e()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / reps
    res_bytes = 4 * (x.numel() + y.numel())
    out = {
        "workload": f"{seconds:.0f} s stereo @ {ORIG_SR} Hz -> {SR} Hz mono, {chunks.shape[0]} chunks of {INPUT_SEC} s, "
                    f"log-mel {tuple(mel.shape)}",
        "value": seconds / (whole_ms * 1e-3), "unit": "audio-s/s", "ms": whole_ms,
        "resample_ms": resample_ms, "logmel_ms": logmel_ms,
        "e2e": {"value": seconds / (e2e_ms * 1e-3), "unit": "audio-s/s", "ms": e2e_ms,
                "h2d_bytes": 4 * x.numel(), "d2h_bytes": 4 * mel.numel()},
        "resample_roofline": {"bound": "fp32 / L1 (171 taps per output sample)", "bytes_alg": res_bytes,
                              "achieved_gbs": res_bytes / (resample_ms * 1e-3) / 1e9,
                              "gflops": 2.0 * 171 * y.numel() / (resample_ms * 1e-3) / 1e9},
    }
    if cpu_seconds > 0:
        out["cpu_baseline"] = cpu_reference(cpu_seconds)
    return out


def main():
    import torch
    p = argparse.ArgumentParser()
    p.add_argument("--seconds", type=float, default=600.0)
    p.add_argument("--cpu-seconds", type=float, default=60.0)
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    print(json.dumps(measure(dev, args.seconds, cpu_seconds=args.cpu_seconds)))


if __name__ == "__main__":
    main()
