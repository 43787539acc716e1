#!/usr/bin/env python3
"""One measurement per BASELINE.json configuration besides the headline (bench.py = configs[1], torchrun = configs[2]):

  configs[0]  default synthetiser + log-mel at 16 kHz (configs/config_default.yaml shapes): one batch of 64
  configs[3]  dense-polyphony stress: ~100 events per segment, every one-shot >= 1 s, >= 20 slices per tile
  configs[4]  long-form inference front (tools/bench_longform.py)
  + the long-form *render*: one 600 s SynthDrum call (inference.py:146)

Each on one B200 with the plan resident (CUDA events), next to the CPU oracle port on the host cores over a bounded
sample.  Prints one JSON object; `python tools/bench_configs.py > profiles/rNN_configs.json`.
"""
from __future__ import annotations

import json
import os
import random
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402


def timed(torch, dev, fn, reps):
    fn(); fn(); torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / reps


def cpu_port(segs, cfg, bank, sr, n):
    """Oracle port on one core over the first n segments (render + torchaudio log-mel): audio-s/s per core."""
    from oracle import mel_oracle, synth_oracle
    nested = bank.to_nested()
    rng = random.Random(1)
    t0 = time.perf_counter()
    wavs = [synth_oracle.render(s, cfg, nested, rng=rng) for s in segs[:n]]
    mel_oracle.logmel_torchaudio(synth_oracle.collate(wavs), sr, 2048, 0.01, 128)
    dt = time.perf_counter() - t0
    return {"value": sum(len(w) for w in wavs) / sr / dt, "unit": "audio-s/s", "cores": 1, "kind": "port",
            "sample": f"{n} segments, {dt:.2f} s wall, torch threads = 1"}


def main():
    import torch
    torch.set_num_threads(1)
    from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum
    from adt_str_b200.config import CONFIG_DEFAULT, SETTING_1, config_default, setting_1
    from adt_str_b200.synthetic import make_bank, make_dense_segment, make_long_form, make_segments
    from adt_str_b200.synthetiser import PlanBuffers
    import bench_longform
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {}

    def front_end(cfg, bank, sr):
        return FrontEnd(SynthDrum(cfg, bank=bank, device=dev), ComputeMelSpectrogram(sr, 2048, 0.01, 128))

    def run(name, fe, batches, sr, chunk_batches, reps, cfg_dict, cpu_n):
        plan = fe.plan_batches(batches, random.Random(1234), chunk_batches)
        buf = PlanBuffers(dev)
        buf._dplan = buf.upload(buf.pack(plan)); buf._resident = plan
        wav, feat = fe._outputs(plan, 0)
        ms = timed(torch, dev, lambda: fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False), reps)
        audio_s = float(int(plan.wave_lengths.sum())) / sr
        flat = [s for b in batches for s in b]
        out[name] = {"segments": plan.n_seg, "events": plan.n_events, "audio_s": audio_s, "ms": ms,
                     "value": audio_s / (ms * 1e-3), "unit": "audio-s/s",
                     "bytes_alg": int(4 * int(plan.wave_lengths.sum()) + 4 * 128 * int(plan.mel_total_rows)
                                      + plan.bank_bytes(fe.synth.bank) + 32 * plan.n_events),
                     "slices_per_tile_mean": float(np.diff(plan.tile_ptr).mean()),
                     "cpu_baseline": cpu_port(flat, cfg_dict, fe.synth.bank, sr, cpu_n)}
        out[name]["path_gbs"] = out[name]["bytes_alg"] / (ms * 1e-3) / 1e9

    # configs[0]: 16 kHz default shapes, ONE batch of 64 (latency) and 256 batches (throughput)
    bank16 = make_bank(10_000, 16000, seed=0)
    fe16 = front_end(config_default(), bank16, 16000)
    segs = make_segments(256 * 64, seed=3)
    run("config0_default_16k_one_batch", fe16, [segs[:64]], 16000, 1, 50, dict(CONFIG_DEFAULT), 64)
    run("config0_default_16k_256_batches", fe16, [segs[b * 64:(b + 1) * 64] for b in range(256)], 16000, 4, 5,
        dict(CONFIG_DEFAULT), 128)
    del fe16, bank16

    # configs[3]: dense polyphony, long one-shots
    bank_long = make_bank(2_000, 24000, min_len=24000, max_len=48000, seed=31)
    fe_d = front_end(setting_1(), bank_long, 24000)
    dense = make_dense_segment()
    run("config3_dense_polyphony_64_batches", fe_d, [[dense] * 64 for _ in range(64)], 24000, 2, 5, dict(SETTING_1), 16)
    del fe_d, bank_long

    # configs[4]: long form - the render of one 600 s call, then the inference front
    bank = make_bank(10_000, 24000, seed=0)
    synth = SynthDrum(setting_1(), bank=bank, device=dev)
    notes = make_long_form(600.0)
    plan = synth.plan([notes], random.Random(5))
    buf = PlanBuffers(dev)
    wav = torch.empty((plan.n_seg, plan.ld_wav), dtype=torch.float32, device=dev)
    ms = timed(torch, dev, lambda: synth.render_plan(plan, out=wav), 5)
    out["config4_long_form_render_600s"] = {"events": plan.n_events, "samples": int(plan.wave_lengths[0]), "ms": ms,
                                            "value": float(plan.wave_lengths[0]) / 24000 / (ms * 1e-3), "unit": "audio-s/s"}
    out["config4_long_form_front"] = bench_longform.measure(dev, 600.0, cpu_seconds=600.0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
