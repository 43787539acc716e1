#!/usr/bin/env python3
"""A render with the FX chain on every segment (reverb + compressor + limiter): the launch ncu captures the FX
kernels from (tools/gpu_round.sh), and a quick timing of the chain."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from adt_str_b200 import SynthDrum  # noqa: E402
from adt_str_b200.config import setting_1  # noqa: E402
from adt_str_b200.synthetic import make_bank, make_segments  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    bank = make_bank(2000, 24000, seed=0)
    cfg = setting_1(use_fx_prob=1.0, use_reverb_prob=1.0, use_compression_prob=1.0, use_limiter_prob=1.0)
    synth = SynthDrum(cfg, bank=bank, device=dev)
    segs = make_segments(512, seed=1, empty_fraction=0.0)
    torch.manual_seed(0)
    plan = synth.plan(segs, random.Random(0))
    for _ in range(3):
        wav = synth.render_plan(plan)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        synth.render_plan(plan)
    b.record()
    torch.cuda.synchronize()
    print(f"512 segments, FX on all: {a.elapsed_time(b) / 3:.3f} ms per render, finite {bool(torch.isfinite(wav).all())}")


if __name__ == "__main__":
    main()
