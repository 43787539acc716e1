"""Digest of rendered waveforms and log-mel for fixed seeded plans: run once in this tree and once in a checkout of
another revision (``git worktree add _r1 <rev>``), compare the printed lines - equal digests = bit-identical output.

    python tools/wav_digest.py            # this tree
    (cd _r1 && python ../tools/wav_digest.py)
"""
import hashlib
import os
import random
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum
from adt_str_b200.config import setting_1
from adt_str_b200.synthetic import make_bank, make_segments


def digest(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    dev = torch.device("cuda", 0)
    bank = make_bank(2000, 24000, seed=0)
    fe = FrontEnd(SynthDrum(setting_1(), bank=bank, device=dev), ComputeMelSpectrogram(24000, 2048, 0.01, 128))
    segs = make_segments(64 * 12, seed=3)
    batches = [segs[b * 64:(b + 1) * 64] for b in range(12)]
    plan = fe.plan_batches(batches, random.Random(42), 4)
    wav, feat = fe.run_plan(plan)
    torch.cuda.synchronize()
    print("setting-1 12 batches  wav", digest(wav), "logmel", digest(feat), "nan", int(torch.isnan(wav).sum()))
    # dense polyphony: many notes per tile (more than 32 on some), long one-shots
    rng = np.random.default_rng(5)
    dense = []
    for _ in range(16):
        on = np.sort(rng.uniform(0, 2.45, 96)).astype(np.float32)
        pitch = rng.choice([42, 48, 36, 38], 96).astype(np.float32)
        vel = rng.integers(10, 127, 96).astype(np.float32)
        dense.append(np.stack([on, on + np.float32(0.1), pitch, vel], 1))
    w2, f2 = fe(dense, random.Random(7))
    torch.cuda.synchronize()
    print("dense 16 x 96 notes   wav", digest(w2), "logmel", digest(f2))


if __name__ == "__main__":
    main()
