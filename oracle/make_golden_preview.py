#!/usr/bin/env python3
"""Freeze the UNMODIFIED reference's preview renderer (utils/drum_audio_render.py:130-173) as tests/golden/preview.npz.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_preview.py

The reference keeps its one-shots in a module cache filled from WAV files on first use; the script fills that cache
with synthetic one-shots instead (nothing else is touched), runs ``synthesize_drums_procedural`` with and without the
pitch mapping, REFUSES to write unless the oracle restatement agrees bit for bit, and stores inputs and outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import preview_oracle  # noqa: E402
import utils.drum_audio_render as ref  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    sr, num_samples = 24000, 24000 * 2 + 123
    pitches = [35, 36, 38, 41, 42, 44, 46, 48, 52, 54, 60]          # GM-custom pitches that have a sample
    oneshots = {}
    for p in pitches:
        n = int(rng.integers(300, 12000))
        t = np.arange(n, dtype=np.float32) / np.float32(n)
        oneshots[p] = (rng.standard_normal(n).astype(np.float32) * np.exp(np.float32(-5.0) * t) *
                       np.float32(rng.uniform(0.2, 1.5))).astype(np.float32)
    e = 120
    onset = np.sort(rng.uniform(0.0, 2.2, e))                       # some notes start past the end of the buffer
    pitch = rng.choice(list(range(35, 82)) + [20, 90], e)           # mapped, unmapped and sample-less pitches
    vel = np.where(rng.random(e) < 0.5, rng.uniform(0.0, 1.0, e), rng.integers(0, 140, e).astype(np.float64))
    notes = np.stack([onset, onset + 0.1, pitch.astype(np.float64), vel], 1)
    ref._ONESHOT_CACHE.clear()
    ref._ONESHOT_CACHE.update(oneshots)
    out = {}
    for tag, mapping in (("mapped", True), ("plain", False)):
        want = ref.synthesize_drums_procedural(notes, num_samples, sr, apply_mapping=mapping)
        got = preview_oracle.synthesize_drums_procedural(notes, num_samples, sr, oneshots, apply_mapping=mapping)
        assert want.dtype == np.float32 and np.array_equal(want, got), tag
        out["wav_" + tag] = want
    quiet = ref.synthesize_drums_procedural(notes[:0], 1000, sr)
    assert not quiet.any()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preview.npz"), notes=notes, sample_rate=sr,
                        num_samples=num_samples, pitches=np.array(pitches), **{f"os_{p}": oneshots[p] for p in pitches}, **out)
    print("wrote tests/golden/preview.npz:", {k: float(np.abs(v).max()) for k, v in out.items()})


if __name__ == "__main__":
    main()
