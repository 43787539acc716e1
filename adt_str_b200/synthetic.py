"""Synthetic one-shot banks and MIDI event streams.

There is no dataset in the container, so every test and benchmark runs on
these generators.  They are deterministic (``numpy.random.default_rng`` with a
fixed seed) and broadband on purpose: decaying white noise keeps every mel band
well above the fp32 rounding floor, which is what makes the 1e-4 log-mel
tolerance meaningful (SURVEY §7 "hard parts").
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .bank import OneShotBank

GROUPS_TAU_08 = ("gold", "100-90", "90-80")  # similarity_threshold 0.8
BANK_PITCHES = tuple(range(35, 61))          # GM-custom 35..60, 26 classes

# pitch distribution of the event streams
_PITCH_P = {36: .20, 38: .20, 42: .30, 44: .05, 41: .08 / 3, 45: .08 / 3, 47: .08 / 3, 46: .05, 48: .10}


def _pitch_table():
    rest = [p for p in BANK_PITCHES if p not in _PITCH_P]
    pitches = list(_PITCH_P) + rest
    probs = list(_PITCH_P.values()) + [0.02 / len(rest)] * len(rest)
    probs = np.asarray(probs, np.float64)
    return np.asarray(pitches, np.int64), probs / probs.sum()


def make_bank(n_oneshots: int = 10_000, sample_rate: int = 24_000, seed: int = 0,
              min_len: int | None = None, max_len: int | None = None,
              pitches: Sequence[int] = BANK_PITCHES, groups: Sequence[str] = GROUPS_TAU_08,
              fixed_len: int | None = None) -> OneShotBank:
    """``n_oneshots`` decaying-noise one-shots spread evenly over pitch x group.

    Length is log-uniform in [0.05 s, 2 s] (1 200..48 000 samples at 24 kHz) unless
    ``fixed_len`` is given; each one-shot is peak-normalised to exactly 1.0 like
    reference ``data_modules/convert_augmented_to_hdf5.py:102``.
    """
    rng = np.random.default_rng(seed)
    min_len = int(0.05 * sample_rate) if min_len is None else min_len
    max_len = int(2.0 * sample_rate) if max_len is None else max_len
    cells = [(p, g) for p in pitches for g in groups]
    nested: Dict[str, Dict[str, Dict[str, np.ndarray]]] = {}
    for i in range(n_oneshots):
        p, g = cells[i % len(cells)]
        n = fixed_len or int(round(np.exp(rng.uniform(np.log(min_len), np.log(max_len)))))
        t = np.arange(n, dtype=np.float32) / np.float32(n)
        x = rng.standard_normal(n, dtype=np.float32) * np.exp(np.float32(-6.0) * t)
        x /= np.abs(x).max()
        nested.setdefault(str(p), {}).setdefault(g, {})[f"os{i:05d}"] = x.astype(np.float32)
    return OneShotBank.from_nested(nested)


def make_segments(n_segments: int, seed: int = 1, input_sec: float = 2.56, mean_events: float = 32.0,
                  max_events: int = 96, empty_fraction: float = 0.05) -> List[np.ndarray]:
    """Random training-shape note lists, one float32 (N, 4) array per segment:
    ``[onset s, offset s, pitch, velocity]`` with offset = onset + 0.1 (reference
    ``data_modules/midi_parser.py:116-126``), velocity ~ randint(10, 127)
    (``modules/midi_tokenizer.py:46``) and ``empty_fraction`` empty segments
    (``configs/train/setting-1.yaml:37``)."""
    rng = np.random.default_rng(seed)
    pitches, probs = _pitch_table()
    out = []
    for _ in range(n_segments):
        if rng.random() < empty_fraction:
            out.append(np.zeros((0, 4), np.float32))
            continue
        e = int(np.clip(rng.poisson(mean_events), 1, max_events))
        onset = np.sort(rng.uniform(0.0, input_sec - 0.11, e)).astype(np.float32)
        notes = np.empty((e, 4), np.float32)
        notes[:, 0] = onset
        notes[:, 1] = onset + np.float32(0.1)
        notes[:, 2] = rng.choice(pitches, size=e, p=probs)
        notes[:, 3] = rng.integers(10, 127, e)
        out.append(notes)
    return out


def make_dense_segment(input_sec: float = 2.56, bpm: float = 180.0) -> np.ndarray:
    """Dense-polyphony stress (BASELINE config 4): 16th-note closed hat and ride,
    8th-note kick/snare and two 8-note 32nd tom fills, about 100 events."""
    sixteenth = 60.0 / bpm / 4.0
    rows = []
    n16 = int((input_sec - 0.11) / sixteenth)
    for i in range(n16):
        t = i * sixteenth
        rows.append((t, 42, 60 + (i * 7) % 60))
        rows.append((t, 48, 50 + (i * 11) % 70))
        if i % 2 == 0:
            rows.append((t, 36 if (i // 2) % 2 == 0 else 38, 100 + (i % 20)))
    for base in (0.8, 2.0):
        for j in range(8):
            rows.append((base + j * sixteenth / 2.0, (41, 45, 47)[j % 3], 90 + j))
    rows.sort(key=lambda r: r[0])
    notes = np.array([(t, 0.0, p, v) for t, p, v in rows], np.float32)
    notes[:, 1] = notes[:, 0] + np.float32(0.1)
    return notes


def make_long_form(duration_sec: float = 600.0, events_per_sec: float = 12.0, seed: int = 2) -> np.ndarray:
    """One long note list (BASELINE config 5: a 10-minute mix rendered by a single
    ``SynthDrum`` call as in reference ``inference.py:146``)."""
    rng = np.random.default_rng(seed)
    pitches, probs = _pitch_table()
    e = int(duration_sec * events_per_sec)
    onset = np.sort(rng.uniform(0.0, duration_sec - 0.2, e)).astype(np.float32)
    notes = np.empty((e, 4), np.float32)
    notes[:, 0] = onset
    notes[:, 1] = onset + np.float32(0.1)
    notes[:, 2] = rng.choice(pitches, size=e, p=probs)
    notes[:, 3] = rng.integers(10, 127, e)
    return notes
