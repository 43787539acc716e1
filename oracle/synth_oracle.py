"""CPU restatement of the reference synthetiser in NumPy float32 (FX chain: the reference's draws and position, DSP from
``oracle/fx_oracle.c`` - parity of that DSP with pedalboard is unpinned, see there).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  Pinned against the running
reference by tests/test_oracle_vs_reference.py (container) and against the
fixtures in tests/golden/ (everywhere).

Follows ``/root/reference/modules/synthetiser.py``:
  __call__ 255-292, drum_rendering 214-239, _vel_to_vol 204-212,
  random_choice_timbre 192-202, tolerance_thr_to_h5_group 171-190,
  VolumeMixer.init_tracks / instrument_mixer 146-156.
Float arithmetic is done in the dtype of the note array (float32 for lists and
float32 arrays - what ``torch.tensor(notes)`` gives), so the integer indices it
yields are the reference's, bit for bit.
"""
from __future__ import annotations

import math
import random as _random

import numpy as np

# utils/mapping_utils.py:56-84 (pitch -> ADTOF class) and :97-106 (class -> label)
_CLASS = dict(zip(range(35, 62), [35, 35, 38, 38, 38, 38, 41, 42, 42, 42, 41, 48, 41, 48, 48, 42, 48, 52,
                                  61, 61, 61, 61, 61, 58, 61, 61, 61]))
_LABEL = {35: "BD", 38: "SD", 41: "TT", 42: "HH", 48: "CY + RD", 52: "Cowbell", 58: "Claves", 61: "Other"}
_INVERSE = {35: [35, 36], 38: [37, 38, 39, 40], 41: [41, 45, 47], 42: [42, 43, 44, 50],
            48: [46, 48, 49, 51], 52: [52], 58: [58], 61: [53, 54, 55, 56, 57, 59, 60]}
_GAIN = {"BD": 1.0, "SD": 1.0, "TT": 1.0, "HH": 0.7, "CY + RD": 0.7, "Cowbell": 0.7, "Claves": 0.7, "Other": 1.0}
_GROUP_OF = {1.0: "gold", 0.9: "100-90", 0.8: "90-80", 0.7: "80-70", 0.6: "70-60", 0.5: "60-50",
             0.4: "50-40", 0.3: "40-30", 0.2: "30-20", 0.1: "20-10", 0.0: "10-0"}


def threshold_groups(tau: float):
    out, t = [], 1.0
    while t >= math.floor(tau * 10) / 10:       # :187
        out.append(_GROUP_OF[round(t, 1)])
        t -= 0.1
    return out


def vel_to_vol(v, dt):
    """:204-212 with every intermediate rounded to ``dt``."""
    if v == 0:
        return dt(0)
    x = dt(min(max(v, dt(0)), dt(127))) / dt(127.0)
    return dt(dt(0.1) + dt(dt(dt(0.9) * dt(dt(np.power(dt(6), x)) - dt(1))) / dt(5)))


def _normal(std, mean, high, low, generator=None):
    """utils/utils.py:266-269 ``draw_from_normal_distribution``, the same torch calls."""
    import torch
    z = torch.randn(1, generator=generator)
    return torch.clamp(torch.clamp(z * std + mean, -1.0, 1.0).abs() * high, low, high).item()


def draw_board(cfg: dict, rng=_random, generator=None) -> list:
    """``BoardChain.get_board`` (:81-87) -> ``[(plugin name, kwargs)]`` in board order; ``rng`` is the ``random``
    stream (coins, reverb), ``generator`` torch's (compressor, limiter; None = the global one)."""
    board = []
    if rng.random() < cfg["use_reverb_prob"]:                                          # :44-61
        room, damping, wet = rng.uniform(0.2, 0.8), rng.uniform(0.2, 0.8), rng.uniform(0.1, 0.4)
        dry, width = 1 - wet, rng.uniform(0.6, 1.0)
        board.append(("Reverb", dict(room_size=room, damping=damping, wet_level=wet, dry_level=dry, width=width)))
    if rng.random() < cfg["use_compression_prob"]:                                     # :63-75
        thr = -_normal(0.15, 0.5, 10, 0, generator)
        ratio = _normal(0.15, 0.5, 10, 1.0, generator)
        attack = _normal(0.05, 0.1, 1000, 0, generator)
        release = _normal(0.15, 0.2, 1000, 0, generator)
        board.append(("Compressor", dict(threshold_db=thr, ratio=ratio, attack_ms=attack, release_ms=release)))
    if rng.random() < cfg["use_limiter_prob"]:                                         # :77-79
        board.append(("Limiter", dict(threshold_db=-_normal(0.2, 0.4, 3, 0, generator))))
    return board


def apply_board(wav: np.ndarray, board: list, sample_rate: int) -> np.ndarray:
    """``_add_fx`` (:121-137) with the FX oracle's DSP (``oracle/fx_oracle.c``; parity with pedalboard unpinned)."""
    from . import fx_oracle
    for name, kw in board:
        kw = {k: float(np.float32(v)) for k, v in kw.items()}                          # pybind: python float -> C++ float
        if name == "Reverb":
            wav = fx_oracle.reverb(wav, sample_rate, kw["room_size"], kw["damping"], kw["wet_level"], kw["dry_level"],
                                   kw["width"])
        elif name == "Compressor":
            wav = fx_oracle.compressor(wav, sample_rate, kw["threshold_db"], kw["ratio"], kw["attack_ms"], kw["release_ms"])
        else:
            wav = fx_oracle.limiter(wav, sample_rate, kw["threshold_db"])
    return wav


def render(notes, cfg: dict, bank: dict, rng=_random, trace: list | None = None, raw: bool = False, generator=None,
           boards: list | None = None):
    """``SynthDrum(cfg)(notes)`` -> float32 waveform.  ``bank[pitch][group][name]`` is the
    HDF5 tree as nested dicts.  If ``trace`` is a list, one dict per note is appended
    (start, copy length, chosen paths, mixup, volume) for the bit-exact index checks.
    ``raw``: stop before the FX coin and return ``(instrument sum, max_volume)``.  With ``use_fx_prob > 0`` a coin hit
    draws the board like the reference and applies the FX oracle; ``boards`` (a list) receives the board or None."""
    sr = cfg["sample_rate"]
    if len(notes) == 0:
        return np.zeros(int(cfg["input_sec"] * sr), np.float32)                       # :257-258
    a = np.asarray(notes)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float32)
    if isinstance(notes, (list, tuple)) and not any(isinstance(r, np.ndarray) for r in notes):
        a = a.astype(np.float32)                                                       # torch.tensor(list of floats)
    dt = a.dtype.type
    end = dt(a[:, 1].max()) + dt(0.1)                                                  # :262
    length = int(cfg["input_sec"] * sr) if end < dt(cfg["input_sec"]) else int(end * dt(sr))  # :243
    tracks, picked, vmax = {}, {}, dt(0)
    groups = threshold_groups(cfg["similarity_threshold"])

    def pick(pitch):                                                                   # :192-202
        if cfg["ADTOF_mapping"]:
            pitch = rng.choice(_INVERSE[pitch])
        ok = [g for g in groups if str(int(pitch)) in bank and g in bank[str(int(pitch))]]
        g = rng.choice(ok)
        name = rng.choice(sorted(bank[str(int(pitch))][g].keys()))
        return f"{int(pitch)}/{g}/{name}"

    for on, off, pitch, vel in a:
        vmax = max(vmax, vel)
        if not (35 <= pitch <= 61 and off >= on):
            raise ValueError(f"Invalid note: {[on, off, pitch, vel]}")
        inst = int(pitch)
        if inst not in picked:
            picked[inst] = (pick(inst), pick(inst))
        tracks.setdefault(inst, np.zeros(length, np.float32))
        pa, pb = (p.split("/") for p in picked[inst])
        main = np.asarray(bank[pa[0]][pa[1]][pa[2]], np.float32)
        sub = np.asarray(bank[pb[0]][pb[1]][pb[2]], np.float32)
        mix = rng.uniform(0, cfg["mixup_range"])                                       # :217
        n = max(len(main), len(sub))
        m = np.zeros(n, np.float32); m[: len(main)] = main
        s = np.zeros(n, np.float32); s[: len(sub)] = sub
        vol = vel_to_vol(vel, dt)
        shot = m * np.float32(1 - mix) + np.float32(mix) * s                           # :223
        shot = shot / np.abs(shot).max()                                               # :225
        shot = shot * np.float32(vol)                                                  # :227
        start = int(on * dt(sr))                                                       # :229
        keep = min(n, length - start)                                                  # :232-237
        tracks[inst][start: start + keep] += shot[:keep]
        if trace is not None:
            trace.append(dict(start=start, len=keep, main=picked[inst][0], sub=picked[inst][1],
                              mixup=mix, vol=float(vol), pitch=inst))
    wav = np.zeros(length, np.float32)                                                 # :149-156
    for inst, trk in tracks.items():
        key = inst if cfg["ADTOF_mapping"] else _CLASS[inst]
        wav += trk * np.float32(_GAIN[_LABEL[key]])
    if raw:
        return wav, np.float32(vel_to_vol(vmax, dt))
    board = None
    if rng.random() < cfg["use_fx_prob"]:                                              # :154
        board = draw_board(cfg, rng, generator)
        wav = apply_board(wav, board, sr)
    if boards is not None:
        boards.append(board)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (wav / np.abs(wav).max() * np.float32(vel_to_vol(vmax, dt))).astype(np.float32)


def bucket_tiles(starts, lens, tile: int, n_tiles: int):
    """tile -> list of event ids whose [start, start+len) touches it (plain loops)."""
    out = [[] for _ in range(n_tiles)]
    for e, (s, n) in enumerate(zip(starts, lens)):
        if n <= 0:
            continue
        for t in range(s // tile, (s + n - 1) // tile + 1):
            out[t].append(e)
    return out


def collate(wavs):
    """pad_sequence(batch_first=True, padding_value=0.0) - data_modules/train_dataset.py:53."""
    out = np.zeros((len(wavs), max(len(w) for w in wavs)), np.float32)
    for i, w in enumerate(wavs):
        out[i, : len(w)] = w
    return out


def chunk_audio(wav: np.ndarray, chunk: int) -> np.ndarray:
    """inference.py:35-48 - non-overlapping chunks, last one zero-padded."""
    n = wav.shape[-1]
    k = -(-n // chunk)
    out = np.zeros((k, chunk), np.float32)
    for i in range(k):
        piece = wav[i * chunk: min((i + 1) * chunk, n)]
        out[i, : len(piece)] = piece
    return out
