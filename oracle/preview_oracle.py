"""CPU restatement of the reference's preview renderer - TEST INFRASTRUCTURE ONLY (only tests/ may import it).

``synthesize_drums_procedural`` (reference ``utils/drum_audio_render.py:130-173``): every note adds its pitch's one-shot,
scaled by ``clip(velocity, 1, 127) / 127``, at ``int(onset * sample_rate)``; the sum is brought under 0.98 full scale.
The reference reads its one-shots from ``one-shot-rendering/<pitch>/*.wav`` into a module cache
(``get_oneshot_waveform`` ``:74-127``); here they are passed in.  Pinned on the running reference by
``oracle/make_golden_preview.py`` (bit-identical) and ``tests/golden/preview.npz``.
"""
from __future__ import annotations

import numpy as np

# GM standard -> GM custom drum pitch (reference utils/mapping_utils.py:3-51), pitches 35..81
GM_CUSTOM_OF = (35, 36, 37, 38, 39, 40, 41, 42, 41, 43, 41, 44, 45, 45, 46, 47, 48, 49, 48, 50, 51, 52, 46, 53, 48,
                54, 54, 54, 54, 54, 54, 54, 52, 52, 55, 55, 56, 56, 57, 57, 58, 58, 58, 59, 59, 60, 60)


def gm_custom(pitch: int) -> int:
    return GM_CUSTOM_OF[pitch - 35] if 35 <= pitch <= 81 else pitch      # mapping.get(pitch, pitch), :120


def synthesize_drums_procedural(notes, num_samples: int, sample_rate: int, oneshots: dict, apply_mapping: bool = True):
    arr = np.asarray(notes, dtype=np.float64)                             # :137-141
    if arr.size == 0:
        return np.zeros(num_samples, dtype=np.float32)
    buf = np.zeros(num_samples, dtype=np.float32)
    max_s = num_samples / float(sample_rate)
    for row in arr:                                                       # :147
        onset, pitch, vel = float(row[0]), int(row[2]), float(row[3])
        if onset >= max_s:
            continue
        i0 = int(onset * sample_rate)
        if i0 >= num_samples:
            continue
        hit = oneshots.get(gm_custom(pitch) if apply_mapping else pitch)  # :160, :116-127
        if hit is None:
            continue
        n = min(len(hit), num_samples - i0)
        if n > 0:
            g = float(np.clip(vel if vel > 1.0 else vel * 127.0, 1.0, 127.0)) / 127.0   # :166
            buf[i0: i0 + n] += hit[:n] * g
    peak = np.abs(buf).max()                                              # :169-171
    if peak > 1e-6:
        buf *= min(1.0, 0.98 / peak)
    return buf
