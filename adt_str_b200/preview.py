"""The reference's drum preview renderer on the GPU (reference ``utils/drum_audio_render.py:130-194``).

``synthesize_drums_procedural`` sums one one-shot per note - the note's pitch picks it, ``clip(velocity, 1, 127) / 127``
scales it, ``int(onset * sample_rate)`` places it - and brings the sum under 0.98 full scale; ``render_drum_preview``
wraps it (and optionally writes the MIDI file, which needs ``pretty_midi`` like the reference).  Here the sum is one
``adtfe_render`` call: the one-shots form a small bank on the device, every note is an event of the tile mixer (its own
group, so that the accumulation runs in note order like the reference's ``buf[i0:i0+n] += hit[:n] * g``), and the
row is left raw (``ADTFE_SEG_RAW``) for the 0.98 limiter.

The mixer scales an event by ``ca * gain / peak`` with ``peak = max |ca * a + cb * b|`` of the note's own mix.  With
``ca = g``, ``cb = 0``, both one-shots the note's and ``gain = fl(g * max|a|)`` the peak is ``fl(g * max|a|)`` too, the
quotient exactly 1 and the coefficient exactly ``g``: what differs from the reference is one rounding per addition
(the mixer accumulates with fused multiply-adds), within 1e-6 of full scale - the tolerance of the test.

There is no CPU path: without ``libadtfe.so`` or an sm_100 device the constructor raises.
"""
from __future__ import annotations

from typing import Mapping, Optional, Tuple, Union

import numpy as np
import torch

from .bank import OneShotBank
from .midi_tokenizer import GM_STANDARD_TO_CUSTOM
from .planner import EVENT_DTYPE, SegmentPlan, assemble

SEG_RAW = 2   # ADTFE_SEG_RAW: the row keeps the raw mix


class PreviewRenderer:
    """``oneshots``: GM-custom pitch -> mono float32 waveform at ``sample_rate`` (what the reference caches from
    ``one-shot-rendering/<pitch>/*.wav``, ``drum_audio_render.py:74-114``)."""

    def __init__(self, oneshots: Mapping[int, np.ndarray], sample_rate: int, device: Optional[torch.device] = None):
        from .config import setting_1
        from .synthetiser import SynthDrum
        self.sample_rate = int(sample_rate)
        self.pitches = sorted(int(p) for p in oneshots)
        nested = {str(p): {"gold": {"hit": np.asarray(oneshots[p], np.float32).reshape(-1)}} for p in self.pitches}
        self.bank = OneShotBank.from_nested(nested)
        self.id_of = {p: self.bank.group_range(p, "gold")[0] for p in self.pitches}
        self.peak_of = {p: np.float32(np.abs(nested[str(p)]["gold"]["hit"]).max(initial=0.0)) for p in self.pitches}
        self.synth = SynthDrum(setting_1(sample_rate=self.sample_rate), bank=self.bank, device=device)

    def synthesize(self, notes, num_samples: int, apply_mapping: bool = True) -> torch.Tensor:
        """-> ``(num_samples,)`` float32 on the device."""
        arr = notes.detach().cpu().numpy() if isinstance(notes, torch.Tensor) else np.asarray(notes, dtype=np.float64)
        arr = arr.reshape(-1, 4).astype(np.float64)
        dev = self.synth.device
        max_s = num_samples / float(self.sample_rate)
        rows = []
        for onset, _off, pitch, vel in arr:                                      # the reference's skips, :148-163
            pitch = int(pitch)
            if onset >= max_s:
                continue
            i0 = int(onset * self.sample_rate)
            if i0 >= num_samples:
                continue
            p = GM_STANDARD_TO_CUSTOM.get(pitch, pitch) if apply_mapping else pitch
            if p not in self.id_of or self.peak_of[p] == 0:                     # no sample (or a silent one: it adds nothing)
                continue
            n = min(int(self.bank.lengths[self.id_of[p]]), num_samples - i0)
            if n <= 0:
                continue
            g = np.float32(float(np.clip(vel if vel > 1.0 else vel * 127.0, 1.0, 127.0)) / 127.0)   # :166
            rows.append((i0, n, self.id_of[p], g, np.float32(g * self.peak_of[p])))
        if not rows or num_samples <= 0:
            return torch.zeros(max(num_samples, 0), dtype=torch.float32, device=dev)
        ev = np.zeros(len(rows), EVENT_DTYPE)
        ev["start"] = [r[0] for r in rows]
        ev["len"] = [r[1] for r in rows]
        ev["main_id"] = ev["sub_id"] = [r[2] for r in rows]
        ev["ca"] = [r[3] for r in rows]
        ev["cb"] = 0.0
        ev["gain"] = [r[4] for r in rows]
        mix_len = self.bank.lengths[ev["main_id"]].astype(np.int32)
        seg = SegmentPlan(int(num_samples), SEG_RAW, 1.0, ev, mix_len, np.arange(len(rows) + 1, dtype=np.int32))
        plan = assemble([seg], self.bank)
        buf = self.synth.render_plan(plan)[0, :num_samples]
        peak = buf.abs().max()
        # buf *= min(1.0, 0.98 / peak) when peak > 1e-6 (:169-171); an all-silent sum is left as it is
        scale = torch.where(peak > 1e-6, torch.clamp(0.98 / peak, max=1.0), torch.ones_like(peak))
        return buf * scale


def synthesize_drums_procedural(notes, num_samples: int, sample_rate: int, oneshots: Mapping[int, np.ndarray],
                                apply_mapping: bool = True) -> np.ndarray:
    """The reference's signature plus the one-shots (it reads them from disk); returns a host float32 array like it."""
    return PreviewRenderer(oneshots, sample_rate).synthesize(notes, num_samples, apply_mapping).cpu().numpy()


def render_drum_preview(notes, num_samples: int, sample_rate: int, oneshots: Mapping[int, np.ndarray],
                        midi_path: Optional[str] = None, apply_mapping: bool = False) -> Tuple[torch.Tensor, str]:
    """``render_drum_preview`` (:176-194): optionally the MIDI file, then the one-shot rendering -> ``(waveform, "oneshot")``."""
    if midi_path is not None:
        save_drum_midi(notes, midi_path)
    wav = PreviewRenderer(oneshots, sample_rate).synthesize(notes, num_samples, apply_mapping)
    return wav.cpu(), "oneshot"


def save_drum_midi(notes: Union[np.ndarray, torch.Tensor], path) -> None:
    """``save_drum_midi`` / ``notes_to_pretty_midi`` (:31-66); needs ``pretty_midi``, like the reference."""
    try:
        import pretty_midi
    except ImportError as exc:   # the reference raises RuntimeError("pretty_midi is required for MIDI export.")
        raise RuntimeError("pretty_midi is required for MIDI export.") from exc
    arr = notes.detach().cpu().numpy() if isinstance(notes, torch.Tensor) else np.asarray(notes, dtype=np.float64)
    pm = pretty_midi.PrettyMIDI()
    inst = pretty_midi.Instrument(program=0, is_drum=True, name="ADT drums")
    for onset, offset, pitch, vel in arr.reshape(-1, 4):
        onset, offset, vel = float(onset), float(offset), float(vel)
        if offset <= onset:
            offset = onset + 0.05
        v = int(round(vel * 127)) if vel <= 1.0 else int(round(vel))
        inst.notes.append(pretty_midi.Note(velocity=int(np.clip(v, 1, 127)), pitch=int(np.clip(int(pitch), 0, 127)),
                                           start=onset, end=offset))
    pm.instruments.append(inst)
    pm.write(str(path))
