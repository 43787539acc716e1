"""ctypes front of the C++ host planner (csrc/planner.cpp).

Produces the same ``RenderPlan`` as ``planner.plan_batch`` - same RNG stream (the MT19937
state of ``random`` is handed over and written back), same indices - about two orders of
magnitude faster.  Falls back to the Python planner for inputs it does not cover (float64
note arrays, exotic rng objects)."""
from __future__ import annotations

import ctypes as C
import random as _random
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .bank import OneShotBank
from .config import SynthDrumConfig
from .mapping import ADTOF_INVERSE, PITCH_MAX, PITCH_MIN, instrument_gain
from .planner import EVENT_DTYPE, PEAK_ITEM_DTYPE, SEGMENT_DTYPE, RenderPlan, plan_batch, similarity_groups

_ERRORS = {1: ValueError, 2: IndexError, 3: KeyError, 4: NotImplementedError}


class NativePlanner:
    def __init__(self, config: SynthDrumConfig, bank: OneShotBank):
        self.config, self.bank = config, bank
        lib = _lib.load()
        groups = similarity_groups(config.similarity_threshold)
        gptr, gfirst, gcount, gain, iptr, ipitch = [0], [], [], [], [0], []
        for pitch in range(PITCH_MIN, PITCH_MAX + 1):
            for g in groups:
                if bank.has_group(pitch, g):
                    first, count = bank.group_range(pitch, g)
                    gfirst.append(first)
                    gcount.append(count)
            gptr.append(len(gfirst))
            try:
                gain.append(instrument_gain(pitch, config.ADTOF_mapping))
            except KeyError:
                gain.append(-1.0)
            ipitch.extend(ADTOF_INVERSE.get(pitch, []))
            iptr.append(len(ipitch))
        a = lambda x, dt: np.ascontiguousarray(np.asarray(x, dt))  # noqa: E731
        self._keep = [a(bank.lengths, np.int32), a(bank.offsets, np.int64), a(gptr, np.int32), a(gfirst, np.int32), a(gcount, np.int32),
                      a(gain, np.float32), a(iptr, np.int32), a(ipitch, np.int32)]
        k = self._keep
        h = C.c_void_p()
        _lib.check(lib.adtfe_planner_create(config.sample_rate, float(config.input_sec), float(config.mixup_range),
                                            float(config.use_fx_prob), int(bool(config.ADTOF_mapping)),
                                            k[0].ctypes.data, k[1].ctypes.data, len(bank), k[2].ctypes.data,
                                            k[3].ctypes.data, k[4].ctypes.data, k[5].ctypes.data, k[6].ctypes.data,
                                            k[7].ctypes.data, C.byref(h)), "adtfe_planner_create")
        self.handle, self.lib = h, lib

    def __del__(self):
        try:
            if self.handle:
                self.lib.adtfe_planner_destroy(self.handle)
        except Exception:
            pass

    @staticmethod
    def _as_f32(notes) -> Optional[np.ndarray]:
        """float32 (n, 4) view of one note list, or None when torch.tensor(notes) would not be float32."""
        if isinstance(notes, np.ndarray):
            if notes.dtype != np.float32:
                return None
            return notes.reshape(-1, 4) if notes.size else notes.reshape(0, 4)
        if len(notes) == 0:
            return np.zeros((0, 4), np.float32)
        if isinstance(notes, (list, tuple)) and all(isinstance(v, (int, float)) for r in notes for v in r):
            return np.asarray(notes, np.float32).reshape(-1, 4)   # python numbers -> torch default float32
        return None

    def plan_batch(self, batch_notes: Sequence, rng=_random, ld_wav: Optional[int] = None) -> RenderPlan:
        arrays = [self._as_f32(n) for n in batch_notes]
        if any(a is None for a in arrays) or not hasattr(rng, "getstate"):
            return plan_batch(batch_notes, self.config, self.bank, rng, ld_wav)
        for a in arrays:
            if a.ndim != 2 or a.shape[1] != 4:
                raise ValueError(f"notes must have shape (N, 4), got {tuple(a.shape)}")
        n_seg = len(arrays)
        counts = np.array([len(a) for a in arrays], np.int32)
        flat = np.ascontiguousarray(np.concatenate(arrays)) if n_seg else np.zeros((0, 4), np.float32)
        version, internal, gauss = rng.getstate()
        state = np.array(internal, np.uint32)
        out = np.zeros(8, np.int64)
        info = np.zeros(2, np.int32)
        rc = self.lib.adtfe_planner_plan(self.handle, flat.ctypes.data, counts.ctypes.data, n_seg, state.ctypes.data,
                                         int(ld_wav or 0), out.ctypes.data, info.ctypes.data)
        rng.setstate((version, tuple(state.tolist()), gauss))
        if rc > 0:
            seg, note = int(info[0]), int(info[1])
            what = f"Invalid note: {arrays[seg][note]}" if rc == 1 and note >= 0 else \
                {2: "Cannot choose from an empty sequence", 3: f"segment {seg} note {note}",
                 4: "the pedalboard FX chain (synthetiser.py:121-137) is outside the GPU path; set use_fx_prob=0"}[rc] \
                if rc != 1 else "Invalid note"
            raise _ERRORS[rc](what)
        _lib.check(rc, "adtfe_planner_plan")
        n_ev, n_grp, _, tps, n_pw, n_te, ld, _ = (int(x) for x in out)
        events = np.empty(n_ev, EVENT_DTYPE)
        mix_len = np.empty(n_ev, np.int32)
        group_ptr = np.empty(n_grp + 1, np.int32)
        segments = np.empty(n_seg, SEGMENT_DTYPE)
        tile_ptr = np.empty(n_seg * tps + 1, np.int32)
        peak_work = np.empty(n_pw, PEAK_ITEM_DTYPE)
        tile_events = np.empty(n_te, np.int32)
        _lib.check(self.lib.adtfe_planner_export(self.handle, events.ctypes.data, mix_len.ctypes.data,
                                                 group_ptr.ctypes.data, segments.ctypes.data, tile_ptr.ctypes.data,
                                                 peak_work.ctypes.data, tile_events.ctypes.data), "adtfe_planner_export")
        return RenderPlan(n_seg, ld, tps, segments, events, mix_len, group_ptr, tile_ptr, tile_events, peak_work,
                          segments["len"].astype(np.int64))
