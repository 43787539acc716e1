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
from .planner import (EVENT_DTYPE, FX_DTYPE, PEAK_ITEM_DTYPE, SEGMENT_DTYPE, RenderPlan, fill_fx_normals, plan_batch,
                      similarity_groups)

_ERRORS = {1: ValueError, 2: IndexError, 3: KeyError}


class PackedGroup:
    """What the pipeline keeps of a group planned and packed natively (``NativePlanner.plan_group_into``): the
    geometry ``FrontEnd.run_plan_host`` and ``GroupResult`` read - the records themselves are in the pinned blob."""
    mel_rows = True   # ragged log-mel rows are part of the blob

    def __init__(self, n_seg, ld_wav, mel_total_rows, wave_lengths, batch_ptr, batch_samples, batch_frames, n_events):
        self.n_seg, self.ld_wav, self.mel_total_rows = n_seg, ld_wav, mel_total_rows
        self.wave_lengths, self.batch_ptr = wave_lengths, batch_ptr
        self.batch_samples, self.batch_frames, self.n_events = batch_samples, batch_frames, n_events


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
        _lib.check(lib.adtfe_planner_set_fx(h, float(config.use_reverb_prob), float(config.use_compression_prob),
                                            float(config.use_limiter_prob)), "adtfe_planner_set_fx")

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
        # torch.tensor(notes) is float32 only for plain python numbers with at least one float among them: numpy
        # scalars (np.float64 is a float subclass) make it float64 and all-int rows int64, and the reference then
        # does its index arithmetic in that type - those inputs take the general planner
        if isinstance(notes, (list, tuple)):
            kinds = {type(v) for r in notes for v in r}
            if kinds <= {int, float} and float in kinds:
                return np.asarray(notes, np.float32).reshape(-1, 4)
        return None

    def plan_batch(self, batch_notes: Sequence, rng=_random, ld_wav: Optional[int] = None, generator=None) -> RenderPlan:
        """``generator``: torch generator of the FX chain's normal draws (None = the global one, like the reference)."""
        arrays = [self._as_f32(n) for n in batch_notes]
        if any(a is None for a in arrays) or not hasattr(rng, "getstate"):
            return plan_batch(batch_notes, self.config, self.bank, rng, ld_wav, generator)
        for a in arrays:
            if a.ndim != 2 or a.shape[1] != 4:
                raise ValueError(f"notes must have shape (N, 4), got {tuple(a.shape)}")
        n_seg = len(arrays)
        counts = np.array([len(a) for a in arrays], np.int32)
        flat = np.ascontiguousarray(np.concatenate(arrays)) if n_seg else np.zeros((0, 4), np.float32)
        version, internal, gauss = rng.getstate()
        state = np.array(internal, np.uint32)
        out = np.zeros(9, np.int64)
        info = np.zeros(2, np.int32)
        rc = self.lib.adtfe_planner_plan(self.handle, flat.ctypes.data, counts.ctypes.data, n_seg, state.ctypes.data,
                                         int(ld_wav or 0), out.ctypes.data, info.ctypes.data)
        rng.setstate((version, tuple(state.tolist()), gauss))
        if rc > 0:
            seg, note = int(info[0]), int(info[1])
            what = f"Invalid note: {arrays[seg][note]}" if rc == 1 and note >= 0 else \
                {2: "Cannot choose from an empty sequence", 3: f"segment {seg} note {note}"}[rc] \
                if rc != 1 else "Invalid note"
            raise _ERRORS[rc](what)
        _lib.check(rc, "adtfe_planner_plan")
        n_ev, n_grp, _, tps, n_pw, n_te, ld, _, n_fx = (int(x) for x in out)
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
        plan = RenderPlan(n_seg, ld, tps, segments, events, mix_len, group_ptr, tile_ptr, tile_events, peak_work,
                          segments["len"].astype(np.int64))
        plan.sample_rate = int(self.config.sample_rate)
        if n_fx:   # the planner drew the coins and the reverb from `random`; the dynamics parameters are torch's
            fx = np.empty(n_fx, FX_DTYPE)
            _lib.check(self.lib.adtfe_planner_export_fx(self.handle, fx.ctypes.data), "adtfe_planner_export_fx")
            plan.fx = fill_fx_normals(fx, generator)
        return plan

    # ---- a whole group of batches planned and packed into a pinned blob without the interpreter in between
    def plan_group(self, group: Sequence[Sequence], mt_state: np.ndarray) -> Optional[np.ndarray]:
        """Plan all segments of ``group`` (batches of float32 (N, 4) note arrays) with the MT19937 state
        ``mt_state`` (uint32[625], advanced in place - the caller's private stream).  The plan stays inside the
        native planner until ``pack_group``.  Returns the counts of ``adtfe_planner_plan``, or None when the notes
        are not plain float32 arrays (the caller then takes the general path; the state is untouched)."""
        flat = [n for b in group for n in b]
        for a in flat:
            if type(a) is not np.ndarray or a.dtype != np.float32 or a.ndim != 2 or a.shape[1] != 4:
                return None
        n_seg = len(flat)
        if n_seg == 0 or any(len(b) == 0 for b in group):
            return None
        counts = np.fromiter(map(len, flat), np.int32, n_seg)
        notes = np.concatenate(flat) if n_seg > 1 else np.ascontiguousarray(flat[0])
        out = np.zeros(9, np.int64)
        info = np.zeros(2, np.int32)
        rc = self.lib.adtfe_planner_plan(self.handle, notes.ctypes.data, counts.ctypes.data, n_seg,
                                         mt_state.ctypes.data, 0, out.ctypes.data, info.ctypes.data)
        if rc > 0:
            seg, note = int(info[0]), int(info[1])
            raise _ERRORS[rc](f"Invalid note: {flat[seg][note]}" if rc == 1 and note >= 0 else
                              f"segment {seg} note {note} (planner status {rc})")
        _lib.check(rc, "adtfe_planner_plan")
        return out

    def pack_group(self, sizes: Sequence[int], hop: int, wpi: int, chunk_batches: int, host_ptr: int, capacity: int):
        """The plan of the last ``plan_group`` as ``len(sizes)`` collated batches, written into the blob at
        ``host_ptr`` (``adtfe_planner_pack_batches``).  Returns ``(status, shape, bytes needed, chunks, width, frames,
        fx offset)``; status -3 = the blob is too small (nothing written)."""
        sizes = np.ascontiguousarray(sizes, np.int32)
        nb = len(sizes)
        chunks = np.zeros((nb + 1) * 4, np.int32)
        width, frames = np.zeros(nb, np.int64), np.zeros(nb, np.int64)
        shape = _lib.Plan()
        need, fx_off = C.c_size_t(), C.c_size_t()
        rc = self.lib.adtfe_planner_pack_batches(self.handle, sizes.ctypes.data, nb, int(chunk_batches), int(hop),
                                                 int(wpi), host_ptr, capacity, C.byref(shape), chunks.ctypes.data,
                                                 width.ctypes.data, frames.ctypes.data, C.byref(need), C.byref(fx_off))
        if rc not in (0, -3):
            _lib.check(rc, "adtfe_planner_pack_batches")
        return rc, shape, need.value, chunks, width, frames, fx_off.value

    def plan_group_into(self, group: Sequence[Sequence], mt_state: np.ndarray, acquire, hop: int, wpi: int,
                        chunk_batches: int = 1, generator=None) -> Optional[PackedGroup]:
        """``plan_group``, then ``buf = acquire()`` (the ``PlanBuffers`` whose pinned blob receives the plan - the
        caller may block there until a buffer set is free), then ``pack_group`` into it."""
        out = self.plan_group(group, mt_state)
        if out is None:
            return None
        buf = acquire()
        sizes = np.fromiter(map(len, group), np.int32, len(group))
        rc, shape, need, chunks, width, frames, fx_off = self.pack_group(sizes, hop, wpi, chunk_batches,
                                                                         buf.host.data_ptr(), buf.host.numel())
        if rc == -3:   # ADTFE_ERR_WORKSPACE: grow the pinned blob and pack again
            buf.reserve(need)
            rc, shape, need, chunks, width, frames, fx_off = self.pack_group(sizes, hop, wpi, chunk_batches,
                                                                             buf.host.data_ptr(), buf.host.numel())
        _lib.check(rc, "adtfe_planner_pack_batches")
        if shape.n_fx:   # the FX records sit in the blob with NaN dynamics parameters: fill them in place (torch's stream)
            fx = buf.host.numpy()[fx_off: fx_off + FX_DTYPE.itemsize * shape.n_fx].view(FX_DTYPE)
            fill_fx_normals(fx, generator)
        buf.adopt(shape, need, chunks)
        n_seg = int(out[2])
        o = buf.offsets[1]
        seg = np.frombuffer(buf.host.numpy()[o: o + SEGMENT_DTYPE.itemsize * n_seg].tobytes(), SEGMENT_DTYPE)
        ptr = np.zeros(len(sizes) + 1, np.int64)
        np.cumsum(sizes, out=ptr[1:])
        return PackedGroup(n_seg, int(out[6]), int(shape.mel_total_rows), seg["len"].astype(np.int64), ptr, width,
                           frames, int(out[0]))
