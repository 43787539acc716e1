"""Host planner: notes + RNG draws -> flat event records bucketed by output tile.

This is the integer / float32 bookkeeping of the reference's per-note loop
(``modules/synthetiser.py:255-292``) separated from the audio arithmetic, so
that the GPU only sees plain arrays:

* RNG call order of the reference, reproduced draw for draw from the same
  ``random`` stream - per *new* instrument ``choice(groups)``, ``choice(keys)``
  for the main then the sub timbre (``:192-202, 274-281``); per note
  ``uniform(0, mixup_range)`` (``:217``); one final ``random()`` for the FX
  coin (``:154``).
* index rules in the dtype the reference computes them in (float32 unless the
  caller hands in float64 notes): ``note_start = int(onset * sr)`` (``:229``),
  ``wave_length = int(max(max_offset + 0.1, input_sec) * sr)`` (``:262, 243``),
  copy length ``min(max(len_a, len_b), wave_length - note_start)``
  (``:230-237``).  These must be bit-exact; tests/test_planner.py checks them
  against the running reference.
* velocity -> volume curve (``:204-212``) and the per-instrument gain
  (``:104-113, 152-153``) folded into one float32 gain per event.

Events are emitted in *track order* - instruments by first appearance, notes in
input order inside an instrument - which is the order the reference accumulates
them in (per-instrument tracks ``:290``, then ``instrument_mixer`` ``:149-153``).
The tile mixer adds them in exactly this order, so the output is deterministic.
"""
from __future__ import annotations

import math
import random as _random
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np
import torch

from .bank import OneShotBank
from .config import SynthDrumConfig
from .mapping import (ADTOF_INVERSE, PITCH_MAX, PITCH_MIN, SIMILARITY_GROUPS,
                      instrument_gain)

TILE = 2048       # output samples owned by one CTA of the tile mixer (ADTFE_TILE)

#: numpy view of ``adtfe_event`` in include/adtfe.h (32 bytes)
EVENT_DTYPE = np.dtype([("start", "<i4"), ("len", "<i4"), ("main_id", "<i4"), ("sub_id", "<i4"),
                        ("ca", "<f4"), ("cb", "<f4"), ("gain", "<f4"), ("seg", "<i4")])
#: numpy view of ``adtfe_segment`` (16 bytes)
SEGMENT_DTYPE = np.dtype([("len", "<i4"), ("flags", "<i4"), ("max_volume", "<f4"), ("first_event", "<i4")])
#: numpy view of ``adtfe_peak_item`` (40 bytes)
PEAK_ITEM_DTYPE = np.dtype([("a_off", "<i8"), ("b_off", "<i8"), ("la", "<i4"), ("lb", "<i4"), ("mix_len", "<i4"),
                            ("first_event", "<i4"), ("n_events", "<i4"), ("chunk", "<i4")])
#: numpy view of ``adtfe_mel_row`` (16 bytes)
MEL_ROW_DTYPE = np.dtype([("out_row", "<i8"), ("count", "<i4"), ("flags", "<i4")])
MEL_ROW_SILENT = 1  # ADTFE_MEL_ROW_SILENT: the row is all zeros (an empty segment), its frames are exact zeros
#: numpy view of ``adtfe_chunk`` (16 bytes)
CHUNK_DTYPE = np.dtype([("seg", "<i4"), ("event", "<i4"), ("peak_work", "<i4"), ("fx_row", "<i4")])
#: numpy view of ``adtfe_fx`` (48 bytes): the keyword arguments the reference hands to pedalboard
FX_DTYPE = np.dtype([("seg", "<i4"), ("flags", "<i4"), ("room_size", "<f4"), ("damping", "<f4"), ("wet_level", "<f4"),
                     ("dry_level", "<f4"), ("width", "<f4"), ("comp_threshold_db", "<f4"), ("comp_ratio", "<f4"),
                     ("comp_attack_ms", "<f4"), ("comp_release_ms", "<f4"), ("lim_threshold_db", "<f4")])
FX_REVERB, FX_COMPRESSOR, FX_LIMITER = 1, 2, 4   # ADTFE_FX_*
SEG_EMPTY = 0      # no notes: all-zero waveform of int(input_sec*sr) samples, no normalisation
SEG_NORMALISE = 1  # wav / max|wav| * max_volume (NaN when the mix is all zero, like the reference)


def similarity_groups(threshold: float) -> List[str]:
    """Sub-groups admitted by a similarity threshold, best first
    (reference ``synthetiser.py:168-190``; same float loop, so 0.8 -> 3 groups)."""
    floor = math.floor(threshold * 10) / 10
    level, groups = 1.0, []
    while level >= floor:
        groups.append(SIMILARITY_GROUPS[10 - int(round(round(level, 1) * 10))])
        level -= 0.1
    return groups


def velocity_to_volume(velocity: np.ndarray) -> np.ndarray:
    """``0.1 + 0.9 * (6**(clamp(v,0,127)/127) - 1) / 5`` in the dtype of ``velocity``;
    exactly 0 for v == 0 (reference ``synthetiser.py:204-212``)."""
    dt = velocity.dtype.type
    x = np.clip(velocity, dt(0), dt(127)) / dt(127.0)
    p = np.power(dt(6), x)
    vol = dt(0.1) + (dt(0.9) * (p - dt(1))) / dt(5)
    return np.where(velocity == 0, dt(0), vol).astype(velocity.dtype)


def notes_to_array(notes) -> np.ndarray:
    """(N, 4) array in the dtype ``torch.tensor(notes)`` would infer
    (reference ``synthetiser.py:259``): python floats -> float32, numpy/torch keep theirs."""
    if isinstance(notes, torch.Tensor):
        t = notes.detach().cpu()
    else:
        t = torch.tensor(notes)
    if not t.dtype.is_floating_point or t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    a = t.numpy()
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError(f"notes must have shape (N, 4), got {tuple(a.shape)}")
    return a


@dataclass
class SegmentPlan:
    wave_length: int
    flags: int
    max_volume: float
    events: np.ndarray            # EVENT_DTYPE, track order, seg field = 0
    mix_len: np.ndarray           # int32 per event: max(len_a, len_b) before truncation
    group_ptr: np.ndarray         # int32 (n_groups+1,): events of one instrument are contiguous
    fx: np.ndarray = None         # FX_DTYPE (1,) when the FX coin hit (seg field = 0), else None


def normal_draw(std: float, mean: float, high_bound: float, low_bound: float, generator=None) -> float:
    """``draw_from_normal_distribution`` of the reference (``utils/utils.py:266-269``), the same torch calls - so the
    same float32 arithmetic on the same draw of the same generator (``generator=None``: torch's global one, which is
    what the reference uses)."""
    z = torch.randn(1, generator=generator)
    return torch.clamp(torch.clamp(z * std + mean, -1.0, 1.0).abs() * high_bound, low_bound, high_bound).item()


# draw_from_normal_distribution(std, mean, high_bound, low_bound) of every parameter, in the order the reference draws
# them: _add_compression's four (threshold is negated), then _add_limiter's one (negated) - synthetiser.py:63-79
_COMP_DRAWS = (("comp_threshold_db", -1.0, 0.15, 0.5, 10.0, 0.0), ("comp_ratio", 1.0, 0.15, 0.5, 10.0, 1.0),
               ("comp_attack_ms", 1.0, 0.05, 0.1, 1000.0, 0.0), ("comp_release_ms", 1.0, 0.15, 0.2, 1000.0, 0.0))
_LIM_DRAWS = (("lim_threshold_db", -1.0, 0.2, 0.4, 3.0, 0.0),)


def fill_fx_normals(fx: np.ndarray, generator=None) -> np.ndarray:
    """The compressor and limiter parameters of FX records, drawn in the reference's order - record by record (one
    ``SynthDrum.__call__`` each), ``_add_compression``'s four draws then ``_add_limiter``'s one
    (``synthetiser.py:63-79,81-86``).  The native planner leaves them NaN: it owns the ``random`` stream, torch's
    generator lives here.

    The reference draws one value per call (``torch.randn(1)``, ``utils/utils.py:266-269``).  For fewer than 16
    elements ``torch.randn`` fills a tensor through the same scalar sampler, one element after the other (Box-Muller,
    the second value of a pair kept in the generator), so ``randn(k)`` with ``k <= 15`` IS the next ``k`` single
    draws - tests/test_fx.py holds the two against each other - and the float32 arithmetic behind the draw is
    elementwise.  Drawing a batch's parameters fifteen at a time instead of one at a time takes the interpreter out
    of the planner threads (1.25 ms of GIL-bound calls per batch of the stock FX configuration before)."""
    if len(fx) == 0:
        return fx
    flags = fx["flags"]
    comp, lim = (flags & FX_COMPRESSOR) != 0, (flags & FX_LIMITER) != 0
    # (record, parameter) of every draw, in draw order: row-major over [comp, comp, comp, comp, lim] per record
    rec, par = np.nonzero(np.stack([comp, comp, comp, comp, lim], 1))
    n = len(rec)
    if n == 0:
        return fx
    z = torch.empty(n, dtype=torch.float32)
    for i in range(0, n, 15):   # one sampler call per fifteen draws; the arithmetic follows once, elementwise
        torch.randn(min(15, n - i), generator=generator, out=z[i:i + 15])
    table = torch.tensor([d[2:] for d in _COMP_DRAWS + _LIM_DRAWS], dtype=torch.float32)[torch.from_numpy(par)]
    std, mean, high, low = table[:, 0], table[:, 1], table[:, 2], table[:, 3]
    v = torch.minimum(torch.maximum(torch.clamp(z * std + mean, -1.0, 1.0).abs() * high, low), high).numpy()
    for k, (field, sign, *_rest) in enumerate(_COMP_DRAWS + _LIM_DRAWS):
        sel = par == k
        fx[field][rec[sel]] = sign * v[sel]
    return fx


def draw_fx(config: SynthDrumConfig, rng, generator=None) -> np.ndarray:
    """``BoardChain.get_board`` (``synthetiser.py:81-87``): which plugins, with which parameters - the coins and the
    reverb's uniforms from ``rng`` (the ``random`` stream), the dynamics parameters from torch's generator."""
    fx = np.zeros(1, FX_DTYPE)
    if rng.random() < config.use_reverb_prob:               # _add_reverb :44-61
        room = rng.uniform(0.2, 0.8)
        damping = rng.uniform(0.2, 0.8)
        wet = rng.uniform(0.1, 0.4)
        width = rng.uniform(0.6, 1.0)
        fx["flags"] |= FX_REVERB
        fx["room_size"], fx["damping"], fx["wet_level"], fx["dry_level"], fx["width"] = room, damping, wet, 1 - wet, width
    if rng.random() < config.use_compression_prob:          # _add_compression :63-75
        fx["flags"] |= FX_COMPRESSOR
        fx["comp_threshold_db"] = -normal_draw(0.15, 0.5, 10, 0, generator)
        fx["comp_ratio"] = normal_draw(0.15, 0.5, 10, 1.0, generator)
        fx["comp_attack_ms"] = normal_draw(0.05, 0.1, 1000, 0, generator)
        fx["comp_release_ms"] = normal_draw(0.15, 0.2, 1000, 0, generator)
    if rng.random() < config.use_limiter_prob:              # _add_limiter :77-79
        fx["flags"] |= FX_LIMITER
        fx["lim_threshold_db"] = -normal_draw(0.2, 0.4, 3, 0, generator)
    return fx


def plan_segment(notes, config: SynthDrumConfig, bank: OneShotBank, rng=_random, generator=None) -> SegmentPlan:
    """Plan one ``SynthDrum.__call__``.  ``rng`` is the ``random`` module (default, so a
    seeded reference run and a seeded run here draw the same numbers) or a ``random.Random``;
    ``generator``: the torch generator of the FX chain's normal draws (None = the global one, like the reference)."""
    sr = config.sample_rate
    if len(notes) == 0:  # synthetiser.py:257-258
        return SegmentPlan(int(config.input_sec * sr), SEG_EMPTY, 0.0,
                           np.zeros(0, EVENT_DTYPE), np.zeros(0, np.int32), np.zeros(1, np.int32))
    a = notes_to_array(notes)
    dt = a.dtype.type
    n = a.shape[0]

    # ---- segment length (synthetiser.py:262, 243)
    end = dt(a[:, 1].max()) + dt(0.1)
    if end < dt(config.input_sec):
        wave_length = int(config.input_sec * sr)
    else:
        wave_length = int(end * dt(sr))

    # ---- per-note draws, in the reference's order
    groups = similarity_groups(config.similarity_threshold)
    chosen, rank = {}, {}
    main_id = np.empty(n, np.int32)
    sub_id = np.empty(n, np.int32)
    order_rank = np.empty(n, np.int32)
    alpha = np.empty(n, np.float64)
    gains = np.empty(n, np.float32)

    def choose(pitch: int) -> int:  # synthetiser.py:192-202
        if config.ADTOF_mapping:
            pitch = rng.choice(ADTOF_INVERSE[pitch])
        valid = [g for g in groups if bank.has_group(int(pitch), g)]
        g = rng.choice(valid)
        first, count = bank.group_range(int(pitch), g)
        return first + rng.choice(range(count))

    onset, offset, pitch = a[:, 0], a[:, 1], a[:, 2]
    for i in range(n):
        if not (PITCH_MIN <= pitch[i] <= PITCH_MAX and offset[i] >= onset[i]):
            raise ValueError(f"Invalid note: {a[i]}")  # synthetiser.py:270-271
        if onset[i] < 0:
            raise ValueError(f"Invalid note: {a[i]} (negative onset)")
        inst = int(pitch[i])
        if inst != pitch[i]:
            raise KeyError(inst)  # the reference's float-keyed track dict misses here too
        if inst not in chosen:
            chosen[inst] = (choose(inst), choose(inst))
            rank[inst] = len(rank)
        main_id[i], sub_id[i] = chosen[inst]
        order_rank[i] = rank[inst]
        alpha[i] = rng.uniform(0, config.mixup_range)  # synthetiser.py:217
    for inst in rank:  # gain lookup happens in instrument_mixer, after the loop
        g = instrument_gain(inst, config.ADTOF_mapping)
        gains[pitch == inst] = g
    fx = draw_fx(config, rng, generator) if rng.random() < config.use_fx_prob else None  # synthetiser.py:154

    # ---- vectorised index rules
    vel = a[:, 3]
    vol = velocity_to_volume(vel)
    max_volume = float(velocity_to_volume(np.array([max(dt(0), vel.max())], a.dtype))[0])
    start = (onset * dt(sr)).astype(np.int64)              # float product in dt, C truncation
    mix_len = np.maximum(bank.lengths[main_id], bank.lengths[sub_id]).astype(np.int64)
    length = np.minimum(mix_len, wave_length - start)       # synthetiser.py:232-237
    length = np.maximum(length, 0)

    ev = np.zeros(n, EVENT_DTYPE)
    ev["start"], ev["len"] = start, length
    ev["main_id"], ev["sub_id"] = main_id, sub_id
    ev["ca"] = (1.0 - alpha).astype(np.float32)             # python float (1 - mixup) -> f32 (synthetiser.py:223)
    ev["cb"] = alpha.astype(np.float32)
    ev["gain"] = vol.astype(np.float32) * gains
    perm = np.argsort(order_rank, kind="stable")
    ev, mix_len = ev[perm], mix_len[perm].astype(np.int32)
    sorted_rank = order_rank[perm]
    group_ptr = np.concatenate([[0], np.flatnonzero(np.diff(sorted_rank)) + 1, [n]]).astype(np.int32)
    return SegmentPlan(wave_length, SEG_NORMALISE, max_volume, ev, mix_len, group_ptr, fx)


@dataclass
class RenderPlan:
    """A batch of segment plans flattened for the device (all little-endian, C layout)."""
    n_seg: int
    ld_wav: int                   # row pitch of the (n_seg, ld_wav) waveform matrix, multiple of 4
    tiles_per_seg: int
    segments: np.ndarray          # SEGMENT_DTYPE (n_seg,)
    events: np.ndarray            # EVENT_DTYPE (n_events,)
    mix_len: np.ndarray           # int32 (n_events,)
    group_ptr: np.ndarray         # int32 (n_groups+1,)
    tile_ptr: np.ndarray          # int32 (n_seg*tiles_per_seg+1,)
    tile_events: np.ndarray       # int32 (n_refs,) event ids, ascending inside a tile
    peak_work: np.ndarray         # PEAK_ITEM_DTYPE (n_peak_work,): (group, chunk) items of the peak pass
    wave_lengths: np.ndarray = field(default=None)  # int64 (n_seg,)
    # ---- several collated batches in one plan (set_batches): ragged log-mel rows + render chunks
    batch_ptr: np.ndarray = field(default=None)      # int64 (n_batches+1,): segments of batch b
    batch_samples: np.ndarray = field(default=None)  # int64 (n_batches,): collated width of batch b
    batch_frames: np.ndarray = field(default=None)   # int64 (n_batches,): log-mel frames of batch b
    mel_rows: np.ndarray = field(default=None)       # MEL_ROW_DTYPE (n_seg,)
    mel_total_rows: int = 0
    chunks: np.ndarray = field(default=None)         # CHUNK_DTYPE (n_chunks+1,)
    fx: np.ndarray = field(default=None)             # FX_DTYPE (n_fx,), ascending seg: segments with an FX chain
    sample_rate: int = 0                             # needed by the FX kernels (filter tunings)

    def set_batches(self, sizes: Sequence[int], n_frames, chunk_batches: int = 1) -> "RenderPlan":
        """Mark the plan as ``len(sizes)`` collated batches laid end to end (``sizes[b]`` segments
        each).  Every batch keeps its own width - the longest of *its* segments, as
        ``collate_fn``'s ``pad_sequence`` gives (train_dataset.py:53) - and therefore its own
        frame count ``n_frames(width)`` (model.py:95-97).  One chunk of the render per
        ``chunk_batches`` batches."""
        sizes = np.asarray(sizes, np.int64)
        if sizes.sum() != self.n_seg or (sizes <= 0).any():
            raise ValueError("batch sizes must be positive and add up to the number of segments")
        ptr = np.concatenate([[0], np.cumsum(sizes)])
        width = np.maximum.reduceat(self.wave_lengths, ptr[:-1]) if self.n_seg else np.zeros(0, np.int64)
        frames = np.array([n_frames(int(w)) for w in width], np.int64)
        rows = np.zeros(self.n_seg, MEL_ROW_DTYPE)
        row0 = np.concatenate([[0], np.cumsum(sizes * frames)])
        batch_of = np.repeat(np.arange(len(sizes)), sizes)
        rows["count"] = frames[batch_of]
        rows["flags"] = np.where(self.segments["flags"] == SEG_EMPTY, MEL_ROW_SILENT, 0)
        rows["out_row"] = row0[batch_of] + (np.arange(self.n_seg) - ptr[batch_of]) * frames[batch_of]
        cptr = np.unique(np.concatenate([ptr[::max(1, int(chunk_batches))], ptr[-1:]]))
        chunks = np.zeros(len(cptr), CHUNK_DTYPE)
        chunks["seg"] = cptr
        first_event = np.concatenate([self.segments["first_event"].astype(np.int64), [self.n_events]])
        chunks["event"] = first_event[cptr]
        chunks["peak_work"] = np.searchsorted(self.peak_work["first_event"], chunks["event"], side="left")
        if self.fx is not None and len(self.fx):
            chunks["fx_row"] = np.searchsorted(self.fx["seg"], chunks["seg"], side="left")
        self.batch_ptr, self.batch_samples, self.batch_frames = ptr, width.astype(np.int64), frames
        self.mel_rows, self.mel_total_rows, self.chunks = rows, int(row0[-1]), chunks
        return self

    def split(self, wav, feat):
        """Per-batch views of the outputs of a plan with batches: ``[(wav_b (B, Lmax_b), mel_b (B, T_b, n_mels))]``."""
        out = []
        row0 = 0
        for b in range(len(self.batch_frames)):
            s0, s1 = int(self.batch_ptr[b]), int(self.batch_ptr[b + 1])
            t = int(self.batch_frames[b])
            n = (s1 - s0) * t
            out.append((wav[s0:s1, : int(self.batch_samples[b])], feat[row0: row0 + n].view(s1 - s0, t, -1)))
            row0 += n
        return out

    @property
    def n_events(self) -> int:
        return int(self.events.shape[0])

    @property
    def n_groups(self) -> int:
        return int(self.group_ptr.shape[0] - 1)

    def bank_bytes(self, bank: OneShotBank) -> int:
        """Sum over (segment, distinct one-shot) of 4*min(len_u, L_seg - first_start_u): the
        bank bytes one segment cannot avoid reading."""
        if self.n_events == 0:
            return 0
        ev = self.events
        ids = np.concatenate([ev["main_id"], ev["sub_id"]]).astype(np.int64)
        seg = np.concatenate([ev["seg"], ev["seg"]]).astype(np.int64)
        start = np.concatenate([ev["start"], ev["start"]]).astype(np.int64)
        key = seg * (len(bank) + 1) + ids
        order = np.lexsort((start, key))               # by (segment, one-shot), earliest start first
        k = key[order]
        first = np.ones(len(order), bool)
        first[1:] = k[1:] != k[:-1]
        sel = order[first]
        room = self.segments["len"][seg[sel]].astype(np.int64) - start[sel]
        total = np.minimum(bank.lengths[ids[sel]].astype(np.int64), room).clip(0).sum()
        return 4 * int(total)

    def bytes_alg(self, bank: OneShotBank, n_frames: int, n_mels: int) -> int:
        """Algorithmic bytes of render + log-mel for this batch:
        wav write + mel write + distinct one-shot reads + 32 B per event."""
        return (4 * int(self.segments["len"].sum()) + 4 * n_frames * n_mels * self.n_seg
                + self.bank_bytes(bank) + EVENT_DTYPE.itemsize * self.n_events)


def bucket_tiles(start: np.ndarray, length: np.ndarray, seg: np.ndarray, n_seg: int, tiles_per_seg: int):
    """CSR ``tile -> event ids``.  An event lands in every tile its
    ``[start, start+len)`` touches; ids ascend inside a tile (= track order)."""
    n_tiles = n_seg * tiles_per_seg
    live = length > 0
    first = start // TILE
    count = np.where(live, (start + length - 1) // TILE - first + 1, 0).astype(np.int64)
    ev_id = np.repeat(np.arange(len(start), dtype=np.int64), count)
    within = np.arange(int(count.sum()), dtype=np.int64) - np.repeat(np.cumsum(count) - count, count)
    tile = seg.astype(np.int64)[ev_id] * tiles_per_seg + first.astype(np.int64)[ev_id] + within
    order = np.argsort(tile, kind="stable")
    tile_events = ev_id[order].astype(np.int32)
    tile_ptr = np.zeros(n_tiles + 1, np.int64)
    np.add.at(tile_ptr, tile + 1, 1)
    return np.cumsum(tile_ptr).astype(np.int32), tile_events


def assemble(plans: Sequence[SegmentPlan], bank: OneShotBank, ld_wav: int | None = None) -> RenderPlan:
    n_seg = len(plans)
    max_len = max((p.wave_length for p in plans), default=0)
    if ld_wav is None:
        ld_wav = -(-max_len // 4) * 4
    if ld_wav < max_len or ld_wav % 4:
        raise ValueError("ld_wav must be a multiple of 4 and cover the longest segment")
    tiles_per_seg = -(-ld_wav // TILE) if ld_wav else 0
    segments = np.zeros(n_seg, SEGMENT_DTYPE)
    evs, mls, gps, base = [], [], [np.zeros(1, np.int32)], 0
    for s, p in enumerate(plans):
        segments[s] = (p.wave_length, p.flags, p.max_volume, base)
        e = p.events.copy()
        e["seg"] = s
        evs.append(e)
        mls.append(p.mix_len)
        gps.append(p.group_ptr[1:] + base)
        base += len(e)
    events = np.concatenate(evs) if evs else np.zeros(0, EVENT_DTYPE)
    mix_len = np.concatenate(mls).astype(np.int32) if mls else np.zeros(0, np.int32)
    group_ptr = np.concatenate(gps).astype(np.int32)
    tile_ptr, tile_events = bucket_tiles(events["start"].astype(np.int64), events["len"].astype(np.int64),
                                         events["seg"], n_seg, tiles_per_seg)
    fx_rows = []
    for s, p in enumerate(plans):
        if p.fx is not None:
            r = p.fx.copy()
            r["seg"] = s
            fx_rows.append(r)
    plan = RenderPlan(n_seg, ld_wav, tiles_per_seg, segments, events, mix_len, group_ptr, tile_ptr,
                      tile_events, peak_work_items(events, mix_len, group_ptr, bank),
                      np.array([p.wave_length for p in plans], np.int64))
    plan.fx = np.concatenate(fx_rows) if fx_rows else None
    return plan


PEAK_NOTES = 8  # notes of a group per peak work item (ADTFE_PEAK_NOTES): one warp of the peak pass each


def peak_work_items(events: np.ndarray, mix_len: np.ndarray, group_ptr: np.ndarray, bank: OneShotBank) -> np.ndarray:
    """One record per PEAK_NOTES notes of a group (the notes of one instrument in one segment), bank lookups resolved;
    ``chunk`` stays 0."""
    n_groups = len(group_ptr) - 1
    if n_groups <= 0:
        return np.zeros(0, PEAK_ITEM_DTYPE)
    size = np.diff(group_ptr).astype(np.int64)
    parts = np.maximum(1, -(-size // PEAK_NOTES))                       # items per group
    group = np.repeat(np.arange(n_groups, dtype=np.int64), parts)
    k = np.arange(len(group), dtype=np.int64) - np.repeat(np.cumsum(parts) - parts, parts)   # item index inside its group
    out = np.zeros(len(group), PEAK_ITEM_DTYPE)
    head = group_ptr[:-1].astype(np.int64)[group]
    main, sub = events["main_id"][head], events["sub_id"][head]
    out["a_off"], out["b_off"] = bank.offsets[main], bank.offsets[sub]
    out["la"], out["lb"] = bank.lengths[main], bank.lengths[sub]
    out["mix_len"] = mix_len[head]
    out["first_event"] = head + PEAK_NOTES * k
    out["n_events"] = np.minimum(PEAK_NOTES, size[group] - PEAK_NOTES * k)
    return out


def plan_batch(batch_notes: Sequence, config: SynthDrumConfig, bank: OneShotBank, rng=_random,
               ld_wav: int | None = None, generator=None) -> RenderPlan:
    """Plan ``len(batch_notes)`` independent ``SynthDrum.__call__``s, in order (the RNG
    streams advance exactly as that many reference calls would advance them)."""
    plan = assemble([plan_segment(n, config, bank, rng, generator) for n in batch_notes], bank, ld_wav)
    plan.sample_rate = int(config.sample_rate)
    return plan
