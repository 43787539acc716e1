"""``SynthDrum`` - drop-in for reference ``modules/synthetiser.py:159-292``.

``SynthDrum(config)(notes, eval_rendering=False) -> FloatTensor[L]`` keeps the
reference's signature, RNG stream, error behaviour and CPU return value, but the
audio is rendered on a B200: the host planner (``planner.py``) turns notes into
event records and tile buckets, ``libadtfe``'s mixer kernels do the overlap-add
against a one-shot bank that stays resident in HBM.  ``render_batch`` is the
entry the training loop should call from the main process (one launch per batch,
output left on the device for ``ComputeMelSpectrogram``); calling ``__call__``
from forked DataLoader workers, as the reference does
(``data_modules/train_dataset.py:228``), cannot work with CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
import random as _random
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .bank import OneShotBank
from .config import SynthDrumConfig
from .planner import RenderPlan, plan_batch


class DeviceBank:
    """The packed bank uploaded once to one GPU (replaces h5py.File per note, synthetiser.py:273)."""

    def __init__(self, bank: OneShotBank, device: torch.device):
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.adtfe_bank_create(bank.pcm.ctypes.data, bank.pcm.size, bank.offsets.ctypes.data,
                                         bank.lengths.ctypes.data, len(bank), device.index, C.byref(h)),
                   "adtfe_bank_create")
        self.handle, self.lib, self.device = h, lib, device

    def __del__(self):
        try:
            if self.handle:
                self.lib.adtfe_bank_destroy(self.handle)
        except Exception:
            pass


class PlanBuffers:
    """Pinned host blob + device blob for one planned batch (grown on demand, reused)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.host = torch.empty(0, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(0, dtype=torch.uint8, device=device)
        self.workspace = torch.empty(0, dtype=torch.uint8, device=device)
        self.nbytes = 0
        self._uploaded: Optional[torch.cuda.Event] = None   # recorded behind the last H2D copy of `host`

    def fence(self) -> None:
        """Block until the last ``upload`` has read the pinned blob: the copy is asynchronous, so the host must not
        rewrite ``host`` (next ``pack`` / ``reserve`` / in-place native packing) while it is still in flight."""
        if self._uploaded is not None:
            self._uploaded.synchronize()
            self._uploaded = None

    def pack(self, plan: RenderPlan) -> _lib.Plan:
        lib = _lib.load()
        self.fence()
        batched = plan.mel_rows is not None
        self._chunks = np.ascontiguousarray(plan.chunks) if batched else None  # host memory the library reads
        n_fx = 0 if plan.fx is None else len(plan.fx)
        shape = _lib.Plan(None, None, None, None, None, plan.n_events, plan.n_seg, plan.tiles_per_seg,
                          len(plan.peak_work), plan.ld_wav, None,
                          plan.mel_total_rows if batched else 0,
                          int(plan.batch_frames.max()) if batched and len(plan.batch_frames) else 0,
                          len(self._chunks) - 1 if batched else 0,
                          self._chunks.ctypes.data if batched else None, len(plan.tile_events),
                          n_fx, None, int(plan.sample_rate))
        off = (C.c_size_t * 7)()
        fixed = C.c_size_t()
        _lib.check(lib.adtfe_plan_blob_layout(C.byref(shape), C.byref(off), C.byref(fixed)), "adtfe_plan_blob_layout")
        total = (fixed.value + 4 * len(plan.tile_events) + 15) & ~15
        if self.host.numel() < total:
            cap = max(total, 2 * self.host.numel(), 1 << 16)
            self.host = torch.empty(cap, dtype=torch.uint8).pin_memory()
            self.dev = torch.empty(cap, dtype=torch.uint8, device=self.device)
        h = self.host.numpy()
        rows = plan.mel_rows if batched else np.zeros(0, np.uint8)
        fx = plan.fx if n_fx else np.zeros(0, np.uint8)
        for o, arr in zip(off, (plan.events, plan.segments, plan.tile_ptr, plan.peak_work, rows, fx, plan.tile_events)):
            raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            h[o: o + raw.size] = raw
        self.nbytes = total
        need = lib.adtfe_render_workspace_bytes(plan.n_events, plan.n_seg, plan.tiles_per_seg, len(plan.tile_events))
        if self.workspace.numel() < need:
            self.workspace = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=self.device)
        self.offsets = [int(o) for o in off]
        self.shape = shape
        return shape

    def reserve(self, nbytes: int) -> None:
        """Pinned and device blobs of at least ``nbytes`` (contents are not kept)."""
        self.fence()
        if self.host.numel() < nbytes:
            cap = max(int(nbytes), 2 * self.host.numel(), 1 << 16)
            self.host = torch.empty(cap, dtype=torch.uint8).pin_memory()
            with torch.cuda.device(self.device):
                self.dev = torch.empty(cap, dtype=torch.uint8, device=self.device)

    def adopt(self, shape: _lib.Plan, nbytes: int, chunks: np.ndarray) -> None:
        """Take over a blob written in place by ``adtfe_planner_pack_batches`` (``shape`` as it returned it;
        ``chunks``: the int32 array behind ``shape.chunks_host``, kept alive here)."""
        lib = _lib.load()
        self._chunks = chunks
        off = (C.c_size_t * 7)()
        fixed = C.c_size_t()
        _lib.check(lib.adtfe_plan_blob_layout(C.byref(shape), C.byref(off), C.byref(fixed)), "adtfe_plan_blob_layout")
        self.nbytes = int(nbytes)
        need = lib.adtfe_render_workspace_bytes(shape.n_events, shape.n_seg, shape.tiles_per_seg, shape.n_tile_events)
        if self.workspace.numel() < need:
            with torch.cuda.device(self.device):
                self.workspace = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=self.device)
        self.offsets = [int(o) for o in off]
        self.shape = shape

    def upload(self, shape: _lib.Plan) -> _lib.Plan:
        """H2D of the packed blob on the current stream; returns the plan with device pointers."""
        self.dev[: self.nbytes].copy_(self.host[: self.nbytes], non_blocking=True)
        with torch.cuda.device(self.device):
            self._uploaded = torch.cuda.Event()
            self._uploaded.record()
        base = self.dev.data_ptr()
        o = self.offsets
        return _lib.Plan(base + o[0], base + o[1], base + o[2], base + o[6], base + o[3],
                         shape.n_events, shape.n_seg, shape.tiles_per_seg, shape.n_peak_work, shape.ld_wav,
                         base + o[4] if shape.mel_total_rows > 0 else None, shape.mel_total_rows,
                         shape.mel_max_count, shape.n_chunks, shape.chunks_host, shape.n_tile_events,
                         shape.n_fx, base + o[5] if shape.n_fx > 0 else None, shape.sample_rate)


class SynthDrum:
    def __init__(self, config: SynthDrumConfig, bank: Optional[OneShotBank] = None,
                 device: Optional[torch.device] = None):
        self.config = config
        self.sample_rate = config.sample_rate
        self.oneshot_path = f"{config.oneshot_path}@{self.sample_rate}.hdf5"  # synthetiser.py:163
        self.similarity_threshold = config.similarity_threshold
        self.ADTOF_mapping = config.ADTOF_mapping
        self._bank = bank
        self._device = torch.device(device) if device is not None else None
        self._device_bank: Optional[DeviceBank] = None
        self._buffers: Optional[PlanBuffers] = None
        self._native_planner = None
        #: who runs the FX chain of segments whose FX coin hit (``use_fx_prob``): "gpu" - the kernels of csrc/fx.cu
        #: (the published JUCE algorithms pedalboard wraps, everything stays on the device) - or "pedalboard": the
        #: reference's own library on the host, for users who have it installed and want its exact DSP
        self.fx_backend = "gpu"

    # ------------------------------------------------------------------ bank
    @property
    def bank(self) -> OneShotBank:
        """Packed bank: the one given, else ``<oneshot_path>@<sr>.npz``, else the reference's
        ``.hdf5`` converted on the fly (needs h5py).  Opened lazily, like the reference."""
        if self._bank is None:
            packed = f"{self.config.oneshot_path}@{self.sample_rate}.npz"
            if os.path.exists(packed):
                self._bank = OneShotBank.load(packed)
            elif os.path.exists(self.oneshot_path):
                self._bank = OneShotBank.from_hdf5(self.oneshot_path)
            else:
                raise FileNotFoundError(f"no one-shot bank at {packed} or {self.oneshot_path}")
        return self._bank

    @property
    def device(self) -> torch.device:
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("SynthDrum needs a CUDA device (sm_100a); there is no CPU path")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    def device_bank(self) -> DeviceBank:
        if self._device_bank is None:
            self._device_bank = DeviceBank(self.bank, self.device)
        return self._device_bank

    def buffers(self) -> PlanBuffers:
        if self._buffers is None:
            self._buffers = PlanBuffers(self.device)
        return self._buffers

    # ------------------------------------------------------------------ plan
    def plan(self, batch_notes: Sequence, rng=_random, ld_wav: Optional[int] = None, native: bool = True,
             generator=None) -> RenderPlan:
        """Plan a batch on the host.  ``native`` uses the C++ planner (same RNG stream, same
        indices); it falls back to ``planner.plan_batch`` by itself for float64 note arrays.
        ``generator``: torch generator for the FX chain's normal draws (None: the global one, like the reference)."""
        if not native:
            return plan_batch(batch_notes, self.config, self.bank, rng, ld_wav, generator)
        if self._native_planner is None:
            from .native_planner import NativePlanner
            self._native_planner = NativePlanner(self.config, self.bank)
        return self._native_planner.plan_batch(batch_notes, rng, ld_wav, generator)

    def plan_batches(self, batches: Sequence[Sequence], n_frames, rng=_random, chunk_batches: int = 1) -> RenderPlan:
        """Plan several batches as ONE device plan (one H2D copy, one launch per kernel): the segments
        of all batches in order - the RNG stream advances exactly as planning them one by one would -
        with per-batch collated widths and frame counts (``RenderPlan.set_batches``).
        ``n_frames`` is ``ComputeMelSpectrogram.n_frames``."""
        flat = [notes for b in batches for notes in b]
        return self.plan(flat, rng).set_batches([len(b) for b in batches], n_frames, chunk_batches)

    # ---------------------------------------------------------------- render
    def _render_through_pedalboard(self, plan: RenderPlan, out: Optional[torch.Tensor]) -> torch.Tensor:
        """``fx_backend = "pedalboard"``: the rows with an FX record are rendered RAW (ADTFE_SEG_RAW: no normalisation,
        no GPU FX), brought to the host, sent through the plugins the reference builds (``_add_fx``,
        synthetiser.py:121-137) with the parameters the plan drew, normalised like ``_normalize_audio(wav) * max_volume``
        (:142-144, 156) and written back.  Slow (one round trip per batch) but it is the reference's own DSP."""
        import copy
        import pedalboard
        raw = copy.copy(plan)
        raw.segments = plan.segments.copy()
        raw.segments["flags"][plan.fx["seg"]] = 2      # ADTFE_SEG_RAW
        raw.fx = None
        out = self.render_plan(raw, out)
        sr = int(self.config.sample_rate)
        for r in plan.fx:
            seg, n = int(r["seg"]), int(plan.segments["len"][r["seg"]])
            if plan.segments["flags"][seg] == 0:
                continue
            x = out[seg, :n].cpu().numpy()
            board = pedalboard.Pedalboard([])
            if r["flags"] & 1:
                board.append(pedalboard.Reverb(room_size=float(r["room_size"]), damping=float(r["damping"]),
                                               wet_level=float(r["wet_level"]), dry_level=float(r["dry_level"]),
                                               width=float(r["width"]), freeze_mode=0.0))
            if r["flags"] & 2:
                board.append(pedalboard.Compressor(threshold_db=float(r["comp_threshold_db"]), ratio=float(r["comp_ratio"]),
                                                   attack_ms=float(r["comp_attack_ms"]),
                                                   release_ms=float(r["comp_release_ms"])))
            if r["flags"] & 4:
                board.append(pedalboard.Limiter(threshold_db=float(r["lim_threshold_db"])))
            y = torch.from_numpy(np.asarray(board(x[None, :].T.astype(np.float32), sample_rate=sr)).T[0].copy())
            y = y / y.abs().max() * float(plan.segments["max_volume"][seg])
            out[seg, :n] = y.to(out.device)
        return out

    def render_plan(self, plan: RenderPlan, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Enqueue one planned batch on the current stream -> (n_seg, ld_wav) float32 on the device."""
        if plan.fx is not None and len(plan.fx) and self.fx_backend == "pedalboard":
            return self._render_through_pedalboard(plan, out)
        dev = self.device
        bank = self.device_bank()
        with torch.cuda.device(dev):
            buf = self.buffers()
            shape = buf.pack(plan)
            if out is None:
                out = torch.empty((plan.n_seg, plan.ld_wav), dtype=torch.float32, device=dev)
            elif out.shape != (plan.n_seg, plan.ld_wav) or not out.is_contiguous() or out.dtype != torch.float32:
                raise ValueError("out must be a contiguous float32 (n_seg, ld_wav) tensor")
            if plan.n_seg and plan.ld_wav:
                dplan = buf.upload(shape)
                stream = torch.cuda.current_stream(dev).cuda_stream
                _lib.check(bank.lib.adtfe_render(bank.handle, C.byref(dplan), out.data_ptr(),
                                                 buf.workspace.data_ptr(), buf.workspace.numel(), stream),
                           "adtfe_render")
        return out

    def render_batch(self, batch_notes: Sequence, rng=_random) -> Tuple[torch.Tensor, torch.Tensor]:
        """Render ``len(batch_notes)`` segments in one launch.  Returns the collated
        ``(B, Lmax)`` waveform matrix on the device - what ``collate_fn``'s ``pad_sequence``
        (train_dataset.py:53) would build - and the per-segment lengths."""
        plan = self.plan(batch_notes, rng)
        wav = self.render_plan(plan)
        l_max = int(plan.wave_lengths.max()) if plan.n_seg else 0
        return wav[:, :l_max], torch.from_numpy(plan.wave_lengths.copy())

    def __call__(self, notes, eval_rendering=False):
        if eval_rendering:
            # the reference reads an attribute it never defines (synthetiser.py:287-288)
            raise NotImplementedError("eval_rendering is not implemented (undefined in the reference as well)")
        if len(notes) == 0:  # synthetiser.py:257-258, no device work needed
            return torch.zeros(int(self.config.input_sec * self.config.sample_rate))
        plan = self.plan([notes])
        wav = self.render_plan(plan)
        return wav[0, : int(plan.wave_lengths[0])].cpu()

    def close(self) -> None:
        self._device_bank = None
        self._buffers = None
