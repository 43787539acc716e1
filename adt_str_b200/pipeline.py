"""Host-side pipeline: note lists in host memory -> log-mel in pinned host memory.

The reference renders inside forked DataLoader workers, each with its own ``random``
state (``train.py:235-237``, ``data_modules/train_dataset.py:213-229``).  Here the
workers are *threads* that only plan - the C++ planner runs without the GIL - each with
its own planner handle; every group of batches gets its own ``random`` stream derived from
(seed, rank, group index), so a seed reproduces the same draws whatever thread plans a group; the main thread enqueues one
``adtfe_frontend_host`` call per group of batches (plan blob H2D -> render -> log-mel on the
current stream, the D2H on a copy stream so that it overlaps the next group's kernels).
``n_sets`` rotating buffer sets (pinned blob, device blob, workspace, device outputs, pinned
log-mel) let planning of the next groups, the GPU work of the current one and the copies of
the previous one overlap.

    pipe = HostPipeline(frontend, workers=8, seed=0)
    for result in pipe.run(groups):          # groups: iterable of [batch, ...]; batch: [notes, ...]
        for wav_len, logmel in result.batches():   # logmel: pinned (B, T, n_mels) float32 view
            ...
        result.release()                     # the set may be reused from here on
"""
from __future__ import annotations

import random
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from .frontend import FrontEnd
from .planner import RenderPlan
from .synthetiser import PlanBuffers


class _Set:
    """One rotating buffer set."""

    def __init__(self, device: torch.device):
        self.buf = PlanBuffers(device)
        self.wav: Optional[torch.Tensor] = None
        self.feat: Optional[torch.Tensor] = None
        self.host: Optional[torch.Tensor] = None
        self.done = torch.cuda.Event()
        self.free = threading.Event()
        self.free.set()


class GroupResult:
    """Outputs of one group of batches; valid until ``release()``."""

    def __init__(self, plan: RenderPlan, s: _Set, n_mels: int):
        self.plan, self._set, self._n_mels = plan, s, n_mels
        self.h2d_bytes = s.buf.nbytes
        self.d2h_bytes = plan.mel_total_rows * n_mels * 4

    def wait(self) -> "GroupResult":
        self._set.done.synchronize()
        return self

    def batches(self):
        """``[(lengths (B,), logmel (B, T, n_mels) pinned host view)]`` - call after ``wait()``."""
        p, out, row0 = self.plan, [], 0
        for b in range(len(p.batch_frames)):
            s0, s1, t = int(p.batch_ptr[b]), int(p.batch_ptr[b + 1]), int(p.batch_frames[b])
            n = (s1 - s0) * t
            out.append((p.wave_lengths[s0:s1], self._set.host[row0:row0 + n].view(s1 - s0, t, self._n_mels)))
            row0 += n
        return out

    def release(self) -> None:
        self._set.free.set()


class HostPipeline:
    def __init__(self, frontend: FrontEnd, workers: int = 4, n_sets: int = 4, seed: int = 0, chunk_batches: int = 8,
                 rank: Optional[int] = None, compute_streams: Optional[int] = None):
        """``compute_streams``: streams the groups' kernels rotate over.  1 = the stream that is current when ``run``
        enqueues (the default without FX).  With the FX chain on (``use_fx_prob > 0``) the default is one per buffer
        set: a group's FX rows are recursions over the row's samples that finish ~2.5 ms after its last mixer launch
        whatever the group's size (a warp or two per SM), and on one stream every group would wait for the previous
        group's chain; on several, the chains of the groups in flight run beside each other and under the next renders."""
        self.fe = frontend
        if rank is None:   # the process's rank under torch.distributed / torchrun, else 0
            import os
            import torch.distributed as dist
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else int(os.environ.get("RANK", "0"))
        self.rank = int(rank)
        self.workers = max(1, int(workers))
        self.n_sets = max(2, int(n_sets))
        self.chunk_batches = chunk_batches
        self.device = frontend.synth.device
        self.n_mels = frontend.mel.compute_spec.n_mels
        self._sets = [_Set(self.device) for _ in range(self.n_sets)]
        self._local = threading.local()
        self._seed = seed
        self._groups_seen = 0   # group index across run() calls: part of every group's RNG seed
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="adtfe-plan")
        self._copy_stream = torch.cuda.Stream(self.device)   # D2H of group g overlaps the kernels of group g+1
        if compute_streams is None:
            compute_streams = self.n_sets if frontend.synth.config.use_fx_prob > 0 else 1
        self._compute = [torch.cuda.Stream(self.device) for _ in range(compute_streams)] if compute_streams > 1 else []
        self._enqueued = 0

    # ---- worker side: plan + pack into the set's pinned blob (no CUDA calls besides the event wait)
    def _worker_state(self):
        st = getattr(self._local, "st", None)
        if st is None:
            from .native_planner import NativePlanner
            st = self._local.st = dict(rng=random.Random(0),
                                       planner=NativePlanner(self.fe.synth.config, self.fe.synth.bank),
                                       mt=np.zeros(625, np.uint32))
        return st

    def _plan(self, group: Sequence[Sequence], s: _Set, index: int):
        st = self._worker_state()
        # One RNG stream per GROUP, derived from (seed, rank, group index): which thread plans a group is up to the
        # executor, the draws are not - a seed reproduces the same timbres and mixups run after run, and ranks that
        # share a seed still render different augmentations (DataLoader workers: seed + worker id, per rank).
        group_seed = (self._seed * 1_000_003 + self.rank) * 1_000_003 + index
        st["rng"].seed(group_seed)
        st["mt"][:] = np.array(st["rng"].getstate()[1], np.uint32)   # the same stream, kept native
        gen = None
        if self.fe.synth.config.use_fx_prob > 0:   # the FX chain's normal draws: torch's stream, one per group as well
            gen = st.setdefault("gen", torch.Generator())
            gen.manual_seed(group_seed % (1 << 63))

        def acquire():
            s.free.wait()          # the consumer released the set ...
            s.free.clear()
            s.done.synchronize()   # ... and the GPU has left it
            return s.buf

        # fast path: plan, wait for the buffer set, pack - two library calls, no interpreter work per note or record
        spec = self.fe.mel.compute_spec
        plan = st["planner"].plan_group_into(group, st["mt"], acquire, spec.hop_length, self.fe.mel.window_pad_idxs,
                                             self.chunk_batches, gen)
        if plan is not None:
            return plan, s.buf.shape
        # general path (python lists, float64 notes ...): the worker's random.Random continues the native stream
        version, _, gauss = st["rng"].getstate()
        st["rng"].setstate((version, tuple(int(v) for v in st["mt"]), gauss))
        flat = [notes for b in group for notes in b]
        plan = st["planner"].plan_batch(flat, st["rng"], None, gen).set_batches([len(b) for b in group],
                                                                                 self.fe.mel.n_frames, self.chunk_batches)
        st["mt"][:] = np.array(st["rng"].getstate()[1], np.uint32)
        shape = acquire().pack(plan)
        return plan, shape

    # ---- main thread: enqueue in order
    def run(self, groups: Iterable[Sequence[Sequence]]) -> Iterator[GroupResult]:
        """Yields a ``GroupResult`` per group, in order, as soon as its work is *enqueued*;
        ``result.wait()`` blocks until the log-mel is in host memory."""
        pending = deque()
        it = iter(groups)
        k = 0
        exhausted = False
        while True:
            while not exhausted and len(pending) < self.n_sets:
                try:
                    g = next(it)
                except StopIteration:
                    exhausted = True
                    break
                s = self._sets[k % self.n_sets]
                pending.append((self._pool.submit(self._plan, g, s, self._groups_seen), s))
                self._groups_seen += 1
                k += 1
            if not pending:
                return
            fut, s = pending.popleft()
            plan, _shape = fut.result()
            self._enqueue(plan, s)
            yield GroupResult(plan, s, self.n_mels)

    def _enqueue(self, plan: RenderPlan, s: _Set) -> None:
        dev = self.device
        if s.wav is None or s.wav.numel() < plan.n_seg * plan.ld_wav:   # flat, grown on demand, viewed per plan
            s.wav = torch.empty(int(plan.n_seg * plan.ld_wav * 1.02) + 64, dtype=torch.float32, device=dev)
        if s.feat is None or s.feat.shape[0] < plan.mel_total_rows:
            rows = int(plan.mel_total_rows * 1.02) + 64
            s.feat = torch.empty((rows, self.n_mels), dtype=torch.float32, device=dev)
            s.host = torch.empty((rows, self.n_mels), dtype=torch.float32).pin_memory()
        wav = s.wav[: plan.n_seg * plan.ld_wav].view(plan.n_seg, plan.ld_wav)
        if self._compute:   # the set's buffers are free (the worker waited for its `done`): any stream may take them
            st = self._compute[self._enqueued % len(self._compute)]
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                self.fe.run_plan_host(plan, s.host, None, buffers=s.buf, wav=wav, feat=s.feat, packed=True,
                                      copy_stream=self._copy_stream)
        else:
            self.fe.run_plan_host(plan, s.host, None, buffers=s.buf, wav=wav, feat=s.feat, packed=True,
                                  copy_stream=self._copy_stream)
        self._enqueued += 1
        s.done.record(self._copy_stream)

    def close(self) -> None:
        self._pool.shutdown(wait=True)
