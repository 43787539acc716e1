#!/usr/bin/env python3
"""Where the end-to-end time goes: planner threads alone, D2H alone, pipeline with and without the D2H."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum, HostPipeline
from adt_str_b200.config import setting_1
from adt_str_b200.native_planner import NativePlanner
from adt_str_b200.synthetic import make_bank, make_segments

bank = make_bank(10000, 24000, seed=0)
segs = make_segments(128 * 64, seed=1)
batches = [segs[i * 64:(i + 1) * 64] for i in range(128)]
dev = torch.device("cuda", 0)
synth = SynthDrum(setting_1(), bank=bank, device=dev); mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
fe = FrontEnd(synth, mel)
audio_s = sum(2.56 for _ in segs)
print("cpus", len(os.sched_getaffinity(0)))
for group in (8, 16, 32):
    groups = [batches[i:i + group] for i in range(0, len(batches), group)]
    for workers in (1, 4, 8, 16):
        planners = {}
        def plan(g):
            import threading
            tid = threading.get_ident()
            if tid not in planners:
                planners[tid] = (NativePlanner(synth.config, bank), random.Random(tid))
            pl, rng = planners[tid]
            flat = [n for b in g for n in b]
            return pl.plan_batch(flat, rng).set_batches([len(b) for b in g], mel.n_frames, 4)
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(plan, groups[:workers]))
            t0 = time.perf_counter(); res = list(ex.map(plan, groups)); dt = time.perf_counter() - t0
        print(f"plan only: group {group:3d} workers {workers:2d}: {dt*1e3:7.1f} ms for 128 batches -> {audio_s/dt/1e3:8.0f} k audio-s/s")
# D2H alone
x = torch.empty((128 * 64 * 250, 128), device=dev); h = torch.empty(x.shape).pin_memory()
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H {x.numel()*4/1e9:.2f} GB in {dt*1e3:.1f} ms = {x.numel()*4/dt/1e9:.1f} GB/s")
for workers in (4, 8, 16):
    for group in (8, 16, 32):
        groups = [batches[i:i + group] for i in range(0, len(batches), group)]
        pipe = HostPipeline(fe, workers=workers, n_sets=4, seed=3)
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            infl = []
            for r in pipe.run(groups):
                infl.append(r)
                if len(infl) > 2: infl.pop(0).wait().release()
            for r in infl: r.wait().release()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        pipe.close()
        print(f"pipeline: workers {workers:2d} group {group:3d}: {dt*1e3:7.1f} ms -> {audio_s/dt/1e3:8.0f} k audio-s/s")
