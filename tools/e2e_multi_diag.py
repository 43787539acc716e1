#!/usr/bin/env python3
"""Where the end-to-end time goes when several ranks share one host: per rank, all ranks at the same time (barriers):
planning alone (native plan + pack, 4 threads), device work + D2H alone (plans packed beforehand), D2H alone, and the
whole HostPipeline.    python -m torch.distributed.run --nproc-per-node N tools/e2e_multi_diag.py"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from concurrent.futures import ThreadPoolExecutor
from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum, HostPipeline
from adt_str_b200.config import setting_1
from adt_str_b200.native_planner import NativePlanner
from adt_str_b200.synthetic import make_bank, make_segments

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group(os.environ.get("DIAG_BACKEND", "gloo"))
NB = 256
bank = make_bank(10000, 24000, seed=0)
segs = make_segments(NB * 64, seed=1 + rank)
batches = [segs[i * 64:(i + 1) * 64] for i in range(NB)]
groups = [batches[i:i + 8] for i in range(0, NB, 8)]
synth = SynthDrum(setting_1(), bank=bank, device=dev); mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
fe = FrontEnd(synth, mel)
out = {}

def sync():
    torch.cuda.synchronize(dev); dist.barrier()

# 1. planning alone
for workers in (1, 2, 4, 8):
    planners = {}
    def plan(g):
        import threading
        tid = threading.get_ident()
        if tid not in planners:
            planners[tid] = (NativePlanner(synth.config, bank), np.array(random.Random(tid).getstate()[1], np.uint32),
                             np.zeros(8 << 20, np.uint8))
        pl, mt, blob = planners[tid]
        pl.plan_group(g, mt)
        return pl.pack_group([len(b) for b in g], 240, 5, 4, blob.ctypes.data, blob.size)[0]
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(plan, groups[:workers]))
        sync(); t0 = time.perf_counter(); res = list(ex.map(plan, groups)); dt = time.perf_counter() - t0
    assert all(r == 0 for r in res)
    out[f"plan_{workers}thr_ms"] = round(dt * 1e3, 1)
# 2. D2H alone
x = torch.empty((NB * 64 * 250, 128), device=dev); h = torch.empty(x.shape).pin_memory()
for _ in range(2):
    sync(); t0 = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(dev); dt = time.perf_counter() - t0
out["d2h_ms"] = round(dt * 1e3, 1); out["d2h_gbs"] = round(x.numel() * 4 / dt / 1e9, 1)
del x, h
# 3. the whole pipeline
for workers, sets in ((4, 4), (8, 6), (2, 4)):
    pipe = HostPipeline(fe, workers=workers, n_sets=sets, seed=3)
    for rep in range(3):
        sync(); t0 = time.perf_counter()
        infl = []
        for r in pipe.run(groups):
            infl.append(r)
            if len(infl) > 2: infl.pop(0).wait().release()
        for r in infl: r.wait().release()
        torch.cuda.synchronize(dev); dt = time.perf_counter() - t0
    pipe.close()
    out[f"pipeline_{workers}w_{sets}s_ms"] = round(dt * 1e3, 1)
# 4. device side alone: one group's packed plan, enqueued 32 times (kernels + D2H, no planning)
pipe = HostPipeline(fe, workers=1, n_sets=2, seed=3)
res = next(iter(pipe.run(groups[:1]))); res.wait()
s = pipe._sets[0]
wav = s.wav[: res.plan.n_seg * res.plan.ld_wav].view(res.plan.n_seg, res.plan.ld_wav)
for rep in range(2):
    sync(); t0 = time.perf_counter()
    for _ in range(len(groups)):
        fe.run_plan_host(res.plan, s.host, None, buffers=s.buf, wav=wav, feat=s.feat, packed=True, copy_stream=pipe._copy_stream)
    torch.cuda.synchronize(dev); dt = time.perf_counter() - t0
out["device_and_d2h_only_ms"] = round(dt * 1e3, 1)
allres = [None] * world
dist.all_gather_object(allres, out)
if rank == 0:
    print("cpus", len(os.sched_getaffinity(0)), "world", world, "(ms per 256-batch step, per rank)")
    for k in out:
        print(f"{k:28s}", [r[k] for r in allres])
dist.destroy_process_group()
