#!/usr/bin/env python3
"""Benchmark of the render + log-mel front end (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                      # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "training-shape batch"): setting-1 shapes - 64 segments
x 2.56 s @ 24 kHz per batch (configs/train/setting-1.yaml:6-11,29-33), random MIDI event
streams and a synthetic 10 000 one-shot bank.  One batch is only ~100 us of GPU
work, so a *step* is ``--batches-per-step`` (256) distinct batches run back to back, each
planned into ONE device plan and run by one ``adtfe_render_logmel`` call - the render chunk by chunk
(one chunk per batch, over the library's internal streams), the log-mel kernel as a single persistent
launch with per-batch frame counts (SURVEY §8d "config 2": persistent launch over >=256 batches).  Per
rank the work is fixed (weak scaling); ranks share nothing but the final statistics.

value  : audio-seconds rendered+featurised per second, plans already resident in HBM.
e2e    : the same through the public API (FrontEnd.__call__ semantics): host planning of the
         note lists, pinned-host plan blob -> H2D, kernels, log-mel -> pinned host (D2H),
         all inside the timed region.
roofline: dominant kernel (fused log-mel) - algorithmic bytes (4*L read + 4*T*n_mels write per
         segment) / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs; the
         whole path's figure (bytes_alg of SURVEY §8d / step time) is reported beside it.
cpu_baseline: the CPU oracle port (oracle/, same arithmetic and library calls as the reference)
         on the box's host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "render+log-mel audio-seconds/sec"
UNIT = "audio-s/s"
SR, INPUT_SEC, BATCH = 24000, 2.56, 64


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batches-per-step", type=int, default=256)
    p.add_argument("--bank-size", type=int, default=10_000)
    p.add_argument("--e2e-steps", type=int, default=3)
    p.add_argument("--cpu-segments", type=int, default=0, help="cpu_baseline sample size (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-long-form", action="store_true", help="skip the long-form inference front side measurement")
    p.add_argument("--e2e-group", type=int, default=8, help="batches per end-to-end plan")
    p.add_argument("--e2e-workers", type=int, default=0,
                   help="planner threads of the end-to-end pipeline (0 = host cores / ranks, at most 8)")
    p.add_argument("--e2e-sets", type=int, default=4, help="rotating buffer sets of the end-to-end pipeline")
    p.add_argument("--chunk-batches", type=int, default=64, help="batches rendered together as one chunk")
    p.add_argument("--no-library-baseline", action="store_true",
                   help="skip the torchaudio (cuFFT + cuBLAS) log-mel on the same GPU, the comparator of BASELINE.md")
    p.add_argument("--no-traffic", action="store_true", help="skip the in-run ncu measurement of roofline.traffic")
    p.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    return p.parse_args()


def workload_name(args):
    return (f"setting-1 training-shape batches: {BATCH} segments x {INPUT_SEC} s @ {SR} Hz, "
            f"{args.batches_per_step} batches per step, {args.bank_size} one-shot synthetic bank")


# --------------------------------------------------------------------------- CPU arm (oracle port)
_CPU_STATE = {}


def _cpu_task(task):
    """One worker task: render `segs` with the oracle, collate, log-mel through torchaudio."""
    import torch
    from oracle import mel_oracle, synth_oracle
    torch.set_num_threads(1)
    seed, lo, hi = task
    st = _CPU_STATE
    rng = random.Random(seed)
    wavs = [synth_oracle.render(s, st["cfg"], st["nested"], rng=rng) for s in st["segs"][lo:hi]]
    batch = synth_oracle.collate(wavs)
    mel = mel_oracle.logmel_torchaudio(batch, SR, 2048, 0.01, 128)
    return float(sum(len(w) for w in wavs)) / SR, float(mel.sum())


def cpu_pool(bank, segs, cfg):
    """Fork a pool of one-thread workers (mirrors DataLoader workers, train.py:235-237).
    Must be called before CUDA is initialised in this process."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = max(1, min(cores, 64))
    _CPU_STATE.update(cfg=cfg, nested=bank.to_nested(), segs=segs)
    ctx = mp.get_context("fork")
    return ctx.Pool(cores), cores


def cpu_run(pool, cores, n_segments, chunk=8, seed=0):
    tasks = [(seed + i, lo, min(lo + chunk, n_segments)) for i, lo in enumerate(range(0, n_segments, chunk))]
    t0 = time.perf_counter()
    res = pool.map(_cpu_task, tasks, chunksize=max(1, len(tasks) // (cores * 4)))
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt


# --------------------------------------------------------------------------- helpers
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="adtfe_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from adt_str_b200.config import SETTING_1
    from adt_str_b200.synthetic import make_bank, make_segments
    bank = make_bank(args.bank_size, SR, seed=0)
    segs = make_segments(4096, seed=1)
    pool, cores = cpu_pool(bank, segs, dict(SETTING_1))
    per_step = args.cpu_segments or min(len(segs), cores * 32)
    try:
        for w in range(args.warmup):
            cpu_run(pool, cores, per_step, seed=1000 * w)
        total_s, total_t = 0.0, 0.0
        for k in range(args.steps):
            s, t = cpu_run(pool, cores, per_step, seed=7 + 1000 * k)
            total_s += s; total_t += t
    finally:
        pool.terminate()
    value = total_s / total_t
    sample = f"{per_step} segments per step ({per_step * INPUT_SEC:.0f} audio-s) of the same workload, {cores} one-thread worker processes"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- GPU library comparator
def library_logmel_module(dev):
    """The reference's ComputeMelSpectrogram (model.py:68-97) as it runs on a GPU: torchaudio MelSpectrogram
    (torch.stft -> cuFFT, |.|^2, MelScale matmul -> cuBLAS) and the reference's own log / clamp / affine / slice
    lines, call for call.  Returns ``f(wave (B, L)) -> (B, T, n_mels)`` (a view, like the reference's)."""
    import torch
    import torchaudio.transforms as T
    hop = int(0.01 * SR)
    spec = T.MelSpectrogram(sample_rate=SR, n_fft=2048, hop_length=hop, n_mels=128, f_min=20.0, power=2).to(dev)
    wpi = int((2048 / 2) // hop + 1)

    def forward(wave):
        with torch.no_grad(), torch.autocast(device_type="cuda", enabled=False):
            x = spec(wave.float())
            x = torch.log(x + 1e-10)
            x = torch.clamp(x, -23, 12)
            x = (x + 23) / (12 + 23)
            return x.permute(0, 2, 1)[:, wpi: -(wpi + 1), :]
    return forward


def measure_traffic(timeout_s=240):
    """roofline.traffic measured BY THIS RUN: a child process replays a resident plan under
    ``ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`` and the log-mel launch's DRAM bytes per segment
    come back.  None when ncu is missing, not permitted or too slow (the committed capture is then used)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    out = tempfile.mktemp(prefix="adtfe_traffic_", suffix=".csv")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:logmel", "-c", "1", "--csv", "--log-file", out, sys.executable, os.path.abspath(__file__),
           "--traffic-child"]
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout_s, env=env)
        n_seg = None
        for line in res.stdout.decode(errors="replace").splitlines():
            if line.startswith("TRAFFIC_CHILD_SEGMENTS"):
                n_seg = int(line.split()[1])
        total = 0.0
        import csv
        with open(out) as f:
            rows = [r for r in csv.reader(f) if len(r) > 3]
        head = next(i for i, r in enumerate(rows) if "Metric Name" in r)
        col = {name: k for k, name in enumerate(rows[head])}
        for r in rows[head + 1:]:
            if r[col["Metric Name"]].startswith("dram__bytes_"):
                v = float(r[col["Metric Value"]].replace(",", ""))
                unit = r[col["Metric Unit"]].lower()
                total += v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        if n_seg and total > 0:
            return total / n_seg
    except Exception:
        pass
    finally:
        try:
            os.unlink(out)
        except OSError:
            pass
    return None


def run_traffic_child(args):
    """Under ncu (see measure_traffic): 32 training batches resident, rendered, then the log-mel launched twice (the
    first launch is skipped by -c 1 only if it were first; both are identical) - waveform rows 0.5 GB >> L2."""
    import torch
    from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum
    from adt_str_b200.config import setting_1
    from adt_str_b200.synthetic import make_bank, make_segments
    from adt_str_b200.synthetiser import PlanBuffers
    dev = torch.device("cuda", 0)
    bank = make_bank(2000, SR, seed=0)
    n_batches = 32
    segs = make_segments(n_batches * BATCH, seed=1)
    fe = FrontEnd(SynthDrum(setting_1(), bank=bank, device=dev), ComputeMelSpectrogram(SR, 2048, 0.01, 128))
    # one render chunk: the log-mel is then ONE launch over all the plan's rows (a chunked plan gets one launch per chunk)
    plan = fe.plan_batches([segs[b * BATCH:(b + 1) * BATCH] for b in range(n_batches)], random.Random(1), n_batches)
    buf = PlanBuffers(dev)
    buf._dplan = buf.upload(buf.pack(plan)); buf._resident = plan
    wav, feat = fe._outputs(plan, 0)
    fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)
    torch.cuda.synchronize(dev)
    print("TRAFFIC_CHILD_SEGMENTS", plan.n_seg, flush=True)
    return 0


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    rank, world, local = dist_env()
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner under some
    # NCCL_DEBUG settings) goes to stderr instead, the line itself to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    from adt_str_b200.config import SETTING_1, setting_1
    from adt_str_b200.synthetic import make_bank, make_segments

    n_batches = args.batches_per_step
    bank = make_bank(args.bank_size, SR, seed=0)                      # replicated: same seed on every rank
    segs = make_segments(n_batches * BATCH, seed=1 + rank)            # each rank its own event streams

    # ---- CPU baseline first: forks must happen before CUDA exists in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pool, cores = cpu_pool(bank, segs, dict(SETTING_1))
        try:
            n = args.cpu_segments or min(len(segs), cores * 256)
            cpu_run(pool, cores, min(n, cores * 8), seed=1)          # warm the workers
            s, t = cpu_run(pool, cores, n, seed=2)
        finally:
            pool.terminate()
        cpu = {"value": s / t, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n} segments ({s:.0f} audio-s) of the step's workload, {cores} one-thread worker "
                         f"processes, {t:.1f} s wall"}

    import torch
    import torch.distributed as dist
    from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum, _lib
    from adt_str_b200.synthetiser import PlanBuffers

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_set = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
        # each rank on the cores next to its GPU, before any pinned buffer exists (first touch = local memory)
        from adt_str_b200.sharding import bind_to_local_cpus
        cpu_set = bind_to_local_cpus(local, rank, world)
    synth = SynthDrum(setting_1(), bank=bank, device=dev)
    mel = ComputeMelSpectrogram(SR, 2048, 0.01, 128)
    fe = FrontEnd(synth, mel)

    # ---- plan the step once and make the plan resident (not timed for `value`): all batches of the
    # step in ONE device plan - one launch of the log-mel kernel, the render chunked per batch
    rng = random.Random(1234 + rank)
    batches = [segs[b * BATCH:(b + 1) * BATCH] for b in range(n_batches)]
    t0 = time.perf_counter()
    plan = fe.plan_batches(batches, rng, args.chunk_batches)
    plan_s = time.perf_counter() - t0
    buf = PlanBuffers(dev)
    buf._dplan = buf.upload(buf.pack(plan))
    buf._resident = plan
    wav, feat = fe._outputs(plan, 0)
    torch.cuda.synchronize(dev)

    audio_s_step = float(int(plan.wave_lengths.sum())) / SR
    n_frames_seg = np.repeat(plan.batch_frames, np.diff(plan.batch_ptr))
    bytes_alg_step = (4 * int(plan.wave_lengths.sum()) + 4 * 128 * int(n_frames_seg.sum()) + plan.bank_bytes(bank)
                      + 32 * plan.n_events)
    n_chunks = len(plan.chunks) - 1
    # OUR kernels per step: peak + slice records (when the chunk has notes), tile mixer, normalise and the chunk's log-mel
    # per chunk (a plan of one chunk: one log-mel launch at the end)
    launches_per_step = int(3 * n_chunks + 2 * (np.diff(plan.chunks["peak_work"]) > 0).sum())

    def step():
        fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)

    # ---- per-kernel share: the two halves alone, CUDA events on the launching stream
    lib, bh, mh = _lib.load(), synth.device_bank().handle, mel._handle(dev).handle
    import ctypes as C
    st = torch.cuda.current_stream(dev).cuda_stream

    def time_loop(fn, reps=3):
        fn(); torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    def render_all():
        _lib.check(lib.adtfe_render(bh, C.byref(buf._dplan), wav.data_ptr(), buf.workspace.data_ptr(),
                                    buf.workspace.numel(), st))

    def logmel_all():
        _lib.check(lib.adtfe_logmel_rows(mh, wav.data_ptr(), plan.n_seg, plan.ld_wav, buf._dplan.mel_rows_dev,
                                         buf._dplan.mel_max_count, feat.data_ptr(), st))

    render_ms, logmel_ms = time_loop(render_all), time_loop(logmel_all)
    # latency of ONE training batch through the same entry (its own plan, resident)
    one_plan = synth.plan(batches[0], random.Random(5))
    one_buf = PlanBuffers(dev)
    one_buf._dplan = one_buf.upload(one_buf.pack(one_plan)); one_buf._resident = one_plan
    one_w, one_f = fe._outputs(one_plan, int(one_plan.wave_lengths.max()))
    one = time_loop(lambda: fe.run_plan(one_plan, buffers=one_buf, wav=one_w, feat=one_f, upload=False), reps=20)

    # ---- the GPU library comparator (the reference's own ComputeMelSpectrogram on the same B200):
    # torchaudio MelSpectrogram + log / clamp / affine over the SAME resident waveform matrix, one call per collated
    # batch of 64 as ADTModel.forward makes them (model.py:248), CUDA-event timed like logmel_ms above
    library = None
    if rank == 0 and not args.no_library_baseline:
        try:
            lib_forward = library_logmel_module(dev)
            lib_dram = None
            try:
                with open(os.path.join(ROOT, "profiles", "r02_library_traffic.json")) as f:
                    lib_dram = float(json.load(f)["dram_bytes_per_segment"]) * plan.n_seg
            except Exception:
                pass
            views = [wav[int(plan.batch_ptr[b]):int(plan.batch_ptr[b + 1]), :int(plan.batch_samples[b])]
                     for b in range(n_batches)]
            render_all(); torch.cuda.synchronize(dev)

            def library_all():
                for v in views:
                    lib_forward(v)

            lib_ms = time_loop(library_all, reps=2)
            ours_one = mel(views[0]); theirs_one = lib_forward(views[0])
            diff = float((ours_one - theirs_one).abs().max())
            lib_one = time_loop(lambda: lib_forward(views[0]), reps=20)
            ours_one_ms = time_loop(lambda: mel(views[0]), reps=20)
            library = {"what": "torchaudio MelSpectrogram (cuFFT + cuBLAS) + log/clamp/affine on the same GPU over the same "
                               "resident waveform rows, one call per batch of 64 (reference model.py:71-97)",
                       "logmel_ms_per_step": lib_ms, "audio_s_per_s": audio_s_step / (lib_ms * 1e-3),
                       "this_repo_logmel_ms_per_step": logmel_ms, "speedup_logmel": lib_ms / logmel_ms,
                       "single_batch_ms": lib_one, "this_repo_single_batch_ms": ours_one_ms,
                       "max_abs_diff_first_batch": diff,
                       "dram_bytes": lib_dram, "dram_bytes_source": "profiles/r02_library_traffic.json x segments (ncu launch list "
                                                                    "of this leg with DRAM bytes: profiles/r02_launches.txt)"}
        except Exception as exc:   # never lose the headline line to the comparator
            library = {"error": repr(exc)}

    # ---- end to end: host notes -> plan -> pinned blob -> H2D -> kernels -> D2H log-mel (pinned), in groups of
    # batches through HostPipeline: planner threads (own RNG streams, like DataLoader workers), rotating buffer
    # sets, everything asynchronous on the current stream
    from adt_str_b200 import HostPipeline
    group = max(1, min(args.e2e_group, n_batches))
    groups = [batches[i:i + group] for i in range(0, n_batches, group)]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    args.e2e_workers = args.e2e_workers or max(1, min(8, cores if cpu_set else cores // max(1, world)))
    pipe = HostPipeline(fe, workers=args.e2e_workers, n_sets=args.e2e_sets, seed=99 + rank, chunk_batches=args.chunk_batches)
    h2d = d2h = 0
    e2e_checksum = 0.0

    def e2e_run(n_steps):
        """`n_steps` steps through the pipeline as one continuous stream of groups (a loader does not drain its
        pipeline between steps); every result is waited for and touched on the host."""
        nonlocal h2d, d2h, e2e_checksum
        h2d = d2h = 0
        inflight = []
        stream_of_groups = (g for _ in range(n_steps) for g in groups)
        for res in pipe.run(stream_of_groups):
            inflight.append(res)
            h2d += res.h2d_bytes
            d2h += res.d2h_bytes
            if len(inflight) > 2:
                r = inflight.pop(0).wait()
                e2e_checksum += float(r.batches()[0][1][0, 0, 0])   # touch the host result
                r.release()
        for r in inflight:
            r.wait().release()
        h2d //= n_steps
        d2h //= n_steps

    e2e_run(1)  # warm-up (allocations, pinned buffers)
    barrier()
    n_e2e = max(1, args.e2e_steps)
    t0 = time.perf_counter()
    e2e_run(n_e2e)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    clocks = sampler.stop() if sampler else None
    pipe.close()

    # ---- the copy ceiling of the end-to-end number: every rank moves one step's log-mel bytes device -> pinned host at
    # the same time, nothing else running (no planning, no kernels) - what the host's links allow at this rank count
    pin = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    srcb = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    n_copies = max(1, d2h // pin.numel())
    pin.copy_(srcb, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_copies):
        pin.copy_(srcb, non_blocking=True)
    barrier()
    d2h_ceiling_ms = 1e3 * (time.perf_counter() - t0) * (d2h / (n_copies * pin.numel()))
    del pin, srcb

    # ---- BASELINE configs[4] beside the headline: the long-form inference front (tools/bench_longform.py)
    long_form = None
    if rank == 0 and world == 1 and not args.no_long_form:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_longform
            long_form = bench_longform.measure(dev, 600.0, cpu_seconds=0.0 if args.no_cpu_baseline else 600.0)
        except Exception as exc:  # never lose the headline line to the side measurement
            long_form = {"error": repr(exc)}

    # ---- beside the headline: the stock training configuration has the FX chain ON (setting-1.yaml:58-61,
    # use_fx_prob 0.3); the same render + log-mel with it, on 64 batches, next to the same 64 batches without
    fx_side = None
    proj_side = None
    if rank == 0 and world == 1 and not args.no_long_form:
        try:
            nb = min(64, n_batches)
            res = {}
            for tag, prob in (("off", 0.0), ("on", 0.3)):
                s2 = SynthDrum(setting_1(use_fx_prob=prob), bank=bank, device=dev)
                s2._device_bank = synth.device_bank()                       # the same resident bank
                fe2 = FrontEnd(s2, mel)
                torch.manual_seed(7)
                p2 = fe2.plan_batches(batches[:nb], random.Random(77), min(args.chunk_batches, 16))
                b2 = PlanBuffers(dev)
                b2._dplan = b2.upload(b2.pack(p2)); b2._resident = p2
                w2, f2 = fe2._outputs(p2, 0)
                res[tag] = time_loop(lambda: fe2.run_plan(p2, buffers=b2, wav=w2, feat=f2, upload=False), reps=3)
                n_fx_rows = 0 if p2.fx is None else len(p2.fx)
                # the same batches end to end (host note lists -> pinned host log-mel) through the pipeline
                pipe2 = HostPipeline(fe2, workers=args.e2e_workers, n_sets=args.e2e_sets, seed=5, chunk_batches=args.chunk_batches)
                g2 = [batches[i:i + group] for i in range(0, nb, group)]
                for rep in range(7):   # the shortest of six runs behind a warm-up (20 ms of wall clock each)
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    infl = []
                    for r2 in pipe2.run(g2):
                        infl.append(r2)
                        if len(infl) > 2:
                            infl.pop(0).wait().release()
                    for r2 in infl:
                        r2.wait().release()
                    if rep:
                        res["e2e_" + tag] = min(res.get("e2e_" + tag, 1e30), 1e3 * (time.perf_counter() - t0))
                pipe2.close()
            fx_side = {"what": f"render + log-mel of {nb} batches with use_fx_prob = 0.3 (reverb / compressor / limiter kernels, "
                               "csrc/fx.cu) against the same batches without FX",
                       "ms_fx_off": res["off"], "ms_fx_on": res["on"], "segments_with_fx": n_fx_rows,
                       "segments": nb * BATCH, "e2e_ms_fx_off": res["e2e_off"], "e2e_ms_fx_on": res["e2e_on"],
                       "e2e": "the same batches through HostPipeline (host planning, H2D, kernels, log-mel D2H)"}
        except Exception as exc:
            fx_side = {"error": repr(exc)}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_projection
            proj_side = bench_projection.measure(dev, rows=int(n_frames_seg.sum()))
            proj_side["single_batch"] = {k: v for k, v in bench_projection.measure(dev, rows=BATCH * 246, reps=20).items()
                                         if k in ("rows", "ms", "library_ms", "speedup_vs_autocast_linear")}
        except Exception as exc:
            proj_side = {"error": repr(exc)}

    # ---- reduce over ranks: units add, time is the max
    from adt_str_b200.sharding import reduce_stats
    total_audio, max_ms = reduce_stats(audio_s_step * args.steps, elapsed_ms)
    total_audio_e2e, max_e2e_ms = reduce_stats(audio_s_step, e2e_ms)
    _, max_ceiling_ms = reduce_stats(0.0, d2h_ceiling_ms)
    total_bytes, _ = reduce_stats(float(bytes_alg_step), 0.0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    traffic, traffic_src = None, None
    if world == 1 and not args.no_traffic:
        per_seg = measure_traffic()
        if per_seg:
            traffic = per_seg * plan.n_seg
            traffic_src = ("measured by this run: ncu dram__bytes_read.sum + dram__bytes_write.sum of one log-mel launch "
                           f"over 2048 resident segments in a child process ({per_seg:.0f} B per segment) x segments")
    if traffic is None:
        try:  # DRAM bytes of the log-mel kernel per segment, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                traffic = float(json.load(f)["dram_bytes_per_segment"]) * plan.n_seg
            traffic_src = "profiles/r02_traffic.json: dram bytes per segment of an ncu --set full capture x segments"
        except Exception:
            pass
    n_seg_step = n_batches * BATCH
    width_seg = np.repeat(plan.batch_samples, np.diff(plan.batch_ptr))
    logmel_bytes = 4 * int(width_seg.sum()) + 4 * 128 * int(n_frames_seg.sum())   # collated rows read + log-mel written
    logmel_gbs = logmel_bytes / (logmel_ms * 1e-3) / 1e9
    path_gbs = (total_bytes / world) / (max_ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": total_audio / (max_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "segments_per_step_per_gpu": n_seg_step,
                   "audio_s_per_step_per_gpu": audio_s_step,
                   "launch": f"one plan per step: render in {n_chunks} chunks (peaks + slice records | tile mixer | normalise as a pipeline over three internal streams), the log-mel of a chunk launched behind its normalisation",
                   "l2": "no flush needed: per step 0.5 GB bank + ~6 GB of distinct outputs >> 126 MB L2",
                   "single_batch_latency_ms": one, "plan_ms_per_batch_host": 1e3 * plan_s / n_batches},
        "clocks": clocks,
        "e2e": {"value": total_audio_e2e / (max_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": max(1, args.e2e_steps),
                "d2h_ceiling": {"value": total_audio_e2e / (max_ceiling_ms * 1e-3), "unit": UNIT,
                                "ms_per_step": max_ceiling_ms, "gbs_per_rank": d2h / (max_ceiling_ms * 1e-3) / 1e9,
                                "what": "the same log-mel bytes, device -> pinned host on every rank at once, nothing "
                                        "else running: the host-link ceiling of e2e at this rank count"},
                "frac_of_d2h_ceiling": max_ceiling_ms / max_e2e_ms,
                "cpu_affinity": None if cpu_set is None else f"{len(cpu_set)} cores local to the GPU",
                "includes": f"host planning of note lists ({args.e2e_workers} planner threads), plan blob H2D, kernels, "
                            f"log-mel D2H into pinned host memory; groups of {group} batches, {args.e2e_sets} buffer sets"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "kernel": "logmel6_kernel", "achieved": logmel_gbs, "peak": peak, "unit": "GB/s",
                     "frac": logmel_gbs / peak, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "bytes_per_launch": logmel_bytes, "avg_launch_ms": logmel_ms,
                     "share_of_step": logmel_ms / (max_ms / args.steps),
                     "render_ms_per_step": render_ms, "logmel_ms_per_step": logmel_ms,
                     "path": {"bytes_alg_per_step": total_bytes / world, "achieved": path_gbs, "frac": path_gbs / peak,
                              "frac_of_nominal_8TBs": path_gbs / 8000.0}},
        "cpu_baseline": cpu,
        "gpu_library_baseline": library,
        "long_form": long_form,
        "fx_chain": fx_side,
        "project_to_mel": proj_side,
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.traffic_child:
        return run_traffic_child(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
