#!/usr/bin/env python3
"""Step time of the bench workload (one plan of --batches training batches, resident) for several render chunk
sizes, with one launch trace per configuration (adtfe_trace_begin / _dump): per kernel the launches, the mean
duration, the busy time and the span.  The library variant comes from ADTFE_LIB (tools/gpu_variants.sh convention).

    python tools/render_sweep.py [--chunks 64,16] [--batches 256] [--steps 6] [--trace gpurun_out]
"""
import argparse
import csv
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum, _lib  # noqa: E402
from adt_str_b200.config import setting_1  # noqa: E402
from adt_str_b200.synthetic import make_bank, make_segments  # noqa: E402
from adt_str_b200.synthetiser import PlanBuffers  # noqa: E402


def summarise(path):
    rows = list(csv.DictReader(open(path)))
    kinds = {}
    for r in rows:
        kinds.setdefault(r["kernel"], []).append((float(r["start_ms"]), float(r["end_ms"])))
    for k, v in kinds.items():
        d = [b - a for a, b in v]
        print(f"    {k:10s} n={len(v):3d} mean {sum(d) / len(d):7.3f} ms  sum {sum(d):7.3f} ms  "
              f"span {min(a for a, _ in v):7.3f} .. {max(b for _, b in v):7.3f} ms", flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--chunks", default="64")
    p.add_argument("--batches", type=int, default=256)
    p.add_argument("--bank-size", type=int, default=10_000)
    p.add_argument("--steps", type=int, default=6)
    p.add_argument("--trace", default="", help="directory for one launch trace per configuration")
    p.add_argument("--render-only", action="store_true")
    args = p.parse_args()
    tag = os.path.basename(os.environ.get("ADTFE_LIB", "base")).replace(".so", "")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    bank = make_bank(args.bank_size, 24000, seed=0)
    segs = make_segments(args.batches * 64, seed=1)
    batches = [segs[b * 64:(b + 1) * 64] for b in range(args.batches)]
    synth = SynthDrum(setting_1(), bank=bank, device=dev)
    fe = FrontEnd(synth, ComputeMelSpectrogram(24000, 2048, 0.01, 128))
    lib = _lib.load()
    import ctypes as C
    for cb in [int(x) for x in args.chunks.split(",")]:
        plan = fe.plan_batches(batches, random.Random(1234), cb)
        buf = PlanBuffers(dev)
        buf._dplan = buf.upload(buf.pack(plan))
        buf._resident = plan
        wav, feat = fe._outputs(plan, 0)
        st = torch.cuda.current_stream(dev).cuda_stream

        def step():
            if args.render_only:
                _lib.check(lib.adtfe_render(synth.device_bank().handle, C.byref(buf._dplan), wav.data_ptr(),
                                            buf.workspace.data_ptr(), buf.workspace.numel(), st))
            else:
                fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)

        for _ in range(2):
            step()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        b.record()
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b) / args.steps
        print(f"{tag} chunk_batches={cb}: {ms:.3f} ms/step ({'render only' if args.render_only else 'render + log-mel'})", flush=True)
        if args.trace:
            lib.adtfe_trace_begin()
            step()
            path = os.path.join(args.trace, f"trace_{tag}_c{cb}.csv")
            _lib.check(lib.adtfe_trace_dump(path.encode()), "adtfe_trace_dump")
            summarise(path)


if __name__ == "__main__":
    main()
