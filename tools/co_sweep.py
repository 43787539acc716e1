#!/usr/bin/env python3
"""Sweep of the pipelined front end (adtfe_render_logmel on a chunked plan): step time of the bench workload for
several log-mel group sizes (ADTFE_CO_GROUP, in chunks; 0 = render everything, then one log-mel launch) and render
chunk sizes.  The library variant comes from ADTFE_LIB (tools/gpu_variants.sh convention).

    python tools/co_sweep.py [--groups 0,2,4,8,16] [--chunks 4] [--batches 256] [--steps 6]
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum  # noqa: E402
from adt_str_b200.config import setting_1  # noqa: E402
from adt_str_b200.synthetic import make_bank, make_segments  # noqa: E402
from adt_str_b200.synthetiser import PlanBuffers  # noqa: E402


def summarise(path):
    """Per kernel: launches, mean duration, busy time (sum of durations), first start and last end."""
    import csv
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
    p.add_argument("--groups", default="0,2,4,8,16")
    p.add_argument("--chunks", default="4")
    p.add_argument("--batches", type=int, default=256)
    p.add_argument("--bank-size", type=int, default=10_000)
    p.add_argument("--steps", type=int, default=6)
    p.add_argument("--trace", default="", help="directory for one launch trace (adtfe_trace_dump) per configuration")
    args = p.parse_args()
    tag = os.path.basename(os.environ.get("ADTFE_LIB", "base"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    bank = make_bank(args.bank_size, 24000, seed=0)
    segs = make_segments(args.batches * 64, seed=1)
    batches = [segs[b * 64:(b + 1) * 64] for b in range(args.batches)]
    fe = FrontEnd(SynthDrum(setting_1(), bank=bank, device=dev), ComputeMelSpectrogram(24000, 2048, 0.01, 128))
    ref = None
    for cb in [int(x) for x in args.chunks.split(",")]:
        plan = fe.plan_batches(batches, random.Random(1234), cb)
        buf = PlanBuffers(dev)
        buf._dplan = buf.upload(buf.pack(plan))
        buf._resident = plan
        wav, feat = fe._outputs(plan, 0)
        audio_s = float(int(plan.wave_lengths.sum())) / 24000
        for g in [int(x) for x in args.groups.split(",")]:
            os.environ["ADTFE_CO_GROUP"] = str(g)
            feat.zero_()
            for _ in range(2):
                fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)
            b.record()
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / args.steps
            if args.trace:
                from adt_str_b200 import _lib
                lib = _lib.load()
                lib.adtfe_trace_begin()
                fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)
                path = os.path.join(args.trace, f"trace_{tag.replace('.so', '')}_c{cb}_g{g}.csv")
                _lib.check(lib.adtfe_trace_dump(path.encode()), "adtfe_trace_dump")
                summarise(path)
            chk = (float(feat.double().nan_to_num().sum()), float(wav.double().nan_to_num().abs().sum()))
            ref = ref or chk
            same = "same" if chk == ref else f"DIFFERENT {chk} vs {ref}"
            print(f"{tag} chunk_batches={cb} co_group={g}: {ms:.3f} ms/step  {audio_s / ms / 1e3:.3f} M audio-s/s  {same}",
                  flush=True)


if __name__ == "__main__":
    main()
