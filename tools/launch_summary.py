#!/usr/bin/env python3
"""Summary of an ncu launch list (``--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv``
of a small bench run): per kernel the launches, total time and DRAM bytes; the shares among THIS repo's kernels
(adtfe::*) - the figure the bench's ``share_of_step`` has to agree with - and, separately, the kernels of the
reference's GPU library pipeline for the log-mel (torchaudio MelSpectrogram: reflection pad, cuFFT, abs/pow, the mel
SGEMM, log / clamp / affine) that the bench's ``gpu_library_baseline`` leg launches.

    python tools/launch_summary.py gpurun_out/launches_r02.csv > profiles/r02_launches.txt
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {n: i for i, n in enumerate(rows[h])}
    scale_t = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    t, n, rd, wr = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float), collections.defaultdict(float)
    for r in rows[h + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        k = r[col["Kernel Name"]].split("(")[0].strip()[:70]
        v = float(r[col["Metric Value"]].replace(",", ""))
        m, u = r[col["Metric Name"]], r[col["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            t[k] += v * scale_t.get(u, 1.0)
            n[k] += 1
        elif m.startswith("dram__bytes_read"):
            rd[k] += v * scale_b.get(u, 1.0)
        elif m.startswith("dram__bytes_write"):
            wr[k] += v * scale_b.get(u, 1.0)
    mine = ("logmel6_kernel", "logmel_kernel", "mix_kernel", "normalise_kernel", "peak_kernel", "slice_kernel", "blockmax_kernel",
            "fx_reverb_kernel", "fx_dynamics_kernel", "fx_mark_kernel", "project_kernel", "resample_kernel", "downmix_kernel")
    # ncu prints the names with or without the namespace, depending on its demangler
    ours = {k: v for k, v in t.items() if "adtfe::" in k or any(m in k for m in mine)}
    lib = {k: v for k, v in t.items() if k not in ours}
    tot_o = sum(ours.values()) or 1.0
    print(f"# {path}: per-launch times are cold-cache and serialised (compare shares, not absolutes)")
    print("\n== this repo's kernels")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
        print(f"{k:72s} n={n[k]:4d} {v:10.1f} us {100 * v / tot_o:5.1f} %   dram read {rd[k] / 1e6:9.1f} MB  write {wr[k] / 1e6:9.1f} MB")
    print("\n== other kernels in the run (the GPU library comparator: torchaudio MelSpectrogram + log / clamp / affine; ATen copies)")
    for k, v in sorted(lib.items(), key=lambda kv: -kv[1])[:16]:
        print(f"{k:72s} n={n[k]:4d} {v:10.1f} us   dram read {rd[k] / 1e6:9.1f} MB  write {wr[k] / 1e6:9.1f} MB")
    fft = [k for k in lib if "fft" in k.lower()]
    if fft:
        calls = max(n[k] for k in fft)
        lt = sum(lib.values())
        lb = sum(rd[k] + wr[k] for k in lib)
        print(f"\nlibrary pipeline: {calls} calls of one batch of 64 segments each: {lt / calls:.1f} us and "
              f"{lb / calls / 1e6:.1f} MB of DRAM traffic per call = {lb / calls / 64 / 1e6:.2f} MB per segment (ncu flushes the "
              "caches between kernels, so every intermediate is counted once written and once read); the fused log-mel "
              "kernel moves 0.35 MB per segment (bench.py roofline.traffic, measured in the same way)")


if __name__ == "__main__":
    main(sys.argv[1])
