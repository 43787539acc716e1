#!/usr/bin/env python3
"""Phase timeline of logmel_kernel from an ADTFE_TRACE build (clock64 stamps per warp).
Builds adt_str_b200/libadtfe_trace.so, runs one 64-segment launch, prints mean cycles per phase."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from adt_str_b200 import _lib, build

def main():
    so = os.path.join(ROOT, "adt_str_b200", "libadtfe_trace.so")
    flags = [f for f in build.NVCC_FLAGS if f != "--use_fast_math=false"]
    subprocess.check_call([build._nvcc(), *flags, "-DADTFE_TRACE", "-o", so, *[os.path.join(build.CSRC, s) for s in build.SOURCES]])
    _lib.LIB_PATH = so
    from adt_str_b200.mel import ComputeMelSpectrogram
    mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    x = torch.randn(64, 63840, device="cuda")
    n_cta = 148
    trace = torch.zeros(n_cta * 16 * 3 * 16, dtype=torch.int64, device="cuda")
    os.environ["ADTFE_TRACE_PTR"] = hex(trace.data_ptr())
    for _ in range(3):
        trace.zero_(); y = mel(x); torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(n_cta, 16, 3, 16)
    names = ["win0", "rdft0", "tw+xchg0", "cdft0", "pstore0", "load+win1", "rdft1", "tw+xchg1", "cdft1", "pstore1",
             "wts+prefetch", "barrier1", "mel", "barrier2"]
    for r in range(3):
        tt = t[:, :, r, :]
        ok = tt[:, :, 14] > 0
        d = np.diff(tt[:, :, :15], axis=2)[ok]
        print(f"round {r}: warps {ok.sum()}  total {int((tt[:,:,14]-tt[:,:,0])[ok].mean())} cycles")
        print("   " + "  ".join(f"{n}={int(m)}" for n, m in zip(names, d.mean(axis=0))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): mel(x)
    e1.record(); torch.cuda.synchronize()
    print("us per launch", e0.elapsed_time(e1) / 20 * 1e3)

if __name__ == "__main__":
    main()
