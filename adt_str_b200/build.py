"""Build libadtfe.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadtfe.so")
SOURCES = ["api.cu", "mixer.cu", "fx.cu", "logmel.cu", "project.cu", "resample.cu", "planner.cpp"]
DEPS = SOURCES + ["common.cuh", "fft_gen.cuh", os.path.join("..", "..", "include", "adtfe.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libadtfe.so cannot be built")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """``defines`` / ``out``: tuning variants (``-DADTFE_MIX_CONSUMERS=2`` ...) built next to the product
    library and selected with the ``ADTFE_LIB`` environment variable (see ``_lib.py``)."""
    gen = os.path.join(CSRC, "fft_gen.cuh")
    if not os.path.exists(gen):
        subprocess.check_call([sys.executable, os.path.join(HERE, "..", "tools", "gen_fft.py")])
    if out == LIB and not defines and not force and not stale():
        return LIB
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [_nvcc(), *flags, *[f"-D{d}" for d in defines], *(["-Xptxas", "-v"] if verbose else []), "-o", out + ".tmp",
           *[os.path.join(CSRC, s) for s in SOURCES]]
    subprocess.check_call(cmd, cwd=CSRC)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs,
                out=os.path.join(HERE, outs[0]) if outs else LIB))
