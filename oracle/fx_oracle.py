"""ctypes front of ``oracle/fx_oracle.c`` - the CPU restatement of the reference's pedalboard FX chain.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**: pedalboard is not installed here, so
the C file restates the published JUCE algorithms pedalboard wraps (header of ``fx_oracle.c``); what IS pinned on the
running reference is everything around the DSP - which plugins are built, with which parameters, from which RNG
draws, in which order, and where the chain sits (``modules/synthetiser.py:30-87,121-137,154``).  For that the
module also provides ``install_pedalboard_stand_in()``: ``pedalboard.Pedalboard / Reverb / Compressor / Limiter``
classes that record their constructor arguments and, when the board is called, run the DSP below - so the
UNMODIFIED reference renders with FX on (``oracle/ref_harness.py``) and its output can be compared with the GPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import types
from typing import List

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fx_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libfxoracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _LIB + ".tmp", _SRC, "-lm"])
        os.replace(_LIB + ".tmp", _LIB)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        f, i64, i = C.c_float, C.c_int64, C.c_int
        lib.fx_reverb.argtypes = [C.c_void_p, i64, i, f, f, f, f, f]
        lib.fx_compressor.argtypes = [C.c_void_p, i64, i, f, f, f, f]
        lib.fx_limiter.argtypes = [C.c_void_p, i64, i, f, f]
        _lib = lib
    return _lib


def _mono(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, np.float32).reshape(-1)).copy()


def reverb(x, sample_rate: int, room_size: float, damping: float, wet_level: float, dry_level: float,
           width: float) -> np.ndarray:
    y = _mono(x)
    if _load().fx_reverb(y.ctypes.data, y.size, int(sample_rate), room_size, damping, wet_level, dry_level, width):
        raise MemoryError("fx_reverb")
    return y


def compressor(x, sample_rate: int, threshold_db: float, ratio: float, attack_ms: float, release_ms: float) -> np.ndarray:
    y = _mono(x)
    _load().fx_compressor(y.ctypes.data, y.size, int(sample_rate), threshold_db, ratio, attack_ms, release_ms)
    return y


def limiter(x, sample_rate: int, threshold_db: float, release_ms: float = 100.0) -> np.ndarray:
    y = _mono(x)
    _load().fx_limiter(y.ctypes.data, y.size, int(sample_rate), threshold_db, release_ms)
    return y


# ---------------------------------------------------------------------------------- pedalboard stand-in
#: every plugin the stand-in classes were asked to build since the last ``CALLS.clear()``: (class name, kwargs)
CALLS: List[tuple] = []


class _Plugin:
    def __init__(self, **kwargs):
        self.kwargs = {k: float(np.float32(v)) for k, v in kwargs.items()}   # pybind converts to C++ float
        CALLS.append((type(self).__name__, dict(kwargs)))


class Reverb(_Plugin):
    def process(self, x, sr):
        k = self.kwargs
        return reverb(x, sr, k["room_size"], k["damping"], k["wet_level"], k["dry_level"], k["width"])


class Compressor(_Plugin):
    def process(self, x, sr):
        k = self.kwargs
        return compressor(x, sr, k["threshold_db"], k["ratio"], k["attack_ms"], k["release_ms"])


class Limiter(_Plugin):
    def process(self, x, sr):
        k = self.kwargs
        return limiter(x, sr, k["threshold_db"], k.get("release_ms", 100.0))


class Pedalboard(list):
    """``board(x (S, C), sample_rate) -> (S, C)``: the plugins in order, every channel on its own (mono here)."""

    def __call__(self, x, sample_rate):
        x = np.asarray(x, np.float32)
        cols = []
        for c in range(x.shape[1]):
            y = x[:, c]
            for plugin in self:
                y = plugin.process(y, sample_rate)
            cols.append(y)
        return np.stack(cols, axis=1)


def install_pedalboard_stand_in() -> None:
    """Put the recording / oracle-DSP classes in ``sys.modules['pedalboard']`` (before the reference is imported)."""
    m = types.ModuleType("pedalboard")
    m.Pedalboard, m.Reverb, m.Compressor, m.Limiter = Pedalboard, Reverb, Compressor, Limiter
    m.__adtfe_stand_in__ = True
    sys.modules["pedalboard"] = m
