"""The C-ABI library loads and exports every symbol include/adtfe.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from adt_str_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()  # no-op when libadtfe.so is current
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "adtfe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adtfe_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in adtfe.h but not exported"
    assert sorted(_lib.EXPORTS) == names      # the binding covers the whole header


def test_struct_sizes_match_numpy_views():
    from adt_str_b200.planner import EVENT_DTYPE, PEAK_ITEM_DTYPE, SEGMENT_DTYPE
    assert EVENT_DTYPE.itemsize == 32 and SEGMENT_DTYPE.itemsize == 16 and PEAK_ITEM_DTYPE.itemsize == 40
    # ... + (n_tile_events, n_fx) + fx_dev + (sample_rate + padding)
    assert C.sizeof(_lib.Plan) == 5 * 8 + 4 * 4 + 8 + 8 + 8 + 4 + 4 + 8 + 8 + 8 + 8
    from adt_str_b200.planner import CHUNK_DTYPE, FX_DTYPE, MEL_ROW_DTYPE
    assert MEL_ROW_DTYPE.itemsize == 16 and CHUNK_DTYPE.itemsize == 16 and FX_DTYPE.itemsize == 48


def test_version_and_argument_errors_without_a_device(lib):
    assert lib.adtfe_version() == 5
    assert lib.adtfe_render_workspace_bytes(10, 2, 30, 70) >= 10 * 48 + 2 * 30 * 4 + 70 * 32
    assert lib.adtfe_render_workspace_bytes(-1, 2, 30, 70) == 0
    shape = _lib.Plan(None, None, None, None, None, 100, 4, 31, 9, 63488)
    off = (C.c_size_t * 7)()
    total = C.c_size_t()
    assert lib.adtfe_plan_blob_layout(C.byref(shape), C.byref(off), C.byref(total)) == 0
    assert list(off)[:3] == [0, 3200, 3264] and all(o % 16 == 0 for o in off) and total.value == off[6]
    assert off[4] - off[3] == 368 and off[6] == off[5] == off[4]   # no mel rows / FX sections without batches / FX
    shape.mel_total_rows = 1000
    shape.n_fx = 3
    assert lib.adtfe_plan_blob_layout(C.byref(shape), C.byref(off), C.byref(total)) == 0
    assert off[5] - off[4] == 4 * 16                            # one adtfe_mel_row per segment
    assert off[6] - off[5] == 3 * 48                            # one adtfe_fx per segment with an FX chain
    assert lib.adtfe_plan_blob_layout(None, C.byref(off), C.byref(total)) == -1
    assert b"null" in lib.adtfe_last_error()
    # null handles are rejected before any CUDA call
    assert lib.adtfe_render(None, C.byref(shape), None, None, 0, None) == -1
    first, count = C.c_int32(), C.c_int32()
    assert lib.adtfe_mel_frames(None, 61440, C.byref(first), C.byref(count)) == -1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from adt_str_b200 import ComputeMelSpectrogram, SynthDrum, setting_1
    from adt_str_b200.synthetic import make_bank
    with pytest.raises(RuntimeError, match="no CPU path"):
        ComputeMelSpectrogram(24000, 2048, 0.01, 128)(torch.zeros(1, 61440))
    synth = SynthDrum(setting_1(), bank=make_bank(78, min_len=100, max_len=500))
    with pytest.raises(RuntimeError, match="no CPU path"):
        synth([[0.1, 0.2, 36, 100]])
    assert synth([]).shape == (61440,)        # the reference's early return needs no device either
    h = C.c_void_p()
    assert _lib.load().adtfe_bank_create(None, 0, None, None, 0, 0, C.byref(h)) == -5   # ADTFE_ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "adt_str_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
