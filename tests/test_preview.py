"""The preview renderer (reference utils/drum_audio_render.py:130-194): oracle restatement against the fixture frozen
from the running reference (and against the reference itself where /root/reference exists), GPU path against the
fixture."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import preview_oracle


def _fixture():
    z = np.load(os.path.join(GOLDEN_DIR, "preview.npz"))
    oneshots = {int(p): z[f"os_{int(p)}"] for p in z["pitches"]}
    return z["notes"], int(z["num_samples"]), int(z["sample_rate"]), oneshots, z["wav_mapped"], z["wav_plain"]


def test_preview_oracle_equals_the_fixture_bit_for_bit():
    notes, n, sr, oneshots, mapped, plain = _fixture()
    assert np.array_equal(preview_oracle.synthesize_drums_procedural(notes, n, sr, oneshots, True), mapped)
    assert np.array_equal(preview_oracle.synthesize_drums_procedural(notes, n, sr, oneshots, False), plain)
    assert not np.array_equal(mapped, plain)                                  # the mapping changes which notes sound
    assert not preview_oracle.synthesize_drums_procedural(notes[:0], 100, sr, oneshots).any()


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils"), reason="the reference tree is only in the build container")
def test_preview_oracle_equals_the_running_reference():
    sys.path.insert(0, "/root/reference")
    try:
        import utils.drum_audio_render as ref
    finally:
        sys.path.remove("/root/reference")
    notes, n, sr, oneshots, _, _ = _fixture()
    ref._ONESHOT_CACHE.clear()
    ref._ONESHOT_CACHE.update(oneshots)          # what get_oneshot_waveform would have read from one-shot-rendering/
    rng = np.random.default_rng(3)
    for _ in range(3):
        sel = notes[np.sort(rng.choice(len(notes), 40, replace=False))]
        for mapping in (True, False):
            want = ref.synthesize_drums_procedural(sel, n - 1000, sr, apply_mapping=mapping)
            got = preview_oracle.synthesize_drums_procedural(sel, n - 1000, sr, oneshots, mapping)
            assert np.array_equal(want, got)
    wav, mode = ref.render_drum_preview(notes, n, sr)
    assert mode == "oneshot" and np.array_equal(wav.numpy(), preview_oracle.synthesize_drums_procedural(notes, n, sr, oneshots, False))


def test_gm_mapping_table_equals_the_product_table():
    from adt_str_b200.midi_tokenizer import GM_STANDARD_TO_CUSTOM
    assert tuple(GM_STANDARD_TO_CUSTOM[p] for p in range(35, 82)) == preview_oracle.GM_CUSTOM_OF


@pytest.mark.gpu
def test_preview_renderer_matches_the_reference_fixture():
    """One adtfe_render call per preview; what differs from the reference is one rounding per addition (fused
    multiply-add): 1e-6 of full scale."""
    import torch
    from adt_str_b200.preview import PreviewRenderer, render_drum_preview
    notes, n, sr, oneshots, mapped, plain = _fixture()
    r = PreviewRenderer(oneshots, sr)
    for mapping, want in ((True, mapped), (False, plain)):
        got = r.synthesize(notes, n, apply_mapping=mapping)
        assert got.shape == (n,) and got.dtype == torch.float32 and got.is_cuda
        assert np.abs(got.cpu().numpy() - want).max() <= 1e-6
    assert abs(float(np.abs(r.synthesize(notes, n).cpu().numpy()).max()) - 0.98) < 1e-6
    assert not r.synthesize(notes[:0], 500).any() and r.synthesize(notes[:0], 500).shape == (500,)
    late = notes.copy(); late[:, 0] += 100.0                                  # every onset past the buffer: silence
    assert not r.synthesize(late, n).any()
    wav, mode = render_drum_preview(torch.from_numpy(notes), n, sr, oneshots)  # tensor notes, mapping off by default
    assert mode == "oneshot" and not wav.is_cuda and np.abs(wav.numpy() - plain).max() <= 1e-6
