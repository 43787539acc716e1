"""Oracle and planner against the LIVE reference (build container only; skipped elsewhere)."""
import random
import warnings

import numpy as np
import pytest

from oracle import mel_oracle, ref_harness, synth_oracle
from adt_str_b200.config import SETTING_1, SynthDrumConfig
from adt_str_b200.planner import plan_segment
from adt_str_b200.synthetic import make_bank, make_dense_segment, make_segments

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")
warnings.filterwarnings("ignore")


@pytest.mark.parametrize("sr,tau,mix", [(24000, 0.8, 0.8), (16000, 0.4, 0.0)])
def test_seeded_reference_run_equals_oracle_and_planner(sr, tau, mix):
    import torch
    cfg = dict(SETTING_1, sample_rate=sr, similarity_threshold=tau, mixup_range=mix, oneshot_path=f"live_{sr}")
    bank = make_bank(156, sample_rate=sr, max_len=6000, seed=3,
                     groups=("gold", "100-90", "90-80", "70-60", "50-40"))
    nested = bank.to_nested()
    segs = make_segments(10, seed=sr) + [make_dense_segment()]
    ref = ref_harness.make_synth(cfg, nested)
    random.seed(77)
    ref_w = [ref(s if len(s) else []).numpy() for s in segs]
    state_after_ref = random.getstate()
    random.seed(77)
    ora_w = [synth_oracle.render(s, cfg, nested) for s in segs]
    assert random.getstate() == state_after_ref          # same number of RNG draws
    for a, b in zip(ref_w, ora_w):
        assert len(a) == len(b) and np.abs(a - b).max() < 1e-6
    random.seed(77)
    plans = [plan_segment(s, SynthDrumConfig(**cfg), bank) for s in segs]
    assert random.getstate() == state_after_ref
    assert [p.wave_length for p in plans] == [len(w) for w in ref_w]
    mel = ref_harness.make_mel(sr, 2048, 0.01, 128)
    batch = synth_oracle.collate(ref_w)
    r = mel(torch.from_numpy(batch)).numpy()
    d = mel_oracle.logmel_direct(batch, sr, 2048, 0.01, 128, np.float32,
                                 fb=mel.compute_spec.mel_scale.fb.numpy(),
                                 window=mel.compute_spec.spectrogram.window.numpy())
    assert np.abs(r - d).max() < 1e-6


def test_float32_index_rule_probe():
    """int((f32(2.5) + 0.1) * 24000) is 62399 in the reference (62400 in float64)."""
    cfg = dict(SETTING_1, oneshot_path="live_probe")
    bank = make_bank(78, max_len=2000, seed=4)
    ref = ref_harness.make_synth(cfg, bank.to_nested())
    notes = [[0.5, 2.5, 36.0, 100.0]]
    random.seed(1)
    n_ref = len(ref(notes))
    random.seed(1)
    assert n_ref == 62399 == plan_segment(notes, SynthDrumConfig(**cfg), bank).wave_length
    random.seed(1)
    assert plan_segment(np.array(notes, np.float64), SynthDrumConfig(**cfg), bank).wave_length == 62400


def test_mel_buffers_equal_reference_state_dict():
    from adt_str_b200.mel import ComputeMelSpectrogram
    import torch
    for sr in (24000, 16000):
        ref = ref_harness.make_mel(sr, 2048, 0.01, 128)
        ours = ComputeMelSpectrogram(sr, 2048, 0.01, 128)
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        for k in rs:
            assert torch.equal(rs[k], os_[k]), k
        ours.load_state_dict(rs, strict=True)
        assert ours.window_pad_idxs == ref.window_pad_idxs
