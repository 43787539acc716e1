"""FX chain (reference modules/synthetiser.py:30-87, 121-137, 154): planning against the running reference.

pedalboard is not installed here, so parity of the DSP itself is UNPINNED (oracle/fx_oracle.c restates the published
JUCE algorithms).  What these tests pin on the unmodified reference is everything around the DSP: that an FX coin hit
builds the plugins the reference builds, with the parameter values the reference draws, consuming the ``random`` and
the torch streams exactly as the reference does - through a stand-in for pedalboard that records the constructor
calls (``oracle.fx_oracle``).  The GPU kernels are compared with the oracle in tests/test_gpu_fx.py.
"""
import random

import numpy as np
import pytest
import torch

from adt_str_b200 import planner
from adt_str_b200.config import SETTING_1, setting_1
from adt_str_b200.synthetic import make_bank, make_segments
from oracle import fx_oracle, ref_harness

needs_reference = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")

FX_CFG = dict(SETTING_1, use_fx_prob=0.6, use_reverb_prob=0.5, use_compression_prob=0.5, use_limiter_prob=0.5)


def _expected_records(calls_per_segment):
    """plugin constructor logs of the reference, one list per SynthDrum call -> FX_DTYPE-like dicts"""
    out = []
    for calls in calls_per_segment:
        if calls is None:
            out.append(None)
            continue
        rec = {"flags": 0}
        for name, kw in calls:
            if name == "Reverb":
                rec["flags"] |= planner.FX_REVERB
                rec.update(room_size=kw["room_size"], damping=kw["damping"], wet_level=kw["wet_level"],
                           dry_level=kw["dry_level"], width=kw["width"])
                assert kw["freeze_mode"] == 0.0
            elif name == "Compressor":
                rec["flags"] |= planner.FX_COMPRESSOR
                rec.update(comp_threshold_db=kw["threshold_db"], comp_ratio=kw["ratio"], comp_attack_ms=kw["attack_ms"],
                           comp_release_ms=kw["release_ms"])
            elif name == "Limiter":
                rec["flags"] |= planner.FX_LIMITER
                rec.update(lim_threshold_db=kw["threshold_db"])
        out.append(rec)
    return out


def _run_reference(segs, bank, seed):
    """The unmodified reference with FX on: per segment the waveform and the plugins it built (None: no coin hit)."""
    fx = ref_harness.enable_fx_stand_in()
    ref = ref_harness.make_synth(dict(FX_CFG), bank.to_nested())
    random.seed(seed)
    torch.manual_seed(seed)
    wavs, logs = [], []
    for notes in segs:
        fx.CALLS.clear()
        hit = [False]
        mixer_cls = ref_harness._import_reference()[0].VolumeMixer
        orig = mixer_cls._add_fx

        def spy(self, x, _orig=orig, _hit=hit):
            _hit[0] = True
            return _orig(self, x)

        mixer_cls._add_fx = spy
        try:
            wavs.append(ref(notes).numpy())
        finally:
            mixer_cls._add_fx = orig
        logs.append(list(fx.CALLS) if hit[0] else None)
    return wavs, logs, random.getstate(), torch.get_rng_state()


@needs_reference
@pytest.mark.parametrize("native", [False, True])
def test_fx_plan_equals_the_reference_draw_for_draw(native):
    bank = make_bank(390, 24000, seed=31, min_len=400, max_len=6000)
    segs = make_segments(40, seed=32, empty_fraction=0.1)
    _, logs, rstate, tstate = _run_reference(segs, bank, seed=77)
    want = _expected_records(logs)
    assert sum(w is not None for w in want) >= 10 and any(w and w["flags"] == 7 for w in want)

    cfg = setting_1(**{k: FX_CFG[k] for k in ("use_fx_prob", "use_reverb_prob", "use_compression_prob", "use_limiter_prob")})
    random.seed(77)
    torch.manual_seed(77)
    if native:
        from adt_str_b200.native_planner import NativePlanner
        plan = NativePlanner(cfg, bank).plan_batch([np.asarray(s, np.float32).reshape(-1, 4) for s in segs], random)
    else:
        plan = planner.plan_batch(segs, cfg, bank)
    assert random.getstate() == rstate                      # the `random` stream ends where the reference's does
    assert torch.equal(torch.get_rng_state(), tstate)       # and so does torch's
    got = {int(r["seg"]): r for r in (plan.fx if plan.fx is not None else [])}
    assert sorted(got) == [s for s, w in enumerate(want) if w is not None]
    for s, w in enumerate(want):
        if w is None:
            continue
        r = got[s]
        assert int(r["flags"]) == w["flags"]
        for key, val in w.items():
            if key != "flags":
                assert r[key] == np.float32(val), (s, key)   # pybind hands pedalboard the float32 of the python float
    assert plan.sample_rate == 24000


@needs_reference
def test_reference_with_fx_equals_oracle_render_plus_oracle_fx():
    """Where the chain sits: instrument sum -> FX -> / max|.| * max_volume (synthetiser.py:149-156).  The synth oracle's
    raw mix through the FX oracle, then normalised, is what the reference (with the same DSP plugged in) returns."""
    from oracle import synth_oracle
    bank = make_bank(390, 24000, seed=33, min_len=400, max_len=6000)
    segs = [s for s in make_segments(12, seed=34, empty_fraction=0.0)]
    wavs, logs, _, _ = _run_reference(segs, bank, seed=5)
    random.seed(5)
    nested = bank.to_nested()
    n_hit = 0
    for notes, ref_wav, calls in zip(segs, wavs, logs):
        raw, vol = synth_oracle.render(notes, dict(FX_CFG), nested, raw=True)
        assert random.random() is not None                  # the FX coin the oracle render does not draw itself
        if calls is not None:
            n_hit += 1
            for name, kw in calls:
                kw = {k: float(np.float32(v)) for k, v in kw.items()}
                if name == "Reverb":
                    raw = fx_oracle.reverb(raw, 24000, kw["room_size"], kw["damping"], kw["wet_level"], kw["dry_level"],
                                           kw["width"])
                    for _ in range(4):
                        random.random()
                elif name == "Compressor":
                    raw = fx_oracle.compressor(raw, 24000, kw["threshold_db"], kw["ratio"], kw["attack_ms"], kw["release_ms"])
                else:
                    raw = fx_oracle.limiter(raw, 24000, kw["threshold_db"])
            for _ in range(3):
                random.random()                             # the three plugin coins
        got = raw / np.abs(raw).max() * vol
        assert np.abs(got - ref_wav).max() <= 2e-6
    assert n_hit >= 3


def test_fx_oracle_basic_properties():
    """Sanity of the restated DSP: the reverb is linear and has a tail, a compressor below threshold is the identity,
    the limiter bounds the output by 1."""
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(24000) * np.exp(-np.arange(24000) / 3000.0)).astype(np.float32)
    x[12000:] = 0.0
    a = fx_oracle.reverb(x, 24000, 0.5, 0.5, 0.3, 0.7, 0.8)
    b = fx_oracle.reverb(0.5 * x, 24000, 0.5, 0.5, 0.3, 0.7, 0.8)
    assert np.abs(a - 2 * b).max() < 1e-4 * np.abs(a).max()
    assert np.abs(a[14000:]).max() > 1e-3                                      # the tail rings on after the input ended
    quiet = 0.01 * x
    assert np.array_equal(fx_oracle.compressor(quiet, 24000, -5.0, 4.0, 10.0, 100.0), quiet)
    loud = fx_oracle.compressor(4 * x, 24000, -5.0, 4.0, 1.0, 100.0)
    assert np.abs(loud).max() < np.abs(4 * x).max()
    lim = fx_oracle.limiter(8 * x, 24000, -1.0)
    assert np.abs(lim).max() <= 1.0


def test_fx_normals_fifteen_at_a_time_are_the_single_draws():
    """fill_fx_normals draws a batch's compressor / limiter parameters with randn(k <= 15) per call; the reference draws
    them with one randn(1) per parameter (utils/utils.py:266-269).  Same values bit for bit, same generator state."""
    import torch
    from adt_str_b200.planner import FX_COMPRESSOR, FX_DTYPE, FX_LIMITER, fill_fx_normals, normal_draw
    rng = np.random.default_rng(0)
    fx = np.zeros(400, FX_DTYPE)
    fx["flags"] = rng.integers(0, 8, len(fx))
    got = fill_fx_normals(fx.copy(), torch.Generator().manual_seed(7))
    g = torch.Generator().manual_seed(7)
    want = fx.copy()
    for r in range(len(want)):
        f = int(want["flags"][r])
        if f & FX_COMPRESSOR:
            want["comp_threshold_db"][r] = -normal_draw(0.15, 0.5, 10, 0, g)
            want["comp_ratio"][r] = normal_draw(0.15, 0.5, 10, 1.0, g)
            want["comp_attack_ms"][r] = normal_draw(0.05, 0.1, 1000, 0, g)
            want["comp_release_ms"][r] = normal_draw(0.15, 0.2, 1000, 0, g)
        if f & FX_LIMITER:
            want["lim_threshold_db"][r] = -normal_draw(0.2, 0.4, 3, 0, g)
    assert got.tobytes() == want.tobytes()
    g2 = torch.Generator().manual_seed(7)
    fill_fx_normals(fx.copy(), g2)
    assert torch.equal(torch.randn(3, generator=g), torch.randn(3, generator=g2))     # both generators stand at the same draw
