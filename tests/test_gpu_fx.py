"""FX chain on the GPU (csrc/fx.cu) against fixtures frozen from the UNMODIFIED reference running with a stand-in
pedalboard whose DSP is oracle/fx_oracle.c (oracle/make_golden.py --fx-only), and against the oracle run on this box.

Tolerance: the fixtures' DSP is sequential float32 (JUCE's evaluation order); the kernels solve the comb filters'
one-pole recurrence with a shuffle scan (re-associated sums) and use CUDA's powf, so waveforms agree to float32
rounding accumulated over the recursions - max-abs <= 2e-5 on rows normalised to <= 1 (measured ~2e-6), which is the
north-star waveform tolerance (1e-5) doubled for the feedback loops.  Parity with pedalboard itself is unpinned.
"""
import random

import numpy as np
import pytest
import torch

from conftest import Golden

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _synth(g, **over):
    import dataclasses
    from adt_str_b200 import SynthDrum
    cfg = dataclasses.replace(g.config(), **over) if over else g.config()
    return SynthDrum(cfg, bank=g.bank, device=torch.device("cuda", 0))


@pytest.mark.parametrize("name", ["fx_24k", "fx_16k_short"])
@pytest.mark.parametrize("native", [True, False])
def test_render_with_fx_matches_the_reference_fixture(name, native):
    import os
    from conftest import GOLDEN_DIR
    g = Golden(name)
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    synth = _synth(g)
    random.seed(g.py_seed)
    torch.manual_seed(int(z["torch_seed"]))
    plan = synth.plan(g.segments, native=native)
    flags = np.zeros(len(g.segments), np.int32)
    if plan.fx is not None:
        flags[plan.fx["seg"]] = plan.fx["flags"]
    assert flags.tolist() == z["fx_flags"].tolist()
    wav = synth.render_plan(plan).cpu().numpy()
    assert plan.wave_lengths.tolist() == g.ref_len.tolist()
    worst = 0.0
    for s, ref in enumerate(g.ref_wavs):
        got = wav[s, : len(ref)]
        assert not wav[s, len(ref):].any()
        worst = max(worst, float(np.abs(got - ref).max()))
    assert worst <= TOL, worst


def test_fx_rows_in_batches_and_chunks_equal_single_calls():
    """FX rows inside a multi-batch, multi-chunk plan (pipeline over the internal streams) come out exactly as in
    single-segment calls with the same parameters, rows without FX are bit-identical to an FX-free render, and the
    log-mel of the fused call sees the FX."""
    from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum
    from adt_str_b200.config import setting_1
    from adt_str_b200.synthetic import make_bank, make_segments
    bank = make_bank(390, 24000, seed=41, min_len=400, max_len=20000)
    dev = torch.device("cuda", 0)
    cfg_fx = setting_1(use_fx_prob=0.5, use_reverb_prob=0.5, use_compression_prob=0.5, use_limiter_prob=0.5)
    segs = make_segments(48, seed=42, empty_fraction=0.1)
    batches = [segs[0:7], segs[7:20], segs[20:21], segs[21:40], segs[40:48]]
    synth = SynthDrum(cfg_fx, bank=bank, device=dev)
    fe = FrontEnd(synth, ComputeMelSpectrogram(24000, 2048, 0.01, 128))
    torch.manual_seed(3)
    plan = fe.plan_batches(batches, random.Random(8), 2)
    assert plan.fx is not None and 8 <= len(plan.fx) <= 40 and plan.chunks["fx_row"][-1] == len(plan.fx)
    wav, feat = fe.run_plan(plan)
    wav2, feat2 = fe.run_plan(plan)
    torch.cuda.synchronize()
    assert torch.equal(wav, wav2) and torch.equal(feat, feat2)          # deterministic
    # segment by segment with the same `random` stream (so the same coins): rows without a record are rendered without
    # FX, rows with one get the board the big plan drew (torch's stream is consumed per plan, so it is copied over)
    has_fx = np.zeros(plan.n_seg, bool)
    has_fx[plan.fx["seg"]] = True
    rng = random.Random(8)
    rows = []
    for s, notes in enumerate(n for b in batches for n in b):
        p1 = synth.plan([notes], rng, ld_wav=plan.ld_wav, generator=torch.Generator().manual_seed(0))
        assert (p1.fx is not None) == bool(has_fx[s])
        if has_fx[s]:
            p1.fx[:] = plan.fx[np.searchsorted(plan.fx["seg"], s)]
            p1.fx["seg"] = 0
        rows.append(synth.render_plan(p1)[0].clone())
    torch.cuda.synchronize()
    for s, row in enumerate(rows):
        assert torch.equal(torch.nan_to_num(wav[s], nan=-7.0), torch.nan_to_num(row[: wav.shape[1]], nan=-7.0)), s
    # log-mel of the fused call = log-mel of the FX'd waveforms
    mel = fe.mel
    for (w_b, f_b) in plan.split(wav, feat):
        assert torch.equal(mel(w_b), f_b)


def test_fx_kernels_against_the_oracle_on_this_box():
    """Reverb / compressor / limiter one at a time and chained, parameters at the edges of the reference's ranges,
    against oracle/fx_oracle.c compiled here."""
    from adt_str_b200.planner import FX_DTYPE
    from adt_str_b200.synthetic import make_bank
    from oracle import fx_oracle, synth_oracle
    from adt_str_b200 import SynthDrum
    from adt_str_b200.config import SETTING_1, setting_1
    bank = make_bank(156, 24000, seed=51, min_len=2000, max_len=30000)
    synth = SynthDrum(setting_1(), bank=bank, device=torch.device("cuda", 0))
    notes = [[0.05, 0.15, 36, 120], [0.4, 0.5, 38, 100], [0.41, 0.51, 42, 90], [1.2, 1.3, 46, 127], [2.3, 2.4, 36, 60]]
    boards = [
        dict(flags=1, room_size=0.2, damping=0.2, wet_level=0.1, dry_level=0.9, width=0.6),
        dict(flags=1, room_size=0.8, damping=0.8, wet_level=0.4, dry_level=0.6, width=1.0),
        dict(flags=2, comp_threshold_db=-10.0, comp_ratio=10.0, comp_attack_ms=0.0, comp_release_ms=1000.0),
        dict(flags=2, comp_threshold_db=-2.5, comp_ratio=1.0, comp_attack_ms=120.0, comp_release_ms=50.0),
        dict(flags=4, lim_threshold_db=-3.0),
        dict(flags=4, lim_threshold_db=-0.0),
        dict(flags=7, room_size=0.5, damping=0.4, wet_level=0.3, dry_level=0.7, width=0.8, comp_threshold_db=-5.0,
             comp_ratio=4.0, comp_attack_ms=80.0, comp_release_ms=200.0, lim_threshold_db=-1.2),
        dict(flags=0),
    ]
    nested = bank.to_nested()
    for b in boards:
        plan = synth.plan([notes], random.Random(1))
        fx = np.zeros(1, FX_DTYPE)
        for k, v in b.items():
            fx[k] = v
        plan.fx = fx
        got = synth.render_plan(plan)[0, : int(plan.wave_lengths[0])].cpu().numpy()
        raw, vol = synth_oracle.render(notes, dict(SETTING_1), nested, rng=random.Random(1), raw=True)
        board = []
        if b["flags"] & 1:
            board.append(("Reverb", {k: b[k] for k in ("room_size", "damping", "wet_level", "dry_level", "width")}))
        if b["flags"] & 2:
            board.append(("Compressor", dict(threshold_db=b["comp_threshold_db"], ratio=b["comp_ratio"],
                                             attack_ms=b["comp_attack_ms"], release_ms=b["comp_release_ms"])))
        if b["flags"] & 4:
            board.append(("Limiter", dict(threshold_db=b["lim_threshold_db"])))
        want = synth_oracle.apply_board(raw, board, 24000)
        want = want / np.abs(want).max() * vol
        assert float(np.abs(got - want).max()) <= TOL, (b, float(np.abs(got - want).max()))


def test_pedalboard_backend_routes_fx_rows_through_the_host_library():
    """``SynthDrum.fx_backend = "pedalboard"``: raw rows (ADTFE_SEG_RAW) -> host -> the plugins the reference builds ->
    normalise -> back.  With the stand-in pedalboard (same DSP as the fixtures) the result is the fixture's."""
    import os
    import sys
    from conftest import GOLDEN_DIR
    from oracle import fx_oracle
    saved = sys.modules.get("pedalboard")
    fx_oracle.install_pedalboard_stand_in()
    try:
        g = Golden("fx_24k")
        z = np.load(os.path.join(GOLDEN_DIR, "fx_24k.npz"))
        synth = _synth(g)
        synth.fx_backend = "pedalboard"
        random.seed(g.py_seed)
        torch.manual_seed(int(z["torch_seed"]))
        wav, lengths = synth.render_batch(g.segments)
        wav = wav.cpu().numpy()
        assert lengths.tolist() == g.ref_len.tolist()
        for s, ref in enumerate(g.ref_wavs):
            assert float(np.abs(wav[s, : len(ref)] - ref).max()) <= 1e-5, s   # rows without FX: the GPU mix; with FX: the same DSP
            assert not wav[s, len(ref):].any()
    finally:
        if saved is None:
            sys.modules.pop("pedalboard", None)
        else:
            sys.modules["pedalboard"] = saved
