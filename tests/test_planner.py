"""Host logic: planner indices bit-exact against the fixtures, bucketing, errors, bank, config."""
import random

import numpy as np
import pytest

from adt_str_b200 import planner
from adt_str_b200.bank import OneShotBank
from adt_str_b200.config import SETTING_1, SynthDrumConfig, deep_merge, setting_1, synth_config_from_sections
from adt_str_b200.mapping import instrument_gain
from adt_str_b200.synthetic import make_bank, make_dense_segment, make_long_form, make_segments
from oracle import synth_oracle


def _track_order(rows):
    """note-order trace rows of one segment -> the planner's track order."""
    rank = {}
    for r in rows:
        rank.setdefault(int(r[5]), len(rank))
    return sorted(range(len(rows)), key=lambda i: (rank[int(rows[i][5])], i))


def test_planner_indices_bit_exact_vs_reference_trace(golden):
    cfg = golden.config()
    random.seed(golden.py_seed)
    plan = planner.plan_batch(golden.segments, cfg, golden.bank)
    assert plan.wave_lengths.tolist() == golden.ref_len.tolist()
    assert plan.segments["len"].tolist() == golden.ref_len.tolist()
    assert plan.ld_wav % 4 == 0 and plan.ld_wav >= golden.ref_len.max()
    for s in range(plan.n_seg):
        sel = golden.trace[:, 0] == s
        rows, mix = golden.trace[sel], golden.trace_mixup[sel]
        ev = plan.events[plan.events["seg"] == s]
        order = _track_order(rows)
        assert len(ev) == len(rows)
        for e, i in zip(ev, order):
            assert (e["start"], e["len"], e["main_id"], e["sub_id"]) == tuple(rows[i][1:5])
            assert e["cb"] == np.float32(mix[i]) and e["ca"] == np.float32(1.0 - mix[i])
        assert (plan.segments["flags"][s] == planner.SEG_EMPTY) == (len(rows) == 0)


def test_tile_buckets_equal_plain_loop(golden):
    random.seed(golden.py_seed)
    plan = planner.plan_batch(golden.segments, golden.config(), golden.bank)
    ev = plan.events
    n_tiles = plan.n_seg * plan.tiles_per_seg
    start = ev["seg"].astype(np.int64) * plan.tiles_per_seg * planner.TILE + ev["start"]
    want = synth_oracle.bucket_tiles(start.tolist(), ev["len"].tolist(), planner.TILE, n_tiles)
    assert plan.tile_ptr[0] == 0 and plan.tile_ptr[-1] == len(plan.tile_events)
    for t in range(n_tiles):
        assert plan.tile_events[plan.tile_ptr[t]: plan.tile_ptr[t + 1]].tolist() == want[t]


def test_groups_cover_events_and_share_oneshots():
    bank = make_bank(156, max_len=3000)
    random.seed(3)
    plan = planner.plan_batch(make_segments(12) + [make_dense_segment()], setting_1(), bank)
    gp = plan.group_ptr
    assert gp[0] == 0 and gp[-1] == plan.n_events and (np.diff(gp) > 0).all()
    for a, b in zip(gp[:-1], gp[1:]):
        e = plan.events[a:b]
        assert len({(x["seg"], x["main_id"], x["sub_id"]) for x in e}) == 1
        assert (plan.mix_len[a:b] == max(bank.lengths[e["main_id"][0]], bank.lengths[e["sub_id"][0]])).all()
    assert (plan.events["len"] <= plan.mix_len).all() and (plan.events["len"] >= 0).all()
    ends = plan.events["start"] + plan.events["len"]
    assert (ends <= plan.segments["len"][plan.events["seg"]]).all()


def test_rng_stream_advances_like_the_reference_call_order():
    bank = make_bank(156, max_len=2000)
    notes = [[0.1, 0.2, 36, 100], [0.2, 0.3, 38, 90], [0.3, 0.4, 36, 80]]
    random.seed(5)
    planner.plan_segment(notes, setting_1(), bank)
    got = random.random()
    random.seed(5)
    for _pitch in (36, 38):          # kick: group, key, group, key, then its mixup ... (synthetiser.py:274-281, 217)
        pass
    draws = []
    random.seed(5)
    g = planner.similarity_groups(0.8)
    for new_instrument in (True, True, False):
        if new_instrument:
            for _ in range(2):
                grp = random.choice(g)
                random.choice(range(2))      # 156 one-shots / 78 cells = 2 per (pitch, group)
        draws.append(random.uniform(0, 0.8))
    random.random()                          # FX coin
    assert got == random.random()


def test_invalid_notes_raise_like_the_reference():
    bank = make_bank(78, min_len=100, max_len=1000)
    with pytest.raises(ValueError, match="Invalid note"):
        planner.plan_segment([[0.1, 0.2, 62, 100]], setting_1(), bank)      # pitch > 61
    with pytest.raises(ValueError, match="Invalid note"):
        planner.plan_segment([[0.3, 0.2, 36, 100]], setting_1(), bank)      # offset < onset
    with pytest.raises(IndexError):                                         # pitch 61 has no one-shots: choice([])
        planner.plan_segment([[0.1, 0.2, 61, 100]], setting_1(), bank)
    fx = planner.plan_segment([[0.1, 0.2, 36, 100]], setting_1(use_fx_prob=1.0, use_reverb_prob=1.0), bank).fx
    assert fx is not None and int(fx["flags"][0]) & planner.FX_REVERB      # an FX coin hit is planned, not refused
    with pytest.raises(KeyError):                                           # ADTOF mode expects class pitches
        planner.plan_segment([[0.1, 0.2, 36, 100]], setting_1(ADTOF_mapping=True), bank)


def test_empty_and_silent_segments():
    bank = make_bank(78, min_len=100, max_len=1000)
    p = planner.plan_segment([], setting_1(), bank)
    assert (p.wave_length, p.flags, len(p.events)) == (61440, planner.SEG_EMPTY, 0)
    p = planner.plan_segment(np.zeros((0, 4), np.float32), config_16k(), bank)
    assert p.wave_length == 40960
    p = planner.plan_segment([[0.1, 0.2, 36, 0]], setting_1(), bank)         # velocity 0 -> volume 0 -> NaN mix
    assert p.flags == planner.SEG_NORMALISE and p.max_volume == 0.0 and p.events["gain"][0] == 0.0


def config_16k():
    return SynthDrumConfig(**dict(SETTING_1, sample_rate=16000))


def test_velocity_curve_and_gains():
    v = planner.velocity_to_volume(np.array([0, 1, 64, 127, 200], np.float32))
    assert v[0] == 0 and abs(v[3] - 1.0) < 1e-6 and v[4] == v[3] and 0.1 < v[1] < v[2] < 1
    assert instrument_gain(36, False) == 1.0 and instrument_gain(42, False) == 0.7 and instrument_gain(46, False) == 0.7
    assert instrument_gain(48, True) == 0.7
    assert planner.similarity_groups(0.8) == ["gold", "100-90", "90-80"]
    assert len(planner.similarity_groups(0.4)) == 7 and planner.similarity_groups(1.1) == []


def test_bank_roundtrip_and_name_order(tmp_path):
    nested = {"36": {"gold": {"b": np.ones(5, np.float32), "a": np.arange(7, dtype=np.float32)}},
              "38": {"90-80": {"z": np.zeros(3, np.float32)}}}
    bank = OneShotBank.from_nested(nested)
    assert bank.names == ["36/gold/a", "36/gold/b", "38/90-80/z"]           # sorted like h5py keys()
    assert bank.group_range(36, "gold") == (0, 2) and not bank.has_group(36, "90-80")
    assert (bank.offsets % 32 == 0).all() and bank.oneshot(0).tolist() == list(range(7))
    path = str(tmp_path / "bank.npz")
    bank.save(path)
    back = OneShotBank.load(path)
    assert back.names == bank.names and np.array_equal(back.pcm, bank.pcm) and back.index == bank.index
    assert set(back.to_nested()["36"]["gold"]) == {"a", "b"}


def test_config_sections_and_merge():
    cfg = {"shared": dict(input_sec=2.56, time_res=0.01, win_length=2048, sample_rate=24000),
           "synthetiser": {k: SETTING_1[k] for k in ("oneshot_path", "similarity_threshold", "max_hat_std_velocity",
                                                     "max_hat_mean_velocity", "max_cymbals_std_velocity",
                                                     "max_cymbals_mean_velocity", "mixup_range", "use_fx_prob",
                                                     "use_reverb_prob", "use_limiter_prob", "use_compression_prob")},
           "tokenizer": {"ADTOF_mapping": False}}
    c = synth_config_from_sections(cfg)
    assert c == setting_1()
    assert deep_merge({"a": {"b": 1, "c": 2}}, {"a": {"b": 3}}) == {"a": {"b": 3, "c": 2}}


def test_bytes_alg_counts_every_term():
    bank = make_bank(156, max_len=3000)
    random.seed(9)
    plan = planner.plan_batch(make_segments(4, empty_fraction=0.0), setting_1(), bank)
    ev = plan.events
    want = 0
    for s in range(plan.n_seg):
        e = ev[ev["seg"] == s]
        firsts = {}
        for x in e:
            for u in (int(x["main_id"]), int(x["sub_id"])):
                firsts[u] = min(firsts.get(u, 1 << 60), int(x["start"]))
        want += sum(4 * max(0, min(int(bank.lengths[u]), int(plan.segments["len"][s]) - st)) for u, st in firsts.items())
    assert plan.bank_bytes(bank) == want
    assert plan.bytes_alg(bank, 246, 128) == want + 4 * int(plan.segments["len"].sum()) + 4 * 246 * 128 * 4 + 32 * len(ev)


def test_long_form_plan_is_one_segment():
    bank = make_bank(156, max_len=3000)
    random.seed(2)
    notes = make_long_form(30.0)
    plan = planner.plan_batch([notes], setting_1(), bank)
    assert plan.n_seg == 1 and plan.wave_lengths[0] >= 29 * 24000 and plan.tiles_per_seg == -(-plan.ld_wav // 2048)


# ------------------------------------------------------------------ C++ planner (csrc/planner.cpp)
def _plans_equal(a, b):
    for f in ("start", "len", "main_id", "sub_id", "ca", "cb", "seg"):
        assert np.array_equal(a.events[f], b.events[f]), f
    assert np.allclose(a.events["gain"], b.events["gain"], rtol=3e-7, atol=0)        # powf vs numpy: 1 ulp
    for f in ("mix_len", "group_ptr", "tile_ptr", "tile_events", "peak_work", "wave_lengths"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for f in ("len", "flags", "first_event"):
        assert np.array_equal(a.segments[f], b.segments[f]), f
    assert np.allclose(a.segments["max_volume"], b.segments["max_volume"], rtol=3e-7)
    assert (a.n_seg, a.ld_wav, a.tiles_per_seg) == (b.n_seg, b.ld_wav, b.tiles_per_seg)


def test_native_planner_equals_python_planner_and_rng_stream(golden):
    from adt_str_b200.native_planner import NativePlanner
    cfg = golden.config()
    random.seed(golden.py_seed)
    want = planner.plan_batch(golden.segments, cfg, golden.bank)
    state = random.getstate()
    random.seed(golden.py_seed)
    got = NativePlanner(cfg, golden.bank).plan_batch(golden.segments)
    assert random.getstate() == state                      # same MT19937 stream, draw for draw
    _plans_equal(got, want)
    assert got.wave_lengths.tolist() == golden.ref_len.tolist()


def test_native_planner_large_batch_and_private_rng():
    from adt_str_b200.native_planner import NativePlanner
    bank = make_bank(624, max_len=30000)
    segs = make_segments(300, seed=8) + [make_dense_segment(), make_long_form(20.0)]
    cfg = setting_1(similarity_threshold=0.4)
    r1, r2 = random.Random(5), random.Random(5)
    want = planner.plan_batch(segs, cfg, bank, rng=r1)
    got = NativePlanner(cfg, bank).plan_batch(segs, rng=r2)
    assert r1.getstate() == r2.getstate()
    _plans_equal(got, want)


def test_native_planner_errors_and_fallback():
    from adt_str_b200.native_planner import NativePlanner
    bank = make_bank(78, min_len=100, max_len=1000)
    npl = NativePlanner(setting_1(), bank)
    ok = np.array([[0.1, 0.2, 36, 100]], np.float32)
    with pytest.raises(ValueError, match="Invalid note"):
        npl.plan_batch([ok, np.array([[0.1, 0.2, 62, 100]], np.float32)])
    with pytest.raises(ValueError, match="Invalid note"):
        npl.plan_batch([np.array([[0.3, 0.2, 36, 100]], np.float32)])
    with pytest.raises(IndexError):
        npl.plan_batch([np.array([[0.1, 0.2, 61, 100]], np.float32)])
    with pytest.raises(KeyError):
        NativePlanner(setting_1(ADTOF_mapping=True), bank).plan_batch([ok])
    fx = NativePlanner(setting_1(use_fx_prob=1.0, use_limiter_prob=1.0), bank).plan_batch([ok, ok]).fx
    assert len(fx) == 2 and fx["seg"].tolist() == [0, 1] and (fx["flags"] & planner.FX_LIMITER).all()
    assert np.isfinite(fx["lim_threshold_db"]).all()          # the torch draws were filled in
    # float64 notes take the Python path (float64 index arithmetic, like torch.tensor(float64 array))
    random.seed(1)
    p64 = npl.plan_batch([np.array([[0.5, 2.5, 36.0, 100.0]], np.float64)])
    random.seed(1)
    p32 = npl.plan_batch([[[0.5, 2.5, 36.0, 100.0]]])
    assert (int(p64.wave_lengths[0]), int(p32.wave_lengths[0])) == (62400, 62399)
    assert npl.plan_batch([]).n_seg == 0


def test_batches_in_one_plan_match_plans_one_by_one():
    """plan_batches = the same events / indices as planning batch by batch with the same RNG
    stream, plus per-batch widths, frame counts, ragged log-mel rows and render chunks."""
    from adt_str_b200 import ComputeMelSpectrogram, SynthDrum
    from adt_str_b200.config import setting_1
    from adt_str_b200.synthetic import make_bank, make_segments
    bank = make_bank(260, 24000, seed=4)
    segs = make_segments(14, seed=8, empty_fraction=0.2)
    batches = [segs[:5], segs[5:6], segs[6:11], segs[11:]]
    synth = SynthDrum(setting_1(), bank=bank)
    mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    rng = random.Random(77)
    big = synth.plan_batches(batches, mel.n_frames, rng)
    state_after = rng.getstate()
    rng = random.Random(77)
    singles = [synth.plan(b, rng) for b in batches]
    assert rng.getstate() == state_after
    assert big.n_seg == 14 and big.ld_wav == max(p.ld_wav for p in singles)
    e0 = r0 = pw0 = s0 = 0
    for b, p in enumerate(singles):
        ev = big.events[e0:e0 + p.n_events]
        for f in ("start", "len", "main_id", "sub_id", "ca", "cb", "gain"):
            assert np.array_equal(ev[f], p.events[f])
        assert np.array_equal(ev["seg"], p.events["seg"] + s0)
        assert np.array_equal(big.wave_lengths[s0:s0 + p.n_seg], p.wave_lengths)
        assert big.batch_samples[b] == p.wave_lengths.max()
        t = mel.n_frames(int(p.wave_lengths.max()))
        assert big.batch_frames[b] == t
        rows = big.mel_rows[s0:s0 + p.n_seg]
        assert np.array_equal(rows["count"], np.full(p.n_seg, t))
        assert np.array_equal(rows["out_row"], r0 + t * np.arange(p.n_seg))
        assert tuple(big.chunks[b]) == (s0, e0, pw0, 0)            # (seg, event, peak_work, fx_row): no FX here
        pw = big.peak_work[pw0:pw0 + len(p.peak_work)]
        assert np.array_equal(pw["first_event"], p.peak_work["first_event"] + e0)
        e0 += p.n_events; r0 += t * p.n_seg; pw0 += len(p.peak_work); s0 += p.n_seg
    assert tuple(big.chunks[-1]) == (14, big.n_events, len(big.peak_work), 0) and big.mel_total_rows == r0
    with pytest.raises(ValueError):
        big.set_batches([7, 6], mel.n_frames)


def test_native_group_pack_equals_python_set_batches_and_pack():
    """adtfe_planner_pack_batches (plan -> plan blob without the interpreter) writes byte for byte what
    RenderPlan.set_batches + the blob layout give for the same RNG stream; the MT19937 state it advances in place is
    the state random.Random ends in."""
    import ctypes as C
    import random
    from adt_str_b200 import _lib
    from adt_str_b200.config import setting_1
    from adt_str_b200.mel import ComputeMelSpectrogram
    from adt_str_b200.native_planner import NativePlanner
    from adt_str_b200.planner import CHUNK_DTYPE
    from adt_str_b200.synthetic import make_bank, make_segments
    bank = make_bank(390, 24000, seed=31)
    segs = make_segments(70, seed=32, empty_fraction=0.1)
    cuts = [0, 9, 10, 26, 33, 41, 58, 70]
    group = [segs[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    import torch
    from adt_str_b200.planner import FX_DTYPE, fill_fx_normals
    lib = _lib.load()
    for cfg, chunk_batches in [(setting_1(), 1), (setting_1(), 2), (setting_1(), 3), (setting_1(), 100),
                               (setting_1(use_fx_prob=0.5), 2), (setting_1(use_fx_prob=1.0, use_reverb_prob=1.0), 3)]:
        planner = NativePlanner(cfg, bank)
        rng = random.Random(555)
        torch.manual_seed(9)
        want = planner.plan_batch([n for b in group for n in b], rng).set_batches([len(b) for b in group], mel.n_frames,
                                                                                 chunk_batches)
        mt = np.array(random.Random(555).getstate()[1], np.uint32)
        counts = planner.plan_group(group, mt)
        assert tuple(mt.tolist()) == rng.getstate()[1]
        assert int(counts[0]) == want.n_events and int(counts[6]) == want.ld_wav
        assert int(counts[8]) == (0 if want.fx is None else len(want.fx)) and (cfg.use_fx_prob == 0) == (want.fx is None)
        rc, shape, need, chunks, width, frames, fx_off = planner.pack_group([len(b) for b in group], 240,
                                                                            mel.window_pad_idxs, chunk_batches, None, 0)
        assert rc == -3 and need > 0                               # size query
        blob = np.full(need + 64, 0xAB, np.uint8)
        rc, shape, need2, chunks, width, frames, fx_off = planner.pack_group([len(b) for b in group], 240,
                                                                             mel.window_pad_idxs, chunk_batches,
                                                                             blob.ctypes.data, blob.size)
        assert rc == 0 and need2 == need
        assert width.tolist() == want.batch_samples.tolist() and frames.tolist() == want.batch_frames.tolist()
        assert shape.mel_total_rows == want.mel_total_rows and shape.mel_max_count == int(want.batch_frames.max())
        assert shape.n_chunks == len(want.chunks) - 1 and shape.sample_rate == 24000
        assert shape.n_tile_events == len(want.tile_events)
        assert chunks.view(CHUNK_DTYPE)[: shape.n_chunks + 1].tolist() == want.chunks.tolist()
        off = (C.c_size_t * 7)()
        fixed = C.c_size_t()
        assert lib.adtfe_plan_blob_layout(C.byref(shape), C.byref(off), C.byref(fixed)) == 0
        fx = want.fx if want.fx is not None else np.zeros(0, FX_DTYPE)
        if len(fx):   # the blob's FX records wait for the torch draws: the same generator state fills them identically
            assert off[5] == fx_off
            torch.manual_seed(9)
            fill_fx_normals(blob[fx_off: fx_off + 48 * len(fx)].view(FX_DTYPE))
            assert want.chunks["fx_row"][-1] == len(fx) and (np.diff(want.chunks["fx_row"]) >= 0).all()
        for o, arr in zip(off, (want.events, want.segments, want.tile_ptr, want.peak_work, want.mel_rows, fx,
                                want.tile_events)):
            raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
            assert np.array_equal(blob[o: o + raw.size], raw)
        assert (blob[need:] == 0xAB).all()                         # nothing written past the blob
    assert planner.plan_group([[np.zeros((1, 4), np.float64)]], mt) is None      # not float32: general path
    with pytest.raises(ValueError):
        planner.plan_group([[np.array([[0.5, 0.4, 36.0, 100.0]], np.float32)]], mt)   # offset < onset


def test_bank_from_hdf5_walks_the_reference_layout():
    """OneShotBank.from_hdf5 over the layout convert_augmented_to_hdf5.py:70-138 writes (<pitch>/<bin>/<name>
    datasets plus a flat `index` group), through the dict-backed h5py stand-in of the oracle harness (h5py itself is
    not installed in this image): the index group is skipped, names come out sorted like h5py's keys()."""
    from oracle import ref_harness
    ref_harness._install_shims()
    rng = np.random.default_rng(3)
    nested = {"36": {"gold": {"kick_b": rng.standard_normal(50).astype(np.float32),
                              "kick_a": rng.standard_normal(33).astype(np.float32)},
                     "100-90": {"k": rng.standard_normal(7).astype(np.float32)}},
              "42": {"90-80": {"hat": rng.standard_normal(19).astype(np.float32)}}}
    on_disk = dict(nested, index={"paths": np.zeros(4, np.float32), "labels": np.zeros(4, np.float32)})
    ref_harness._BANKS["layout_test@24000.hdf5"] = on_disk
    bank = OneShotBank.from_hdf5("layout_test@24000.hdf5")
    want = OneShotBank.from_nested(nested)
    assert bank.names == want.names == ["36/100-90/k", "36/gold/kick_a", "36/gold/kick_b", "42/90-80/hat"]
    assert np.array_equal(bank.pcm, want.pcm) and bank.index == want.index
    assert np.array_equal(bank.oneshot(1), nested["36"]["gold"]["kick_a"])


def test_bank_converter_entry_point(tmp_path, capsys):
    """``python -m adt_str_b200.bank in.hdf5 out.npz``: the converter a maintainer runs where the HDF5 bank lives."""
    from adt_str_b200 import bank as bank_mod
    from oracle import ref_harness
    ref_harness._install_shims()
    rng = np.random.default_rng(4)
    ref_harness._BANKS["cli_test@24000.hdf5"] = {
        "38": {"gold": {"s1": rng.standard_normal(41).astype(np.float32)}},
        "index": {"paths": np.zeros(1, np.float32)}}
    out = str(tmp_path / "cli_test@24000.npz")
    assert bank_mod._main(["cli_test@24000.hdf5", out]) == 0
    assert "1 one-shots" in capsys.readouterr().out
    back = OneShotBank.load(out)
    assert back.names == ["38/gold/s1"] and back.group_range(38, "gold") == (0, 1) and int(back.lengths[0]) == 41
