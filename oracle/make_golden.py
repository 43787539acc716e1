#!/usr/bin/env python3
"""Freeze outputs of the UNMODIFIED reference as fixtures under tests/golden/.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every case it (1) generates a small synthetic bank and note lists, (2) runs the
reference ``SynthDrum`` (h5py/pedalboard shimmed, FX off) and the reference
``ComputeMelSpectrogram`` on the collated batch, (3) runs the NumPy oracle on the same
seeds and REFUSES to write unless it agrees with the reference (lengths equal, waveform
max-abs < 1e-6, float32 log-mel max-abs < 1e-6, float64 log-mel < 2e-5) and (4) stores inputs, reference outputs and the
oracle's per-note trace (start / length / chosen one-shots / mixup - validated by the
waveform agreement) in one .npz per case.
"""
from __future__ import annotations

import os
import random
import sys
import warnings

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from adt_str_b200.config import SETTING_1  # noqa: E402  (plain dict of config values)
from adt_str_b200.synthetic import make_bank, make_segments, make_dense_segment  # noqa: E402
from oracle import mel_oracle, ref_harness, synth_oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (config overrides, bank kwargs, n_segments, events seed, python-random seed)
    "short_24k": (dict(input_sec=0.64), dict(n_oneshots=156, max_len=3000, min_len=200, seed=10), 8, 11, 1234),
    "short_24k_tau04_adtof": (dict(input_sec=0.64, similarity_threshold=0.4, mixup_range=0.3, ADTOF_mapping=True),
                              dict(n_oneshots=156, max_len=2500, min_len=200, seed=12,
                                   groups=("gold", "90-80", "60-50", "50-40", "40-30")), 6, 13, 99),
    "default_16k": (dict(sample_rate=16000), dict(n_oneshots=78, max_len=4000, min_len=300, seed=14, sample_rate=16000),
                    2, 15, 5),
    "setting1_24k": (dict(), dict(n_oneshots=78, max_len=6000, min_len=300, seed=16), 2, 17, 42),
}


def segments_for(cfg, n, seed):
    segs = make_segments(n, seed=seed, input_sec=cfg["input_sec"], mean_events=10 if cfg["input_sec"] < 1 else 32,
                         empty_fraction=0.0)
    segs[1] = np.zeros((0, 4), np.float32)                       # an empty note list
    if cfg["ADTOF_mapping"]:                                      # notes carry ADTOF class pitches then
        for s in segs:                                            # (midi_tokenizer.py:37-40)
            s[:, 2] = [synth_oracle._CLASS[int(p)] for p in s[:, 2]]
    if n > 3:                                                     # a note whose offset pushes the end out
        segs[2] = np.concatenate([segs[2], np.array([[cfg["input_sec"] - 0.04, cfg["input_sec"] + 0.06, 38, 0]],
                                                    np.float32)])  # velocity 0: silent but extends the segment
        segs[3][:, 3] = np.minimum(segs[3][:, 3], 50)             # low max velocity -> max_volume < 1
    return segs


def main():
    if not ref_harness.available():
        raise SystemExit("reference tree not available: fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    for name, (over, bank_kw, n, ev_seed, py_seed) in CASES.items():
        cfg = dict(SETTING_1, **over)
        cfg["oneshot_path"] = f"golden_{name}"
        bank = make_bank(**{"sample_rate": cfg["sample_rate"], **bank_kw})
        nested = bank.to_nested()
        segs = segments_for(cfg, n, ev_seed)
        ref = ref_harness.make_synth(cfg, nested)
        random.seed(py_seed)
        ref_wavs = [ref(s if len(s) else []).numpy() for s in segs]
        random.seed(py_seed)
        traces, ora_wavs = [], []
        for s in segs:
            t = []
            ora_wavs.append(synth_oracle.render(s, cfg, nested, trace=t))
            traces.append(t)
        assert [len(w) for w in ref_wavs] == [len(w) for w in ora_wavs], name
        wav_err = max(float(np.abs(a - b).max()) for a, b in zip(ref_wavs, ora_wavs))
        assert wav_err < 1e-6, (name, wav_err)
        batch = synth_oracle.collate(ref_wavs)
        mel_mod = ref_harness.make_mel(cfg["sample_rate"], cfg["win_length"], cfg["time_res"], 128)
        ref_mel = mel_mod(torch.from_numpy(batch)).numpy()
        geo = (cfg["sample_rate"], cfg["win_length"], cfg["time_res"], 128)
        ora_mel = mel_oracle.logmel_direct(batch, *geo, np.float32, fb=mel_mod.compute_spec.mel_scale.fb.numpy(),
                                           window=mel_mod.compute_spec.spectrogram.window.numpy())
        mel_err = float(np.abs(ref_mel - ora_mel).max())
        assert mel_err < 1e-6, (name, mel_err)          # float32 restatement with the reference's own buffers
        truth = mel_oracle.logmel_direct(batch, *geo, np.float64)
        truth_err = float(np.abs(ref_mel - truth).max())
        assert truth_err < 2e-5, (name, truth_err)      # float64 "truth": the reference's own rounding error
        name_to_id = {nm: i for i, nm in enumerate(bank.names)}
        flat = [(si, t["start"], t["len"], name_to_id[t["main"]], name_to_id[t["sub"]], t["pitch"])
                for si, tr in enumerate(traces) for t in tr]
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            cfg_keys=np.array(sorted(k for k in cfg if k != "oneshot_path")),
            cfg_vals=np.array([repr(cfg[k]) for k in sorted(cfg) if k != "oneshot_path"]),
            py_seed=py_seed,
            bank_pcm=bank.pcm, bank_offsets=bank.offsets, bank_lengths=bank.lengths, bank_names=np.array(bank.names),
            notes=np.concatenate([s.reshape(-1, 4) for s in segs]).astype(np.float32),
            notes_count=np.array([len(s) for s in segs]),
            ref_wav=np.concatenate(ref_wavs), ref_len=np.array([len(w) for w in ref_wavs]),
            trace=np.array(flat, np.int64).reshape(-1, 6),
            trace_mixup=np.array([t["mixup"] for tr in traces for t in tr], np.float64),
            ref_mel=ref_mel,
        )
        print(f"{name}: {n} segments, lengths {[len(w) for w in ref_wavs]}, oracle-vs-reference wav {wav_err:.2e} "
              f"mel {mel_err:.2e} (vs float64 {truth_err:.2e}), mel shape {ref_mel.shape}")


def make_audio_front():
    """tests/golden/audio_front.npz: the reference's own eval / inference audio front (utils/audio_utils.py:12-23,
    inference.py:35-48,75-98) on seeded synthetic signals - inputs and outputs of the unmodified functions."""
    from oracle import audio_oracle
    au, inf = ref_harness.import_audio_front()
    rng = np.random.default_rng(77)

    def signal(ch, n, sr):  # decaying noise bursts + a low tone: broadband, non-stationary, |x| < 1
        t = np.arange(n) / sr
        env = np.exp(-8.0 * ((t * 3.0) % 1.0))
        x = 0.4 * rng.standard_normal((ch, n)) * env + 0.3 * np.sin(2 * np.pi * 110.0 * t)[None]
        return x.astype(np.float32)

    out = {}
    cases = [("stereo_44k1", 2, 44100, 24000, 115000), ("mono_48k", 1, 48000, 24000, 50001),
             ("mono_22k05_16k", 1, 22050, 16000, 30000), ("stereo_24k", 2, 24000, 24000, 62001)]
    mel24 = ref_harness.make_mel(24000, 2048, 0.01, 128)
    mel16 = ref_harness.make_mel(16000, 2048, 0.01, 128)
    names = []
    for name, ch, sr, target, n in cases:
        x = signal(ch, n, sr)
        w = torch.from_numpy(x)
        # inference.py:82-87: resample every channel, then the channel mean
        y = torchaudio.transforms.Resample(sr, target)(w) if sr != target else w
        res_per_channel = y.numpy().copy()
        if y.shape[0] > 1:
            y = y.mean(dim=0, keepdim=True)
        chunk = int(round(2.56 * target))
        chunks = inf._chunk_audio(y, chunk)
        starts = np.array([s for s, _ in chunks], np.int64)
        mat = torch.cat([c for _, c in chunks], 0)
        mel = (mel24 if target == 24000 else mel16)(mat).numpy()
        # utils/audio_utils.py:10-23: mean first, then resample, then normalize (load_and_resample / eval_dataset)
        mono_first = au.resample(w.mean(0), sr, target) if sr != target else w.mean(0)
        norm = au.normalize(mono_first).numpy()
        ora = audio_oracle.long_form_chunks(x, sr, target, 2.56)
        assert ora.shape == tuple(mat.shape) and float(np.abs(ora - mat.numpy()).max()) < 1e-6, name
        if sr != target:
            d = audio_oracle.resample_direct(x, sr, target, np.float64)
            err = float(np.abs(d - res_per_channel).max())
            assert err < 2e-6, (name, err)
        if sr != target:
            out[f"{name}/resampled"] = res_per_channel
        out.update({f"{name}/x": x, f"{name}/sr": np.array([sr, target]),
                    f"{name}/chunk_starts": starts, f"{name}/chunks": mat.numpy(), f"{name}/mel": mel,
                    f"{name}/mono_first_normalized": norm})
        names.append(name)
        print(f"audio_front {name}: {x.shape} @ {sr} -> {res_per_channel.shape} @ {target}, {len(starts)} chunks, "
              f"mel {mel.shape}")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "audio_front.npz"), **out)


def make_tokens():
    """tests/golden/tokens.npz: the reference's MidiTokenizer + collate_fn (token half) on seeded segments."""
    tok, collate = ref_harness.import_tokenizer()
    segs = make_segments(24, seed=91, empty_fraction=0.15)
    out = {"notes": np.concatenate([s.reshape(-1, 4) for s in segs]).astype(np.float32),
           "notes_count": np.array([len(s) for s in segs])}
    for adtof in (False, True):
        for vel in (False, True):
            t = tok.MidiTokenizer(tok.MidiTokenizerConfig(adtof, 1, 2, 0, 3, vel))
            toks = []
            for s in segs:
                if len(s) == 0:
                    toks.append(t.empty_adt_tokens())
                    continue
                notes = t.map_notes_to_Gm_custom(torch.from_numpy(s.copy()))
                toks.append(t.notes_to_adt_tokens(notes))
            key = f"adtof{int(adtof)}_vel{int(vel)}"
            out[f"{key}/tokens"] = torch.cat([x.double() for x in toks]).numpy()
            out[f"{key}/count"] = np.array([len(x) for x in toks])
            out[f"{key}/float"] = np.array([x.is_floating_point() for x in toks])
            dec = [t.decode(x.numpy()) for x in toks]
            out[f"{key}/decoded"] = np.concatenate([d.numpy().reshape(-1, 4) for d in dec]).astype(np.float64)
            out[f"{key}/decoded_count"] = np.array([len(d) for d in dec])
            if collate is not None:
                c = collate([(torch.zeros(4), x.tolist()) for x in toks])
                out[f"{key}/collated"] = c["tokens"].numpy()
                out[f"{key}/collated_lengths"] = c["token_lengths"].numpy()
    np.savez_compressed(os.path.join(OUT, "tokens.npz"), **out)
    print("tokens: 24 segments x 4 tokenizer configurations")


FX_CASES = {
    # name: (config overrides, bank kwargs, n_segments, events seed, python-random seed, torch seed)
    "fx_24k": (dict(use_fx_prob=0.75, use_reverb_prob=0.6, use_compression_prob=0.6, use_limiter_prob=0.6),
               dict(n_oneshots=156, max_len=9000, min_len=300, seed=20), 10, 21, 314, 2718),
    "fx_16k_short": (dict(sample_rate=16000, input_sec=0.64, use_fx_prob=1.0, use_reverb_prob=1.0,
                          use_compression_prob=1.0, use_limiter_prob=1.0),
                     dict(n_oneshots=78, max_len=3000, min_len=200, seed=22, sample_rate=16000), 4, 23, 7, 11),
}


def make_fx():
    """tests/golden/fx_*.npz: the UNMODIFIED reference rendering with the FX chain ON.  pedalboard is not installed, so
    the reference's ``Pedalboard / Reverb / Compressor / Limiter`` names are pointed at the recording stand-in whose
    DSP is ``oracle/fx_oracle.c`` (``ref_harness.enable_fx_stand_in``): which plugins are built, with which parameters,
    from which draws, and where the chain sits are the reference's own; the DSP arithmetic is the restatement's
    (parity with pedalboard itself: unpinned).  The oracle (same seeds) must agree before anything is written."""
    fx = ref_harness.enable_fx_stand_in()
    os.makedirs(OUT, exist_ok=True)
    for name, (over, bank_kw, n, ev_seed, py_seed, torch_seed) in FX_CASES.items():
        cfg = dict(SETTING_1, **over)
        cfg["oneshot_path"] = f"golden_{name}"
        bank = make_bank(**{"sample_rate": cfg["sample_rate"], **bank_kw})
        nested = bank.to_nested()
        segs = segments_for(cfg, n, ev_seed)
        ref = ref_harness.make_synth(cfg, nested)
        random.seed(py_seed)
        torch.manual_seed(torch_seed)
        ref_wavs, plugins = [], []
        for s in segs:
            fx.CALLS.clear()
            ref_wavs.append(ref(s if len(s) else []).numpy())
            plugins.append(sum({"Reverb": 1, "Compressor": 2, "Limiter": 4}[c[0]] for c in fx.CALLS))
        random.seed(py_seed)
        torch.manual_seed(torch_seed)
        traces, ora_wavs, boards = [], [], []
        for s in segs:
            t = []
            ora_wavs.append(synth_oracle.render(s, cfg, nested, trace=t, boards=boards) if len(s) else
                            synth_oracle.render(s, cfg, nested))
            if not len(s):
                boards.append(None)
            traces.append(t)
        assert [len(w) for w in ref_wavs] == [len(w) for w in ora_wavs], name
        wav_err = max(float(np.abs(a - b).max()) for a, b in zip(ref_wavs, ora_wavs))
        assert wav_err < 1e-6, (name, wav_err)
        flags = [0 if b is None else sum({"Reverb": 1, "Compressor": 2, "Limiter": 4}[p[0]] for p in b) for b in boards]
        assert flags == plugins, (flags, plugins)
        batch = synth_oracle.collate(ref_wavs)
        mel_mod = ref_harness.make_mel(cfg["sample_rate"], cfg["win_length"], cfg["time_res"], 128)
        ref_mel = mel_mod(torch.from_numpy(batch)).numpy()
        name_to_id = {nm: i for i, nm in enumerate(bank.names)}
        flat = [(si, t["start"], t["len"], name_to_id[t["main"]], name_to_id[t["sub"]], t["pitch"])
                for si, tr in enumerate(traces) for t in tr]
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            cfg_keys=np.array(sorted(k for k in cfg if k != "oneshot_path")),
            cfg_vals=np.array([repr(cfg[k]) for k in sorted(cfg) if k != "oneshot_path"]),
            py_seed=py_seed, torch_seed=torch_seed,
            bank_pcm=bank.pcm, bank_offsets=bank.offsets, bank_lengths=bank.lengths, bank_names=np.array(bank.names),
            notes=np.concatenate([s.reshape(-1, 4) for s in segs]).astype(np.float32),
            notes_count=np.array([len(s) for s in segs]),
            ref_wav=np.concatenate(ref_wavs), ref_len=np.array([len(w) for w in ref_wavs]),
            trace=np.array(flat, np.int64).reshape(-1, 6),
            trace_mixup=np.array([t["mixup"] for tr in traces for t in tr], np.float64),
            ref_mel=ref_mel, fx_flags=np.array(plugins, np.int32),
        )
        print(f"{name}: {n} segments, plugins per segment {plugins}, oracle-vs-reference wav {wav_err:.2e}")


if __name__ == "__main__":
    import sys
    if "--fx-only" in sys.argv:
        make_fx()
    elif "--audio-front-only" in sys.argv:
        make_audio_front()
    elif "--tokens-only" in sys.argv:
        make_tokens()
    else:
        main()
        make_audio_front()
        make_tokens()
        make_fx()
