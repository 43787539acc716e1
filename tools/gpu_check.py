#!/usr/bin/env python3
"""Quick on-GPU sanity check: parity vs the CPU oracle + rough timings (not a bench)."""
import os, sys, time, random, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adt_str_b200.synthetic import make_bank, make_segments
from adt_str_b200.config import SETTING_1, setting_1
from adt_str_b200.synthetiser import SynthDrum
from adt_str_b200.mel import ComputeMelSpectrogram
from adt_str_b200.frontend import FrontEnd
from oracle import synth_oracle as so, mel_oracle as mo

def main():
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    bank = make_bank(624, max_len=20000)
    nested = bank.to_nested()
    cfg = setting_1()
    synth = SynthDrum(cfg, bank=bank)
    mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    fe = FrontEnd(synth, mel)
    segs = make_segments(16)
    random.seed(7); ref = [so.render(n, dict(SETTING_1), nested) for n in segs]
    random.seed(7); wav, feat = fe(segs)
    torch.cuda.synchronize()
    wav = wav.cpu().numpy(); feat = feat.cpu().numpy()
    refm = so.collate(ref)
    print("wav shape", wav.shape, refm.shape, "max abs err", float(np.nanmax(np.abs(wav - refm))))
    om = mo.logmel_direct(refm, 24000, 2048, 0.01, 128, np.float64)
    err = np.abs(feat - om); rel = err / np.maximum(np.abs(om), 1e-12)
    print("mel shape", feat.shape, "max abs", float(err.max()), "max rel (|ref|>1e-3)", float(rel[np.abs(om) > 1e-3].max()))
    # mel alone on the oracle wav
    f2 = mel(torch.from_numpy(refm).cuda()).cpu().numpy()
    print("mel-only max abs", float(np.abs(f2 - om).max()))
    # timing
    big = make_bank(2600, max_len=48000)
    synth2 = SynthDrum(cfg, bank=big); fe2 = FrontEnd(synth2, mel)
    segs = make_segments(1024, seed=3)
    random.seed(1); t = time.time(); plan = synth2.plan(segs); print("plan s", time.time() - t, plan.n_events)
    buf = synth2.buffers()
    w, f = fe2.run_plan(plan); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for name, fn in (("fused", lambda: fe2.run_plan(plan, wav=w_full, feat=f, upload=False)),
                     ("render", lambda: synth2.render_plan(plan, out=w_full)),
                     ("logmel", lambda: mel(w_full))):
        w_full = torch.empty((plan.n_seg, plan.ld_wav), device="cuda")
        fn(); torch.cuda.synchronize()
        ev[0].record()
        for _ in range(5): fn()
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print(f"{name}: {ms:.3f} ms / 1024 segments -> {1024*2.56/ms*1e3:.3e} audio-s/s")
    print("bytes_alg per seg", plan.bytes_alg(big, 246, 128) / 1024)

if __name__ == "__main__":
    main()
