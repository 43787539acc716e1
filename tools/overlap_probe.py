#!/usr/bin/env python3
"""Can the render kernels of one plan run under the log-mel kernel of another?  Times render(X) alone, logmel(Y) alone
and both at once on two streams (log-mel launched first)."""
import ctypes as C, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum, _lib
from adt_str_b200.config import setting_1
from adt_str_b200.synthetic import make_bank, make_segments
from adt_str_b200.synthetiser import PlanBuffers

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
bank = make_bank(10000, 24000, seed=0)
synth = SynthDrum(setting_1(), bank=bank, device=dev); mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
fe = FrontEnd(synth, mel)
lib, bh, mh = _lib.load(), synth.device_bank().handle, mel._handle(dev).handle
sets = []
for k in range(2):
    segs = make_segments(nb * 64, seed=10 + k)
    plan = fe.plan_batches([segs[i * 64:(i + 1) * 64] for i in range(nb)], random.Random(k), 4)
    buf = PlanBuffers(dev); buf._dplan = buf.upload(buf.pack(plan)); buf._resident = plan
    wav, feat = fe._outputs(plan, 0)
    fe.run_plan(plan, buffers=buf, wav=wav, feat=feat, upload=False)
    sets.append((plan, buf, wav, feat))
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def render(k, st):
    plan, buf, wav, feat = sets[k]
    _lib.check(lib.adtfe_render(bh, C.byref(buf._dplan), wav.data_ptr(), buf.workspace.data_ptr(), buf.workspace.numel(), st.cuda_stream))

def logmel(k, st):
    plan, buf, wav, feat = sets[k]
    _lib.check(lib.adtfe_logmel_rows(mh, wav.data_ptr(), plan.n_seg, plan.ld_wav, buf._dplan.mel_rows_dev, buf._dplan.mel_max_count, feat.data_ptr(), st.cuda_stream))

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

r = timed(lambda: render(0, sa)); l = timed(lambda: logmel(1, sb))
def both():
    logmel(1, sb); render(0, sa)
def both2():
    render(0, sa); logmel(1, sb)
print(f"batches {nb}: render alone {r:.2f} ms, logmel alone {l:.2f} ms, sum {r + l:.2f}; together (logmel first) {timed(both):.2f} ms, (render first) {timed(both2):.2f} ms")
