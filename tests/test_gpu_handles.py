"""Several handles of one kind alive in one process, used alternately and on concurrent streams.

Kernel attributes (the dynamic shared-memory limit) belong to the *function*, internal streams and events to the
*handle*: a second handle must change nothing for the first.  Every case computes a result with one handle alone,
creates a second handle with a different geometry, and then expects the first result again, bit for bit - also while
the second handle's kernels run on another stream.  (The projection's own case, with a library GEMM beside it, is in
test_gpu_projection.py.)"""
import random

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _both(fn_a, fn_b, want_a, want_b, reps=4):
    dev = torch.device("cuda", 0)
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    for _ in range(reps):
        with torch.cuda.stream(sa):
            got_a = fn_a()
        with torch.cuda.stream(sb):
            got_b = fn_b()
    torch.cuda.synchronize()
    assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b)
    for _ in range(reps):                      # and strictly alternating on one stream
        got_a, got_b = fn_a(), fn_b()
    torch.cuda.synchronize()
    assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b)


def test_two_mel_front_ends_of_different_geometry():
    from adt_str_b200 import ComputeMelSpectrogram
    g = torch.Generator().manual_seed(3)
    xa, xb = torch.randn(48, 61440, generator=g).cuda(), torch.randn(40, 40960, generator=g).cuda()
    mel_a = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    want_a = mel_a(xa).clone()
    mel_b = ComputeMelSpectrogram(16000, 2048, 0.01, 64)         # another hop, another filterbank
    want_b = mel_b(xb).clone()
    mel_c = ComputeMelSpectrogram(22050, 2048, 0.0107, 80)       # an odd hop: the generic kernel
    want_c = mel_c(xb).clone()
    assert torch.equal(mel_a(xa), want_a)
    _both(lambda: mel_a(xa), lambda: mel_b(xb), want_a, want_b)
    _both(lambda: mel_c(xb), lambda: mel_a(xa), want_c, want_a)


def test_two_resamplers_of_different_geometry():
    from adt_str_b200 import Resample
    g = torch.Generator().manual_seed(4)
    xa, xb = torch.randn(2, 300000, generator=g).cuda(), torch.randn(3, 200000, generator=g).cuda()
    rs_a = Resample(44100, 24000)      # 147 -> 80: long polyphase table
    want_a = rs_a(xa).clone()
    rs_b = Resample(48000, 16000)      # 3 -> 1
    want_b = rs_b(xb).clone()
    rs_c = Resample(44100, 48000)
    want_c = rs_c(xb).clone()
    assert torch.equal(rs_a(xa), want_a)
    _both(lambda: rs_a(xa), lambda: rs_b(xb), want_a, want_b)
    _both(lambda: rs_c(xb), lambda: rs_a(xa), want_c, want_a)


def test_two_banks_and_synthesisers_render_side_by_side():
    from adt_str_b200 import ComputeMelSpectrogram, FrontEnd, SynthDrum
    from adt_str_b200.config import setting_1
    from adt_str_b200.synthetic import make_bank, make_segments
    dev = torch.device("cuda", 0)
    segs_a, segs_b = make_segments(96, seed=5), make_segments(64, seed=6)
    mel = ComputeMelSpectrogram(24000, 2048, 0.01, 128)
    fe_a = FrontEnd(SynthDrum(setting_1(), bank=make_bank(400, 24000, seed=1), device=dev), mel)
    plan_a = fe_a.plan_batches([segs_a[:64], segs_a[64:]], random.Random(1), 1)

    def run_a():
        return fe_a.run_plan(plan_a)[1]

    want_a = run_a().clone()
    fe_b = FrontEnd(SynthDrum(setting_1(use_fx_prob=1.0), bank=make_bank(156, 24000, max_len=9000, seed=2), device=dev), mel)
    torch.manual_seed(9)
    plan_b = fe_b.plan_batches([segs_b], random.Random(2), 1)

    def run_b():
        return fe_b.run_plan(plan_b)[1]

    want_b = run_b().clone()
    assert torch.equal(run_a(), want_a)
    _both(run_a, run_b, want_a, want_b, reps=3)
