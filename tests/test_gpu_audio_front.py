"""GPU parity of the eval / inference audio front (adtfe_resample, adtfe_downmix, adtfe_peak_normalise, chunks ->
log-mel) against the reference fixtures (tests/golden/audio_front.npz) and the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, assert_logmel_close
from oracle import audio_oracle, mel_oracle

pytestmark = pytest.mark.gpu

RESAMPLE_TOL = 1e-5   # same bar as the rendered waveform: max abs error against the reference CPU path
AUDIO_CASES = ["stereo_44k1", "mono_48k", "mono_22k05_16k", "stereo_24k"]


@pytest.fixture(scope="module", autouse=True)
def _needs_b200():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from adt_str_b200 import _lib
    _lib.check(_lib.load().adtfe_device_ok(0), "adtfe_device_ok")


@pytest.fixture(scope="module")
def front():
    return np.load(os.path.join(GOLDEN_DIR, "audio_front.npz"))


@pytest.mark.parametrize("name", AUDIO_CASES)
def test_long_form_front_matches_reference_fixture(front, name):
    from adt_str_b200 import ComputeMelSpectrogram, LongFormFrontEnd
    from adt_str_b200.audio_utils import normalize, resample
    x = torch.from_numpy(front[f"{name}/x"])
    sr, target = (int(v) for v in front[f"{name}/sr"])
    if sr != target:
        want = front[f"{name}/resampled"]
        got = resample(x, sr, target)                      # CPU in -> CPU out, computed on the GPU
        assert not got.is_cuda and got.shape == want.shape
        assert float(np.abs(got.numpy() - want).max()) <= RESAMPLE_TOL
        truth = audio_oracle.resample_direct(x.numpy(), sr, target, np.float64)
        assert float(np.abs(got.numpy() - truth).max()) <= 2e-6
    fe = LongFormFrontEnd(target, 2.56, ComputeMelSpectrogram(target, 2048, 0.01, 128))
    chunks, mel = fe(x.cuda(), sr)
    want = front[f"{name}/chunks"]
    assert chunks.is_cuda and tuple(chunks.shape) == want.shape
    assert float(np.abs(chunks.cpu().numpy() - want).max()) <= RESAMPLE_TOL
    tail = int(x.shape[1] * target // sr) + 2 - (want.shape[0] - 1) * want.shape[1]
    assert not chunks[-1, tail:].any()                     # the last chunk is zero-padded
    truth = mel_oracle.logmel_direct(want, target, 2048, 0.01, 128, np.float64)
    assert_logmel_close(mel.cpu().numpy(), front[f"{name}/mel"], truth=truth, atol=2e-6)
    # eval_dataset.py:69-70 order: mean first, then resample, then normalize
    from adt_str_b200.audio_utils import downmix
    mono = downmix(x.cuda())
    y = resample(mono, sr, target) if sr != target else mono
    got = normalize(y).cpu().numpy()
    want = front[f"{name}/mono_first_normalized"]
    assert got.shape == want.shape and float(np.abs(got - want).max()) <= RESAMPLE_TOL
    assert float(np.abs(got).max()) == 1.0


@pytest.mark.parametrize("orig,new", [(44100, 24000), (48000, 24000), (44100, 16000), (22050, 24000), (16000, 24000),
                                      (8000, 16000), (48000, 16000), (96000, 16000), (44100, 48000)])
def test_resample_rate_pairs_and_ragged_lengths(orig, new):
    """Every length class: shorter than the filter, one sample, exact multiples of the polyphase period, tile
    boundaries; several rows with a row pitch; against torchaudio on CPU (the reference's call) and float64."""
    from adt_str_b200 import Resample
    rs = Resample(orig, new)
    o = orig // rs.gcd
    g = torch.Generator().manual_seed(orig + new)
    for n in sorted({1, 2, max(1, o - 1), o, o + 1, 7 * o, 4096, 10007, 3 * 4096 * o // (new // rs.gcd) + 5}):
        x = torch.randn(3, n, generator=g)
        want = audio_oracle.resample_torchaudio(x.numpy(), orig, new)
        got = rs(x.cuda())
        assert got.is_cuda and tuple(got.shape) == want.shape, (n, got.shape, want.shape)
        assert float(np.abs(got.cpu().numpy() - want).max()) <= RESAMPLE_TOL, n
        truth = audio_oracle.resample_direct(x.numpy(), orig, new, np.float64)
        assert float(np.abs(got.cpu().numpy() - truth).max()) <= 4e-6, n
    wide = torch.randn(4, 5000, generator=g).cuda()
    view = wide[:, 100:3100]                                # rows with a pitch, unaligned start
    assert torch.equal(rs(view), rs(view.contiguous()))
    lead = torch.randn(2, 3, 1000, generator=g)             # leading dimensions are kept (..., time)
    assert tuple(rs(lead).shape) == (2, 3, rs.output_length(1000))
    assert tuple(rs(torch.zeros(2, 0)).shape) == (2, 0)


def test_resample_properties_at_full_length():
    """BASELINE config 5 size: a 10-minute stereo signal at 44.1 kHz.  Size-independent properties: linearity,
    shift equivariance by one polyphase period (bit for bit away from the edges), determinism, chunk count."""
    from adt_str_b200 import ComputeMelSpectrogram, LongFormFrontEnd, Resample
    rs = Resample(44100, 24000)
    n = 600 * 44100
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(2, n, generator=g, device="cuda") * 0.3
    b = torch.randn(2, n, generator=g, device="cuda") * 0.3
    ya, yb, yab = rs(a), rs(b), rs(a + b)
    assert ya.shape == (2, 600 * 24000)
    assert float((yab - (ya + yb)).abs().max()) <= 2e-6
    assert torch.equal(rs(a), ya)
    shifted = rs(torch.cat([a[:, 147 * 5:], torch.zeros(2, 147 * 5, device="cuda")], 1))   # 147 in -> 80 out
    assert torch.equal(shifted[:, 200:-2000], ya[:, 80 * 5 + 200:80 * 5 + shifted.shape[1] - 2000])
    m = 12_500                                              # a window checked against float64: a piece that starts on
    lo = 147 * m                                            # a polyphase period resamples to the same samples
    ref = audio_oracle.resample_direct(a[0, lo:lo + 100_000].cpu().numpy(), 44100, 24000, np.float64)
    assert float(np.abs(ya[0, 80 * m + 200:80 * m + 50_000].cpu().numpy() - ref[200:50_000]).max()) <= 2e-6
    fe = LongFormFrontEnd(24000, 2.56, ComputeMelSpectrogram(24000, 2048, 0.01, 128))
    chunks, mel = fe(a, 44100)
    assert tuple(chunks.shape) == (235, 61440) and tuple(mel.shape) == (235, 246, 128)
    mono = (ya[0] + ya[1]) / 2
    assert torch.equal(chunks.reshape(-1)[:mono.numel()], mono) and not chunks.reshape(-1)[mono.numel():].any()
    assert torch.isfinite(mel).all() and float(mel.min()) >= 0.0 and float(mel.max()) <= 1.0
    assert torch.equal(mel[:3], fe.mel(chunks[:3]))


def test_normalize_and_downmix_bit_exact_and_special_values():
    from adt_str_b200.audio_utils import downmix, normalize
    g = torch.Generator().manual_seed(8)
    x = torch.randn(100_003, generator=g) * 0.37
    got = normalize(x.cuda())
    assert torch.equal(got.cpu(), x / x.abs().max())        # IEEE division, same maximum: bit for bit
    assert not normalize(x).is_cuda and torch.equal(normalize(x), got.cpu())
    keep = x.cuda()
    normalize(keep)
    assert torch.equal(keep.cpu(), x)                       # a new tensor, the input is untouched
    assert torch.isnan(normalize(torch.zeros(10).cuda())).all()          # 0/0, like the reference
    y = x.clone(); y[17] = float("nan")
    assert torch.isnan(normalize(y.cuda())).all()           # torch.max propagates NaN
    z = x.clone(); z[5] = float("inf")
    nz = normalize(z.cuda()).cpu()
    assert torch.isnan(nz[5]) and float(nz[6]) == 0.0
    for ch in (1, 2, 3, 6):
        s = torch.randn(ch, 40_001, generator=g)
        want = audio_oracle.downmix(s.numpy())
        assert np.array_equal(downmix(s.cuda()).cpu().numpy(), want)
        if ch <= 2:
            assert torch.equal(downmix(s.cuda(), keepdim=True).cpu(), s.mean(0, keepdim=True))
    with pytest.raises(ValueError):
        downmix(torch.zeros(5).cuda())


def test_resampler_abi_status_codes():
    from adt_str_b200 import Resample, _lib
    lib = _lib.load()
    with pytest.raises(_lib.AdtfeError):
        Resample(44100, 22051)(torch.zeros(1, 100).cuda())  # gcd 1: a 44100-phase filter bank, unsupported
    with pytest.raises(TypeError):
        Resample(44100, 24000)(torch.zeros(1, 100, dtype=torch.int32))
    rs = Resample(48000, 24000)
    h = rs._handle(torch.device("cuda", 0)).handle
    assert lib.adtfe_resample_length(h, 1001) == 501 and lib.adtfe_resample_length(h, -1) == -1
    x = torch.zeros(1, 1000, device="cuda")
    y = torch.zeros(1, 400, device="cuda")
    assert lib.adtfe_resample(h, x.data_ptr(), 1, 1000, 1000, y.data_ptr(), 400, None, None) == -1   # ld_out too small
    assert b"ld_out" in lib.adtfe_last_error()
    assert lib.adtfe_resample(h, x.data_ptr(), 1, 10, 1000, y.data_ptr(), 500, None, None) == -1     # ld_in < n_in
    assert lib.adtfe_resample(None, x.data_ptr(), 1, 1000, 1000, y.data_ptr(), 500, None, None) == -1
    out = C.c_void_p()
    k = torch.zeros(4, 100)
    assert lib.adtfe_resampler_create(0, 24000, 6, k.data_ptr(), 0, C.byref(out)) == -1
    assert lib.adtfe_resampler_create(48000, 24000, 13, None, 0, C.byref(out)) == -1
    bits = torch.zeros(1, dtype=torch.int32, device="cuda")
    big = torch.randn(1, 30_000, device="cuda")
    out_t = torch.empty(1, rs.output_length(30_000), device="cuda")
    rs.resample_into(big, out_t, bits)                      # the fused |max| equals the tensor's
    assert float(bits.view(torch.float32)) == float(out_t.abs().max())
    assert lib.adtfe_peak_normalise(out_t.data_ptr(), out_t.numel(), bits.data_ptr(), 1, None) == 0
    torch.cuda.synchronize()
    assert float(out_t.abs().max()) == 1.0
