"""Audio front of eval / inference (resample, normalize, channel mean, chunks) on the CPU: the oracle against the
fixtures written from the running reference (tests/golden/audio_front.npz), the product's host-side pieces
(filter bank, chunk_audio), and - in the build container - both against the live reference functions."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import audio_oracle, ref_harness

AUDIO_CASES = ["stereo_44k1", "mono_48k", "mono_22k05_16k", "stereo_24k"]


@pytest.fixture(scope="module")
def front():
    return np.load(os.path.join(GOLDEN_DIR, "audio_front.npz"))


@pytest.mark.parametrize("name", AUDIO_CASES)
def test_oracle_matches_reference_fixture(front, name):
    x = front[f"{name}/x"]
    sr, target = (int(v) for v in front[f"{name}/sr"])
    if sr != target:
        want = front[f"{name}/resampled"]
        lib = audio_oracle.resample_torchaudio(x, sr, target)            # the reference's own library call
        assert lib.shape == want.shape and np.abs(lib - want).max() <= 1e-6
        for dt, tol in ((np.float64, 2e-6), (np.float32, 3e-6)):         # explicit polyphase sum
            d = audio_oracle.resample_direct(x, sr, target, dt)
            assert d.shape == want.shape and np.abs(d - want).max() <= tol
    chunks = audio_oracle.long_form_chunks(x, sr, target, 2.56)
    want = front[f"{name}/chunks"]
    assert chunks.shape == want.shape and np.abs(chunks - want).max() <= 1e-6
    chunk = int(round(2.56 * target))
    assert front[f"{name}/chunk_starts"].tolist() == [i * chunk for i in range(len(want))]
    mono = audio_oracle.downmix(x)
    y = audio_oracle.resample_torchaudio(mono, sr, target) if sr != target else mono
    got = audio_oracle.normalize(y)
    want = front[f"{name}/mono_first_normalized"]
    assert got.shape == want.shape and np.abs(got - want).max() <= 2e-6
    assert np.abs(got).max() == 1.0


def test_filter_bank_is_torchaudios_bit_for_bit():
    import torchaudio.transforms as T
    from adt_str_b200.audio_utils import Resample, sinc_resample_kernel
    for o, n in [(44100, 24000), (48000, 24000), (44100, 16000), (22050, 24000), (16000, 24000), (8000, 16000),
                 (48000, 16000), (32000, 24000), (96000, 16000)]:
        ref = T.Resample(o, n)
        k, w = sinc_resample_kernel(o, n)
        assert w == ref.width and torch.equal(k, ref.kernel)
        ours = Resample(o, n)
        assert list(ours.state_dict().keys()) == list(ref.state_dict().keys()) == ["kernel"]
        assert ours.output_length(12345) == int(np.ceil((n // ref.gcd) * 12345 / (o // ref.gcd)))
        k64, w64 = audio_oracle.sinc_kernel(o, n)                        # the oracle's NumPy float64 restatement
        assert w64 == w and np.array_equal(k64.astype(np.float32), k.numpy().reshape(k64.shape))
    k, w = sinc_resample_kernel(44100, 24000, 8, 0.9, "sinc_interp_kaiser", 12.0)
    ref = T.Resample(44100, 24000, "sinc_interp_kaiser", 8, 0.9, 12.0)
    assert w == ref.width and torch.equal(k, ref.kernel)
    assert Resample(24000, 24000)(torch.ones(3)) .tolist() == [1.0, 1.0, 1.0]   # same rate: the input itself
    with pytest.raises(ValueError):
        sinc_resample_kernel(44100, 24000, resampling_method="linear")


def _ref_chunks(wav, chunk):
    """inference.py:35-48 semantics, from the oracle: (start, chunk) pairs of a (channels, samples) signal."""
    out = []
    for c in range(wav.shape[0]):
        out.append(audio_oracle.chunk_audio(wav[c].numpy(), chunk))
    n = out[0].shape[0] if out else 0
    return [(i * chunk, np.stack([o[i] for o in out])) for i in range(n)]


@pytest.mark.parametrize("n,chunk,ch", [(1000, 300, 1), (900, 300, 2), (1, 300, 1), (299, 300, 3), (0, 300, 1)])
def test_chunk_audio_mirrors_the_reference(n, chunk, ch):
    from adt_str_b200.inference_front import chunk_audio
    g = torch.Generator().manual_seed(n + ch)
    wav = torch.randn(ch, n, generator=g)
    got = chunk_audio(wav, chunk)
    want = _ref_chunks(wav, chunk)
    assert [s for s, _ in got] == [s for s, _ in want]
    for (_, a), (_, b) in zip(got, want):
        assert tuple(a.shape) == (ch, chunk) and np.array_equal(a.numpy(), b)
    if ref_harness.available():     # and the unmodified reference function itself
        _, inf = ref_harness.import_audio_front()
        live = inf._chunk_audio(wav, chunk)
        assert [s for s, _ in live] == [s for s, _ in got]
        for (_, a), (_, b) in zip(got, live):
            assert torch.equal(a, b)
    with pytest.raises(ValueError):
        chunk_audio(wav[0], chunk)
    with pytest.raises(ValueError):
        chunk_audio(wav, 0)


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")
def test_oracle_against_live_reference_functions():
    au, _ = ref_harness.import_audio_front()
    rng = np.random.default_rng(5)
    for o, n, length in [(44100, 24000, 30011), (48000, 16000, 9000), (16000, 24000, 777), (22050, 24000, 4000)]:
        x = rng.standard_normal(length).astype(np.float32)
        ref = au.resample(torch.from_numpy(x), o, n).numpy()
        d = audio_oracle.resample_direct(x, o, n, np.float64)
        assert d.shape == ref.shape and np.abs(d - ref).max() < 3e-6
        assert np.array_equal(audio_oracle.resample_torchaudio(x, o, n), ref)
        assert np.array_equal(audio_oracle.normalize(ref), au.normalize(torch.from_numpy(ref)).numpy())
    st = rng.standard_normal((2, 5000)).astype(np.float32)
    assert np.array_equal(audio_oracle.downmix(st), torch.from_numpy(st).mean(0).numpy())
