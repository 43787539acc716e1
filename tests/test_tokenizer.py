"""MidiTokenizer mirror (adt_str_b200/midi_tokenizer.py) against fixtures written from the running reference
(tests/golden/tokens.npz) and - in the build container - against the live reference classes: values AND dtypes."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import ref_harness
from adt_str_b200.midi_tokenizer import MidiTokenizer, MidiTokenizerConfig


@pytest.fixture(scope="module")
def fixture():
    z = np.load(os.path.join(GOLDEN_DIR, "tokens.npz"))
    cuts = np.concatenate([[0], np.cumsum(z["notes_count"])])
    segs = [z["notes"][cuts[i]: cuts[i + 1]].copy() for i in range(len(z["notes_count"]))]
    return z, segs


def _split(flat, counts):
    cuts = np.concatenate([[0], np.cumsum(counts)])
    return [flat[cuts[i]: cuts[i + 1]] for i in range(len(counts))]


@pytest.mark.parametrize("adtof", [False, True])
@pytest.mark.parametrize("vel", [False, True])
def test_tokens_decode_and_collate_match_the_reference_fixture(fixture, adtof, vel):
    z, segs = fixture
    key = f"adtof{int(adtof)}_vel{int(vel)}"
    t = MidiTokenizer(MidiTokenizerConfig(adtof, 1, 2, 0, 3, vel))
    mapped = [t.map_notes_to_Gm_custom(torch.from_numpy(s.copy())) if len(s) else s for s in segs]
    toks = t.encode_batch(mapped)
    want = _split(z[f"{key}/tokens"], z[f"{key}/count"])
    assert [len(x) for x in toks] == z[f"{key}/count"].tolist()
    assert [x.is_floating_point() for x in toks] == z[f"{key}/float"].tolist()
    for a, b in zip(toks, want):
        assert np.array_equal(a.double().numpy(), b)
    dec = t.batch_decode([x.numpy() for x in toks])
    want_dec = _split(z[f"{key}/decoded"], z[f"{key}/decoded_count"])
    for a, b in zip(dec, want_dec):
        assert np.array_equal(a.double().numpy().reshape(-1, 4), b)
    c = t.collate_tokens(toks)
    assert c["tokens"].dtype == torch.int64 and c["token_lengths"].dtype == torch.int64
    assert np.array_equal(c["tokens"].numpy(), z[f"{key}/collated"])
    assert np.array_equal(c["token_lengths"].numpy(), z[f"{key}/collated_lengths"])


def test_input_kinds_dtypes_and_errors():
    t = MidiTokenizer(MidiTokenizerConfig(False, 1, 2, 0, 3, True))
    assert t.notes_to_adt_tokens([[0.07, 0.17, 36, 100]]).tolist() == [1, 11, 336, 500, 2]          # ints stay int64
    assert t.notes_to_adt_tokens([[0.07, 0.17, 36, 100]]).dtype == torch.int64
    f32 = torch.tensor([[1.289999, 1.39, 42.0, 64.0]])
    assert t.notes_to_adt_tokens(f32).tolist() == [1.0, 132.0, 342.0, 464.0, 2.0]                   # int(f32 * 100) = 128
    assert t.notes_to_adt_tokens(f32).dtype == torch.float32
    assert t.notes_to_adt_tokens(f32.double()).tolist() == [1.0, 132.0, 342.0, 464.0, 2.0]
    assert t.notes_to_adt_tokens(f32.double().numpy()).dtype == torch.float64
    assert t.empty_adt_tokens().tolist() == [1, 3, 2]
    with pytest.raises(AssertionError):
        t.notes_to_adt_tokens(torch.tensor([[2.97, 3.0, 36.0, 100.0]]))                             # time token 301
    with pytest.raises(KeyError):
        t.map_notes_to_Gm_custom(torch.tensor([[0.0, 0.1, 20.0, 100.0]]))
    assert t.collate_tokens([])["tokens"].shape == (0, 0)
    # a pitch without its time token is dropped, the next pair decodes (velocity defaults to 100)
    assert np.allclose(t.decode([1, 336, 11, 342, 2]).numpy(), [[0.07, 0.17, 42.0, 100.0]])


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")
def test_against_live_reference_classes():
    from adt_str_b200.synthetic import make_segments
    tok, collate = ref_harness.import_tokenizer()
    segs = make_segments(40, seed=7, empty_fraction=0.1)
    for adtof in (False, True):
        for vel in (False, True):
            cfg = (adtof, 1, 2, 0, 3, vel)
            ref, ours = tok.MidiTokenizer(tok.MidiTokenizerConfig(*cfg)), MidiTokenizer(MidiTokenizerConfig(*cfg))
            assert ours.adt_tokens_offset_dict == ref.adt_tokens_offset_dict
            assert ours.GM_standard_midi_to_Gm_custom_map == ref.GM_standard_midi_to_Gm_custom_map
            r_toks, o_toks = [], []
            for s in segs:
                if len(s) == 0:
                    r_toks.append(ref.empty_adt_tokens()); o_toks.append(ours.empty_adt_tokens())
                    continue
                torch.manual_seed(5)
                rn = ref.map_notes_to_Gm_custom(torch.from_numpy(s.copy()), random_velocity=True)
                torch.manual_seed(5)
                on = ours.map_notes_to_Gm_custom(torch.from_numpy(s.copy()), random_velocity=True)
                assert torch.equal(rn, on)
                r_toks.append(ref.notes_to_adt_tokens(rn)); o_toks.append(ours.notes_to_adt_tokens(on))
                assert torch.equal(ref.notes_to_adt_tokens(rn.tolist()), ours.notes_to_adt_tokens(on.tolist()))
            for a, b in zip(r_toks, o_toks):
                assert a.dtype == b.dtype and torch.equal(a, b)
                assert torch.equal(ref.decode(a.numpy()), ours.decode(b.numpy()))
                if not adtof:   # with ADTOF_mapping the reference's own decode raises KeyError on tensor tokens
                    assert torch.equal(ref.decode(a), ours.decode(b))
            if collate is not None:
                want = collate([(torch.zeros(3), x.tolist()) for x in r_toks])
                got = ours.collate_tokens(o_toks)
                assert torch.equal(want["tokens"], got["tokens"])
                assert torch.equal(want["token_lengths"], got["token_lengths"])
