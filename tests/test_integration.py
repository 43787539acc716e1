"""The drop-in inside the reference's own model and data pipeline (CPU box: needs /root/reference)."""
import random

import numpy as np
import pytest
import torch

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")


def _small_config(model_mod):
    import importlib
    cfg_mod = importlib.import_module("config")
    return cfg_mod.ADTModelConfig(input_sec=2.56, time_res=0.01, win_length=2048, sample_rate=24000, enc_layers=1,
                                  dec_layers=1, nhead=2, d_query=16, dropout=0.0, tgt_vocab_size=600, plain=True,
                                  n_mels=128)


def test_reference_model_with_swapped_mel_loads_reference_checkpoints_strictly():
    """``ADTModel`` (model.py:194-226) built with ``adt_str_b200.ComputeMelSpectrogram`` in place of its own: the
    state dict has the same keys, shapes and buffer values, a checkpoint of the unmodified model loads with
    ``strict=True`` (build_model.py:66) and a checkpoint of the swapped model loads into the unmodified one."""
    from adt_str_b200 import integration
    _, model_mod = ref_harness._import_reference()
    original_cls = model_mod.ComputeMelSpectrogram
    torch.manual_seed(0)
    ref_model = model_mod.ADTModel(_small_config(model_mod))
    sd = ref_model.state_dict()
    try:
        integration.install_mel(model_mod)
        swapped = model_mod.ADTModel(_small_config(model_mod))
    finally:
        model_mod.ComputeMelSpectrogram = original_cls
    from adt_str_b200 import ComputeMelSpectrogram
    assert isinstance(swapped.compute_spectrogram, ComputeMelSpectrogram)
    assert swapped.compute_spectrogram.window_pad_idxs == ref_model.compute_spectrogram.window_pad_idxs == 5
    ssd = swapped.state_dict()
    assert list(ssd.keys()) == list(sd.keys())
    for k in sd:
        assert ssd[k].shape == sd[k].shape and ssd[k].dtype == sd[k].dtype, k
    for k in ("compute_spectrogram.compute_spec.spectrogram.window", "compute_spectrogram.compute_spec.mel_scale.fb"):
        assert torch.equal(ssd[k], sd[k]), k                      # the buffers themselves are bit-identical
    missing = swapped.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k in sd:
        assert torch.equal(swapped.state_dict()[k], sd[k])
    back = ref_model.load_state_dict(swapped.state_dict(), strict=True)
    assert not back.missing_keys and not back.unexpected_keys
    # the rest of the model is untouched: the projection the log-mel feeds is the reference's own Linear
    assert swapped.project_to_mel.weight.shape == (32, 128)


def test_collate_notes_is_the_reference_collate_for_the_token_half():
    from adt_str_b200 import integration
    tok_mod, ref_collate = ref_harness.import_tokenizer()
    if ref_collate is None:
        pytest.skip("the reference's dataset module cannot be imported here")
    rng = np.random.default_rng(0)
    synth = integration.DeferredSynth()
    items_ref, items_new = [], []
    for n_tok in (5, 9, 9, 3, 1):
        tokens = rng.integers(0, 500, n_tok).tolist()
        notes = rng.random((n_tok, 4)).astype(np.float32)
        items_ref.append((torch.zeros(100 + n_tok), tokens))
        items_new.append((synth(torch.from_numpy(notes)), tokens))
    items_ref.append((torch.zeros(61440), [2, 0, 3]))             # the dataset's "empty" item: zeros + [BOS, 0, EOS]
    items_new.append((torch.zeros(61440), [2, 0, 3]))
    want, got = ref_collate(items_ref), integration.collate_notes(items_new)
    assert torch.equal(got["tokens"], want["tokens"]) and torch.equal(got["token_lengths"], want["token_lengths"])
    assert got["tokens"].dtype == want["tokens"].dtype and set(got) == {"notes", "tokens", "token_lengths"}
    assert [len(n) for n in got["notes"]] == [5, 9, 9, 3, 1, 0] and all(n.dtype == np.float32 for n in got["notes"])
    with pytest.raises(TypeError):
        integration.collate_notes([(torch.ones(10), [1, 2])])
    with pytest.raises(NotImplementedError):
        synth([[0.1, 0.2, 36, 100]], eval_rendering=True)
