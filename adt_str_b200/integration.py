"""Adapters that put the GPU front end behind the reference's training loop without editing its files.

The reference renders inside forked DataLoader workers (``LakhDataset.__getitem__`` -> ``self.synthetiser(notes)``,
``data_modules/train_dataset.py:213-229``), pads the waveforms in ``collate_fn`` (``:41-56``), copies them to the
device in ``ADTTrainer.compute_loss`` (``train.py:40-70``) and featurises them in ``ADTModel.forward``
(``model.py:248``).  CUDA does not survive a fork, so the drop-in moves the *rendering* to the main process and
keeps everything else where it is:

    import model as ref_model, train as ref_train                      # the reference's own modules
    from adt_str_b200 import SynthDrum, integration

    integration.install_mel(ref_model)                                 # ADTModel builds the CUDA ComputeMelSpectrogram
    dataset = LakhDataset(cfg, tokenizer, integration.DeferredSynth()) # workers hand the NOTES through
    loader  = DataLoader(dataset, collate_fn=integration.collate_notes, ...)
    renderer = integration.BatchRenderer(SynthDrum(synth_cfg))         # main process, owns the bank on the GPU
    for batch in loader:
        batch = renderer(batch)                                        # adds "wavs" (B, Lmax) on the device
        loss = trainer.compute_loss(model, batch)                      # unchanged: wavs.to(device) is a no-op

Token handling is the reference's (``collate_fn`` decrements the lengths that equal the maximum, ``:47-51``).
"""
from __future__ import annotations

import random as _random
from typing import Any, Dict, List, Sequence

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence


class PendingNotes:
    """What ``DeferredSynth`` returns instead of audio: the notes of one item, to be rendered with its batch."""
    __slots__ = ("notes",)

    def __init__(self, notes):
        if isinstance(notes, torch.Tensor):
            notes = notes.detach().cpu().numpy()
        self.notes = np.asarray(notes, np.float32).reshape(-1, 4) if len(notes) else np.zeros((0, 4), np.float32)


class DeferredSynth:
    """Takes ``SynthDrum``'s place inside the dataset (``LakhDataset(config, tokenizer, synthetiser)``,
    ``train_dataset.py:179,228``): same call signature, but the notes travel on to the collate function."""

    def __call__(self, notes, eval_rendering: bool = False) -> PendingNotes:
        if eval_rendering:
            raise NotImplementedError("eval_rendering is not implemented (undefined in the reference as well)")
        return PendingNotes(notes)


def collate_notes(batch: Sequence) -> Dict[str, Any]:
    """``collate_fn`` of the reference (``train_dataset.py:41-56``) for items ``(PendingNotes | waveform, tokens)``:
    the token half is the reference's, key for key; ``"wavs"`` is replaced by ``"notes"`` (one float32 (N, 4) array
    per item).  The all-zero waveform the dataset returns for its "empty" items (``_empty_wav``, ``:214-215``) comes
    back as an empty note list - which renders to the same zeros."""
    pad_token = 1
    notes: List[np.ndarray] = []
    for item in batch:
        first = item[0]
        if isinstance(first, PendingNotes):
            notes.append(first.notes)
        elif isinstance(first, torch.Tensor) and first.ndim == 1 and not bool(first.any()):
            notes.append(np.zeros((0, 4), np.float32))
        else:
            raise TypeError("collate_notes expects items from a dataset built with DeferredSynth()")
    token_lengths = [len(item[1]) for item in batch]
    tokens = [torch.tensor(item[1]) for item in batch]
    max_value = max(token_lengths) if len(token_lengths) > 0 else 0
    if max_value > 0:   # lengths equal to the maximum are decreased by one, like the reference
        token_lengths = [n - 1 if n == max_value else n for n in token_lengths]
    return {
        "notes": notes,
        "tokens": pad_sequence(tokens, batch_first=True, padding_value=pad_token).long(),
        "token_lengths": torch.tensor(token_lengths).long(),
    }


class BatchRenderer:
    """Main-process half of the drop-in: turns ``batch["notes"]`` into ``batch["wavs"]`` - the collated
    ``(B, Lmax)`` float32 matrix ``pad_sequence`` would have built, already on the GPU - with one render per batch.
    With ``mel`` (an ``adt_str_b200.ComputeMelSpectrogram``) the log-mel comes out of the same fused call as
    ``batch["logmel"]``."""

    def __init__(self, synth, mel=None, rng=_random):
        self.synth, self.mel, self.rng = synth, mel, rng
        self._front = None
        if mel is not None:
            from .frontend import FrontEnd
            self._front = FrontEnd(synth, mel)

    def __call__(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        out = dict(batch)
        notes = out.pop("notes")
        if self._front is not None:
            out["wavs"], out["logmel"] = self._front(notes, self.rng)
        else:
            out["wavs"], _ = self.synth.render_batch(notes, self.rng)
        return out


def install_mel(reference_model_module) -> None:
    """Point the reference's ``model`` module at the CUDA ``ComputeMelSpectrogram``: ``ADTModel.__init__``
    (``model.py:218-223``) then builds it, with the same constructor arguments and the same two state-dict buffers, so
    existing checkpoints load strictly (``build_model.py:66``)."""
    from .mel import ComputeMelSpectrogram
    reference_model_module.ComputeMelSpectrogram = ComputeMelSpectrogram


def install_projection(model) -> None:
    """Swap ``model.project_to_mel`` (``nn.Linear(n_mels, d_query * nhead)``, ``model.py:224-226``) of a built
    ``ADTModel`` for the tcgen05 ``ProjectToMel`` sharing the same parameters: state-dict keys, values and the
    optimiser's parameter objects stay what they were."""
    from .projection import ProjectToMel
    model.project_to_mel = ProjectToMel.from_linear(model.project_to_mel)
