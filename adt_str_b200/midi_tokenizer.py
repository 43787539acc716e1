"""Token side of a training item, with the reference's call signatures.

Mirrors ``modules/midi_tokenizer.py`` of pier-maker92/ADT_STR (``MidiTokenizerConfig``, ``MidiTokenizer``:
``map_notes_to_Gm_custom`` :36-47, ``notes_to_adt_tokens`` :49-64, ``empty_adt_tokens`` :66-67, ``decode`` :69-100,
``batch_decode`` :102-103) and the token half of ``collate_fn`` (``data_modules/train_dataset.py:41-56``).  It is host
logic (integers): the per-note Python loops of the reference become array operations over a whole segment or a whole
batch, bit-exact in values *and* dtypes - the reference's tokens of a float note tensor are float32
(``torch.tensor`` of a list that holds 0-dim float tensors), of a list of Python ints int64, and ``collate_fn`` pads
with the literal 1 and shortens the longest lengths by one; all of that is kept.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from .mapping import ADTOF_MAP

# GM standard pitch 35..81 -> GM-custom pitch (utils/mapping_utils.py:3-51), as a compact table
_GM_CUSTOM_OF = (35, 36, 37, 38, 39, 40, 41, 42, 41, 43, 41, 44, 45, 45, 46, 47, 48, 49, 48, 50, 51, 52, 46, 53, 48,
                 54, 54, 54, 54, 54, 54, 54, 52, 52, 55, 55, 56, 56, 57, 57, 58, 58, 58, 59, 59, 60, 60)
GM_STANDARD_TO_CUSTOM = {p: c for p, c in zip(range(35, 82), _GM_CUSTOM_OF)}
TIME_OFFSET, PITCH_OFFSET, VELOCITY_OFFSET = 4, 300, 400     # midi_tokenizer.py:25-29
COLLATE_PAD_TOKEN = 1                                         # train_dataset.py:42 (not the tokenizer's pad_token)


@dataclass
class MidiTokenizerConfig:
    ADTOF_mapping: bool
    BOS_token: int
    EOS_token: int
    pad_token: int
    silence_token: int
    add_velocity: bool


class MidiTokenizer:
    def __init__(self, config: MidiTokenizerConfig):
        self.ADTOF_mapping = config.ADTOF_mapping
        self.ADTOF_map = dict(ADTOF_MAP)
        self.GM_standard_midi_to_Gm_custom_map = dict(GM_STANDARD_TO_CUSTOM)
        self.adt_tokens_offset_dict = {"time": TIME_OFFSET, "pitch": PITCH_OFFSET, "velocity": VELOCITY_OFFSET}
        self.BOS_token = config.BOS_token
        self.EOS_token = config.EOS_token
        self.pad_token = config.pad_token
        self.silence_token = config.silence_token
        self.add_velocity = config.add_velocity
        lut = np.full(128, -1, np.int64)
        for p, c in GM_STANDARD_TO_CUSTOM.items():
            lut[p] = self.ADTOF_map[c] if self.ADTOF_mapping else c
        self._pitch_lut = lut

    # ---- midi_tokenizer.py:36-47 (in place, like the reference)
    def map_notes_to_Gm_custom(self, notes, random_velocity=False):
        keys = notes[:, 2].to(torch.int64) if isinstance(notes, torch.Tensor) else np.asarray(notes[:, 2]).astype(np.int64)
        idx = np.asarray(keys)
        bad = (idx < 0) | (idx > 127)
        mapped = self._pitch_lut[np.clip(idx, 0, 127)]
        if bad.any() or (mapped < 0).any():
            first = int(idx[bad | (mapped < 0)][0])
            raise KeyError(first)                       # the dict lookup of the reference
        if isinstance(notes, torch.Tensor):
            notes[:, 2] = torch.from_numpy(mapped)
            if random_velocity:
                notes[:, 3] = torch.randint(10, 127, (notes.shape[0],))
        else:
            notes[:, 2] = mapped
            if random_velocity:
                notes[:, 3] = torch.randint(10, 127, (notes.shape[0],)).numpy()
        return notes

    # ---- midi_tokenizer.py:49-64
    def _token_array(self, notes: np.ndarray) -> np.ndarray:
        """Float note rows (n, 4) -> the token values in the notes' dtype (BOS ... EOS)."""
        dt = notes.dtype.type
        onset = np.trunc(notes[:, 0] * dt(100)).astype(np.int64)        # int(onset * 100), product in the notes' dtype
        time = onset + TIME_OFFSET
        assert (time < PITCH_OFFSET).all(), "Time token is out of range"
        cols = [time.astype(notes.dtype), notes[:, 2] + dt(PITCH_OFFSET)]
        if self.add_velocity:
            cols.append(notes[:, 3] + dt(VELOCITY_OFFSET))
        body = np.stack(cols, axis=1).reshape(-1)
        return np.concatenate([np.array([self.BOS_token], notes.dtype), body, np.array([self.EOS_token], notes.dtype)])

    def notes_to_adt_tokens(self, notes, **kwargs):
        "Notes is intended to be all the notes in one segment"
        if isinstance(notes, torch.Tensor) and notes.is_floating_point() and notes.dim() == 2:
            return torch.from_numpy(self._token_array(notes.detach().cpu().numpy()))
        if isinstance(notes, np.ndarray) and notes.dtype.kind == "f" and notes.ndim == 2:
            # rows of a float ndarray are NumPy scalars: torch.tensor(list) gives their dtype (float32 / float64)
            return torch.from_numpy(self._token_array(notes))
        tokens = [self.BOS_token]                        # anything else: the reference's loop, verbatim semantics
        for note in notes:
            onset, _, pitch, velocity = note
            onset = int(onset * 100)
            time = onset + TIME_OFFSET
            assert time < PITCH_OFFSET, "Time token is out of range"
            tokens.extend([time, pitch + PITCH_OFFSET])
            if self.add_velocity:
                tokens.extend([velocity + VELOCITY_OFFSET])
        tokens.append(self.EOS_token)
        return torch.tensor(tokens)

    def empty_adt_tokens(self):
        return torch.tensor([self.BOS_token, self.silence_token, self.EOS_token])

    # ---- batch entries (what the main process calls once the dataset hands over notes instead of audio)
    def encode_batch(self, batch_notes: Sequence) -> List[torch.Tensor]:
        """One token tensor per segment; an empty note list gives ``empty_adt_tokens()`` (train_dataset.py:214-215)."""
        return [self.empty_adt_tokens() if len(n) == 0 else self.notes_to_adt_tokens(n) for n in batch_notes]

    @staticmethod
    def collate_tokens(token_list: Sequence) -> dict:
        """The token half of ``collate_fn`` (train_dataset.py:41-56): pad with 1, ``.long()``, and the lengths with the
        longest ones shortened by one."""
        lengths = np.array([len(t) for t in token_list], np.int64)
        width = int(lengths.max()) if len(lengths) else 0
        out = np.full((len(token_list), width), COLLATE_PAD_TOKEN, np.int64)
        for i, t in enumerate(token_list):
            a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
            out[i, : len(a)] = a.astype(np.int64)        # .long() truncates like astype
        if width > 0:
            lengths = np.where(lengths == width, lengths - 1, lengths)
        return {"tokens": torch.from_numpy(out), "token_lengths": torch.from_numpy(lengths)}

    # ---- midi_tokenizer.py:69-103
    def decode(self, tokens):
        """Token sequence -> ``(n, 4)`` rows ``[onset, onset + 0.1, pitch, velocity]``, as array operations.

        What the reference's loop does, stated as masks: a token in ``[TIME_OFFSET, PITCH_OFFSET)`` at position ``i``
        opens an onset; a pitch token counts only when position ``i - 1`` holds a time token, a velocity token only when
        ``i - 2`` does; BOS / EOS never count.  The three accepted streams are then paired IN ORDER and cut to the
        shortest (``zip`` over the dict values) - not matched by position - and without any accepted velocity every
        note gets 100.  Result dtypes follow what ``torch.tensor`` makes of the reference's per-token arithmetic:
        float tokens keep their dtype, integer tensors give float32, integer NumPy arrays float64, Python numbers
        float32 (double arithmetic, rounded once)."""
        if isinstance(tokens, torch.Tensor):
            arr, kind = tokens.detach().cpu().numpy(), "torch"
        elif isinstance(tokens, np.ndarray):
            arr, kind = tokens, "numpy"
        else:
            arr, kind = np.asarray(list(tokens)), "python"
        arr = arr.reshape(-1)
        special = (arr == self.BOS_token) | (arr == self.EOS_token)
        is_time = ~special & (arr >= TIME_OFFSET) & (arr < PITCH_OFFSET)
        before = lambda k: np.concatenate([np.zeros(min(k, len(arr)), bool), is_time[: max(len(arr) - k, 0)]])  # noqa: E731
        pitch_ok = ~special & (arr >= PITCH_OFFSET) & (arr < VELOCITY_OFFSET) & before(1)
        vel_ok = ~special & (arr >= VELOCITY_OFFSET) & before(2)
        if arr.dtype.kind == "f":
            work = arr.dtype                                   # tensor / NumPy-scalar arithmetic stays in the tokens' dtype
        elif kind == "torch":
            work = np.dtype(np.float32)                        # integer tensor / 100 -> torch's default float
        else:
            work = np.dtype(np.float64)                        # np.int64 / 100 and python int / 100 are doubles
        onset = (arr[is_time].astype(work) - work.type(TIME_OFFSET)) / work.type(100)
        pitch = arr[pitch_ok].astype(work) - work.type(PITCH_OFFSET)
        if self.ADTOF_mapping and len(pitch):
            if kind == "torch":
                raise KeyError(tokens[int(np.flatnonzero(pitch_ok)[0])] - PITCH_OFFSET)   # a tensor is not a key of the map
            pitch = np.array([self.ADTOF_map[int(v)] if float(v).is_integer() else self.ADTOF_map[v] for v in pitch], work)
        velocity = arr[vel_ok].astype(work) - work.type(VELOCITY_OFFSET) if vel_ok.any() else np.full(len(onset), 100, work)
        n = min(len(onset), len(pitch), len(velocity))
        if n == 0:
            return torch.tensor([])
        rows = np.stack([onset[:n], onset[:n] + work.type(0.1), pitch[:n], velocity[:n]], axis=1)
        if kind == "python":
            rows = rows.astype(np.float32)                     # torch.tensor(list of python floats)
        return torch.from_numpy(np.ascontiguousarray(rows))

    def batch_decode(self, tokens):
        return [self.decode(token) for token in tokens]
