"""Packed one-shot bank.

The reference keeps one gzip float32 HDF5 dataset per one-shot under
``<pitch>/<similarity-group>/<name>`` and re-opens the file for every note
(``modules/synthetiser.py:163,273-284``; layout written by
``data_modules/convert_augmented_to_hdf5.py:70-138``).  Here the whole bank is
one flat float32 array that lives in HBM for the life of the process:

* ``pcm``      float32 (total,)  every one-shot, start padded to 32 floats (128 B)
                                 so bulk copies and float4 loads stay aligned
* ``offsets``  int64   (n,)      start of one-shot ``i`` in ``pcm`` (floats)
* ``lengths``  int32   (n,)      true length in samples
* ``index``    {(pitch, group): (first_id, count)} - ids inside one
                                 (pitch, group) are contiguous and ordered by
                                 *sorted name*, which is the order
                                 ``list(h5_group.keys())`` yields, so
                                 ``random.choice`` picks the same one-shot.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Mapping, Tuple

import numpy as np

ALIGN = 32  # floats


class OneShotBank:
    def __init__(self, pcm: np.ndarray, offsets: np.ndarray, lengths: np.ndarray,
                 index: Dict[Tuple[int, str], Tuple[int, int]], names: List[str]):
        self.pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        self.index = dict(index)
        self.names = list(names)
        if not (len(self.offsets) == len(self.lengths) == len(self.names)):
            raise ValueError("bank arrays disagree on the number of one-shots")
        if len(self.offsets) and (self.offsets % 4).any():
            raise ValueError("one-shot starts must be 16-byte aligned")
        # pitches that have at least one group, for the '"<pitch>/<group>" in file' test
        self._pitches = {p for (p, _g) in self.index}

    # ------------------------------------------------------------------ build
    @classmethod
    def from_nested(cls, nested: Mapping[str, Mapping[str, Mapping[str, np.ndarray]]]) -> "OneShotBank":
        """``nested[pitch][group][name] -> 1-D float32`` (the HDF5 tree shape)."""
        chunks, offsets, lengths, names, index = [], [], [], [], {}
        cursor = 0
        for pitch in sorted(nested, key=lambda s: int(s)):
            for group in sorted(nested[pitch]):
                members = nested[pitch][group]
                first = len(names)
                for name in sorted(members):
                    x = np.asarray(members[name], dtype=np.float32).reshape(-1)
                    padded = -(-len(x) // ALIGN) * ALIGN
                    buf = np.zeros(padded, np.float32)
                    buf[: len(x)] = x
                    chunks.append(buf)
                    offsets.append(cursor)
                    lengths.append(len(x))
                    names.append(f"{int(pitch)}/{group}/{name}")
                    cursor += padded
                index[(int(pitch), group)] = (first, len(names) - first)
        pcm = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
        return cls(pcm, np.array(offsets, np.int64), np.array(lengths, np.int32), index, names)

    @classmethod
    def from_hdf5(cls, path: str) -> "OneShotBank":
        """Convert the reference's HDF5 bank.  Needs h5py, which this image lacks;
        run it where the bank was built."""
        try:
            import h5py  # type: ignore
        except ImportError as e:  # pragma: no cover - h5py absent here
            raise ImportError("h5py is required to convert an HDF5 one-shot bank") from e
        nested: Dict[str, Dict[str, Dict[str, np.ndarray]]] = {}
        with h5py.File(path, "r") as f:
            for pitch in f.keys():
                if pitch == "index":
                    continue
                for group in f[pitch].keys():
                    for name in f[pitch][group].keys():
                        nested.setdefault(pitch, {}).setdefault(group, {})[name] = f[pitch][group][name][...]
        return cls.from_nested(nested)

    # ------------------------------------------------------------------- io
    def save(self, path: str) -> None:
        meta = {"index": [[p, g, a, n] for (p, g), (a, n) in self.index.items()], "names": self.names}
        np.savez(path, pcm=self.pcm, offsets=self.offsets, lengths=self.lengths,
                 meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))

    @classmethod
    def load(cls, path: str) -> "OneShotBank":
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())
        index = {(int(p), g): (int(a), int(n)) for p, g, a, n in meta["index"]}
        return cls(z["pcm"], z["offsets"], z["lengths"], index, meta["names"])

    # --------------------------------------------------------------- queries
    def __len__(self) -> int:
        return len(self.lengths)

    def has_group(self, pitch: int, group: str) -> bool:
        return (pitch, group) in self.index

    def group_range(self, pitch: int, group: str) -> Tuple[int, int]:
        return self.index[(pitch, group)]

    def oneshot(self, i: int) -> np.ndarray:
        o = int(self.offsets[i])
        return self.pcm[o: o + int(self.lengths[i])]

    def to_nested(self) -> Dict[str, Dict[str, Dict[str, np.ndarray]]]:
        out: Dict[str, Dict[str, Dict[str, np.ndarray]]] = {}
        for i, full in enumerate(self.names):
            pitch, group, name = full.split("/", 2)
            out.setdefault(pitch, {}).setdefault(group, {})[name] = self.oneshot(i).copy()
        return out

    @property
    def nbytes(self) -> int:
        return self.pcm.nbytes


def _main(argv=None) -> int:
    """``python -m adt_str_b200.bank <oneshot@SR.hdf5> <oneshot@SR.npz>``: the reference's HDF5 bank
    (``data_modules/convert_augmented_to_hdf5.py:70-138``) -> the packed form ``SynthDrum`` loads.  Needs h5py (run it
    where the bank was built)."""
    import argparse
    ap = argparse.ArgumentParser(prog="python -m adt_str_b200.bank", description=_main.__doc__)
    ap.add_argument("hdf5")
    ap.add_argument("npz")
    args = ap.parse_args(argv)
    bank = OneShotBank.from_hdf5(args.hdf5)
    bank.save(args.npz)
    print(f"{len(bank)} one-shots, {bank.nbytes / 1e6:.1f} MB of PCM, {len(bank.index)} (pitch, group) cells -> {args.npz}")
    return 0


if __name__ == "__main__":
    raise SystemExit(_main())
