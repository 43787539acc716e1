"""Run the UNMODIFIED reference ``SynthDrum`` / ``ComputeMelSpectrogram`` from
``/root/reference`` (build container only - that tree does not exist on the GPU box).

TEST INFRASTRUCTURE ONLY.  ``modules/synthetiser.py`` imports ``h5py`` and
``pedalboard``, neither of which is installed, so two stand-in modules are put in
``sys.modules`` before the import (SURVEY §8c "shim recipe"):

* ``h5py.File(path, "r")`` -> context manager over an in-RAM nested dict
  ``bank[pitch][group][name] -> float32 array`` with sorted ``keys()``, path
  membership and ``[...]`` reads - the subset the synthetiser touches
  (``synthetiser.py:196,199,273,283-284``);
* ``pedalboard`` -> inert classes (FX are off: ``use_fx_prob = 0``).

Nothing from the reference is copied; it is imported where it lies.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict

import numpy as np

REFERENCE_ROOT = os.environ.get("ADT_REFERENCE_ROOT", "/root/reference")
_BANKS: Dict[str, dict] = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "synthetiser.py"))


class _Dataset:
    def __init__(self, data):
        self._d = data

    def __getitem__(self, _key):
        return np.array(self._d, dtype=np.float32, copy=True)


class _Group:
    def __init__(self, tree):
        self._t = tree

    def _walk(self, path):
        node = self._t
        for part in str(path).split("/"):
            if not isinstance(node, dict) or part not in node:
                raise KeyError(path)
            node = node[part]
        return node

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self._walk(path)
        return _Group(node) if isinstance(node, dict) else _Dataset(node)

    def keys(self):
        return sorted(self._t.keys())


class _File(_Group):
    def __init__(self, path, mode="r"):
        if path not in _BANKS:
            raise FileNotFoundError(path)
        super().__init__(_BANKS[path])

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _install_shims() -> None:
    if "h5py" not in sys.modules:
        m = types.ModuleType("h5py")
        m.File = _File
        sys.modules["h5py"] = m
    if "pedalboard" not in sys.modules:
        m = types.ModuleType("pedalboard")
        for name in ("Pedalboard", "Reverb", "Compressor", "Limiter"):
            setattr(m, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
        sys.modules["pedalboard"] = m


def _import_reference():
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    synth = importlib.import_module("modules.synthetiser")
    model = importlib.import_module("model")
    return synth, model


def enable_fx_stand_in():
    """Let the UNMODIFIED reference run with ``use_fx_prob > 0``: the names ``modules.synthetiser`` imported from
    ``pedalboard`` (``synthetiser.py:11``) are pointed at the recording stand-in of ``oracle/fx_oracle.py`` - plugin
    classes that log their constructor arguments and, when the board is called, run the CPU restatement of the JUCE
    DSP.  Nothing in the reference tree is touched; returns ``fx_oracle`` (its ``CALLS`` list is the log)."""
    from . import fx_oracle
    synth, _ = _import_reference()
    for name in ("Pedalboard", "Reverb", "Compressor", "Limiter"):
        setattr(synth, name, getattr(fx_oracle, name))
    return fx_oracle


def register_bank(oneshot_path: str, sample_rate: int, nested: dict) -> None:
    """Make ``f"{oneshot_path}@{sample_rate}.hdf5"`` resolve to ``nested`` (synthetiser.py:163)."""
    _BANKS[f"{oneshot_path}@{sample_rate}.hdf5"] = nested


def make_synth(cfg_dict: dict, nested_bank: dict):
    """The reference's ``SynthDrum`` over an in-RAM bank."""
    synth, _ = _import_reference()
    register_bank(cfg_dict["oneshot_path"], cfg_dict["sample_rate"], nested_bank)
    return synth.SynthDrum(synth.SynthDrumConfig(**cfg_dict))


def make_mel(sample_rate: int, win_length: int, time_res: float, n_mels: int):
    """The reference's ``ComputeMelSpectrogram`` (model.py:68-97)."""
    _, model = _import_reference()
    return model.ComputeMelSpectrogram(sample_rate, win_length, time_res, n_mels)


def import_audio_front():
    """The reference's ``utils.audio_utils`` (resample / normalize, :17-23) and ``inference`` (``_chunk_audio``,
    :35-48) modules, unmodified.  ``inference.py`` imports packages that are not installed here (pretty_midi,
    omegaconf ...); empty stand-in modules let the import through - none of them is touched by ``_chunk_audio``."""
    import importlib
    import importlib.machinery
    _import_reference()  # transformers (via model.py) must be imported before the stand-ins exist
    for name in ("pretty_midi", "omegaconf", "mir_eval", "soundfile", "librosa", "matplotlib", "accelerate"):
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = object
    audio_utils = importlib.import_module("utils.audio_utils")
    inference = importlib.import_module("inference")
    return audio_utils, inference


def import_tokenizer():
    """The reference's ``modules.midi_tokenizer`` module and ``data_modules.train_dataset.collate_fn``, unmodified
    (``collate_fn`` is None when the dataset module cannot be imported here)."""
    import importlib
    _import_reference()
    tok = importlib.import_module("modules.midi_tokenizer")
    try:
        import_audio_front()    # stand-ins for the uninstalled packages the dataset module pulls in
        argv, sys.argv = sys.argv, ["train_dataset", "unused.yaml"]   # the module parses argv at import (:232-234)
        try:
            collate = importlib.import_module("data_modules.train_dataset").collate_fn
        finally:
            sys.argv = argv
    except Exception:
        collate = None
    return tok, collate
