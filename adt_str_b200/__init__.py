"""B200-native render + log-mel front end with the call signatures of
pier-maker92/ADT_STR's ``SynthDrum`` (``modules/synthetiser.py``) and
``ComputeMelSpectrogram`` (``model.py:68-97``).  See DESIGN.md."""
from .config import SharedConfig, SynthDrumConfig, setting_1, config_default  # noqa: F401
from .bank import OneShotBank  # noqa: F401

__all__ = ["SharedConfig", "SynthDrumConfig", "OneShotBank", "setting_1", "config_default",
           "SynthDrum", "ComputeMelSpectrogram", "FrontEnd", "HostPipeline", "Resample", "LongFormFrontEnd", "MidiTokenizer", "MidiTokenizerConfig", "ProjectToMel"]


def __getattr__(name):  # lazy: importing the package must not need the CUDA library
    if name == "SynthDrum":
        from .synthetiser import SynthDrum
        return SynthDrum
    if name == "ComputeMelSpectrogram":
        from .mel import ComputeMelSpectrogram
        return ComputeMelSpectrogram
    if name == "FrontEnd":
        from .frontend import FrontEnd
        return FrontEnd
    if name == "HostPipeline":
        from .pipeline import HostPipeline
        return HostPipeline
    if name == "Resample":
        from .audio_utils import Resample
        return Resample
    if name == "LongFormFrontEnd":
        from .inference_front import LongFormFrontEnd
        return LongFormFrontEnd
    if name == "ProjectToMel":
        from .projection import ProjectToMel
        return ProjectToMel
    if name in ("MidiTokenizer", "MidiTokenizerConfig"):
        from . import midi_tokenizer
        return getattr(midi_tokenizer, name)
    raise AttributeError(name)
