"""Configuration dataclasses for the render + log-mel front end.

Field names and order mirror the reference so that
``SynthDrumConfig(**synthetiser_section, **shared_section, ADTOF_mapping=...)``
(reference ``train.py:275-281``, ``inference.py:140-145``) keeps working:

* ``SharedConfig``      <- reference ``config.py:8-13``
* ``SynthDrumConfig``   <- reference ``modules/synthetiser.py:14-27``
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Mapping

import yaml


@dataclass
class SharedConfig:
    input_sec: float
    time_res: float
    win_length: int
    sample_rate: int


@dataclass
class SynthDrumConfig(SharedConfig):
    oneshot_path: str
    similarity_threshold: float
    max_hat_std_velocity: float
    max_hat_mean_velocity: float
    max_cymbals_std_velocity: float
    max_cymbals_mean_velocity: float
    ADTOF_mapping: bool
    mixup_range: float
    use_fx_prob: float
    use_reverb_prob: float
    use_limiter_prob: float
    use_compression_prob: float


#: the ``synthetiser:`` + ``shared:`` blocks of reference ``configs/train/setting-1.yaml:29-33,50-61``
#: with the FX chain off: the measured path of SURVEY §8d / BASELINE.json runs with ``use_fx_prob = 0`` (the yaml has
#: 0.3; ``setting_1(use_fx_prob=0.3)`` or ``synth_config_from_sections`` on the yaml gives the training distribution -
#: the FX chain runs on the GPU, csrc/fx.cu).
SETTING_1 = dict(
    input_sec=2.56, time_res=0.01, win_length=2048, sample_rate=24000,
    oneshot_path="oneshot", similarity_threshold=0.8,
    max_hat_std_velocity=0.15, max_hat_mean_velocity=0.1,
    max_cymbals_std_velocity=0.15, max_cymbals_mean_velocity=0.65,
    ADTOF_mapping=False, mixup_range=0.8, use_fx_prob=0.0,
    use_reverb_prob=0.5, use_limiter_prob=0.5, use_compression_prob=0.5,
)

#: ``shared:`` of reference ``configs/config_default.yaml:11-15`` (sr 16 kHz); the
#: default config has no ``synthetiser:`` block, so setting-1's is used.
CONFIG_DEFAULT = dict(SETTING_1, sample_rate=16000)


def setting_1(**overrides: Any) -> SynthDrumConfig:
    return SynthDrumConfig(**{**SETTING_1, **overrides})


def config_default(**overrides: Any) -> SynthDrumConfig:
    return SynthDrumConfig(**{**CONFIG_DEFAULT, **overrides})


def deep_merge(base: Dict[str, Any], override: Mapping[str, Any]) -> Dict[str, Any]:
    """Nested dict merge, ``override`` wins (reference ``utils/config_utils.py:4-14``
    does this through OmegaConf, which is not a dependency here)."""
    out = dict(base)
    for k, v in override.items():
        if isinstance(v, Mapping) and isinstance(out.get(k), Mapping):
            out[k] = deep_merge(dict(out[k]), v)
        else:
            out[k] = v
    return out


def load_yaml_config(default_path: str, experiment_path: str | None = None) -> Dict[str, Any]:
    with open(default_path) as f:
        cfg = yaml.safe_load(f) or {}
    if experiment_path:
        with open(experiment_path) as f:
            cfg = deep_merge(cfg, yaml.safe_load(f) or {})
    return cfg


def synth_config_from_sections(cfg: Mapping[str, Any]) -> SynthDrumConfig:
    """``synthetiser ∪ shared ∪ {ADTOF_mapping}`` as reference ``train.py:274-281``."""
    section = dict(cfg.get("synthetiser", {}))
    section.update(cfg.get("shared", {}))
    section.setdefault("ADTOF_mapping", cfg.get("tokenizer", {}).get("ADTOF_mapping", False))
    return SynthDrumConfig(**section)
