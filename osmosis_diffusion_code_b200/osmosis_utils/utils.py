"""Host-side helpers of the sampling path: config loading and the sampling-pattern phase logic.

Mirrors, from the reference `osmosis_utils/utils.py`: `load_yaml` / `arguments_from_file` (:357-360, :466-476),
`str2bool` (:384-395), `get_depth_value` (:529-541), `is_freeze_phi` (:571-590), `set_alternate_length`
(:595-630).  Everything image / visualisation related in that file is out of scope.
"""
from __future__ import annotations

import argparse

import numpy as np
import yaml


def load_yaml(file_path: str) -> dict:
    with open(file_path) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def arguments_from_file(config_file_path: str) -> argparse.Namespace:
    """Top-level YAML keys become attributes; nested mappings stay dicts (they are splatted into factories)."""
    ns = argparse.Namespace()
    for key, value in load_yaml(config_file_path).items():
        setattr(ns, key, value)
    return ns


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = v.lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")


def get_depth_value(value_raw, **kwargs):
    if isinstance(value_raw, float):
        return value_raw
    if isinstance(value_raw, int):
        return float(value_raw)
    if isinstance(value_raw, str):
        return np.fromstring(value_raw, dtype=float, sep=",")
    if isinstance(value_raw, (np.ndarray, np.generic)):
        return value_raw
    raise NotImplementedError


def _outside(sample_pattern, lo_key, hi_key, time_index, num_timesteps):
    return time_index > sample_pattern[hi_key] * num_timesteps or time_index < sample_pattern[lo_key] * num_timesteps


def is_freeze_phi(sample_pattern, time_index, num_timesteps):
    """phi is frozen outside the guidance window and outside [update_end, update_start] * T."""
    if sample_pattern is None or sample_pattern["pattern"] == "original":
        return False
    if _outside(sample_pattern, "stop_guidance", "start_guidance", time_index, num_timesteps):
        return True
    return _outside(sample_pattern, "update_end", "update_start", time_index, num_timesteps)


def set_alternate_length(sample_pattern, time_index, num_timesteps):
    """Number of x / phi alternations at this step (gibbsDDRM's M); 1 outside [s_end, s_start] * T."""
    if sample_pattern is None or sample_pattern["pattern"] == "original":
        return 1
    assert sample_pattern["update_start"] > sample_pattern["update_end"]
    assert sample_pattern["s_start"] > sample_pattern["s_end"]
    if sample_pattern["local_M"] > 1:
        assert sample_pattern["update_start"] >= sample_pattern["s_start"]
        assert sample_pattern["s_end"] >= sample_pattern["update_end"]
    for lo, hi in (("stop_guidance", "start_guidance"), ("update_end", "update_start"), ("s_end", "s_start")):
        if _outside(sample_pattern, lo, hi, time_index, num_timesteps):
            return 1
    return sample_pattern["local_M"]
