"""Shared test helpers: golden loading, YAML -> oracle specs, tolerant comparison."""
import json
import os

import numpy as np
import torch
import yaml

from oracle import osmosis_oracle as orc
from osmosis_diffusion_code_b200.synthetic import synth_state_dict
from tests.golden.cases import SMALL_UNET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = dict(np.load(os.path.join(GOLD, "small_golden.npz")))
    return _golden


def small_specs():
    return [(k, tuple(s)) for k, s in json.load(open(os.path.join(GOLD, "small_unet_param_specs.json")))]


_sd = None


def small_state_dict():
    global _sd
    if _sd is None:
        _sd = synth_state_dict(small_specs(), SMALL_UNET["num_channels"], seed=7, delta=0.05)
    return _sd


def small_cfg():
    return orc.UNetConfig.from_create_model_kwargs(**SMALL_UNET)


def load_yaml_cfg(name, respacing=None):
    cfg = yaml.load(open(os.path.join(ROOT, "configs", name)), Loader=yaml.FullLoader)
    if respacing is not None:
        cfg["diffusion"]["timestep_respacing"] = respacing
    return cfg


oracle_specs_from_cfg = orc.specs_from_config


def maxdiff(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max())


def rel_err(a, b):
    """max |a-b| / max |b| - the normalised max error used for all floating-point parity bars."""
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
