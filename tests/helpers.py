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


def _floats(s):
    if isinstance(s, (int, float)):
        return (float(s),)
    return tuple(float(v) for v in str(s).split(","))


def oracle_specs_from_cfg(cfg, B=1):
    """(Tables, OperatorSpec, GuidanceSpec, initial phis) from a parsed upstream YAML."""
    d = cfg["diffusion"]
    tab = orc.make_tables(d["steps"], d["noise_schedule"], d.get("timestep_respacing", ""))
    o = cfg["measurement"]["operator"]
    kind = o["name"]
    val = _floats(o["value"])
    if kind == "underwater_physical_revised":
        names = ["phi_a", "phi_b", "phi_inf"]
    else:
        names = ["phi_ab", "phi_inf"]
    eta = tuple(float(o.get(n + "_eta", 1e-5)) if o.get(n + "_learn_flag", True) else 0.0 for n in names)
    phis = [torch.tensor(_floats(o[n]), dtype=torch.float32).repeat(B, 1)[..., None, None] for n in names]
    op = orc.OperatorSpec(kind, o.get("depth_type"), val if len(val) > 1 else val[0], eta)
    p, sp = cfg["conditioning"]["params"], cfg["sample_pattern"]
    clip = p.get("gradient_clip", "False").split(",")
    g = orc.GuidanceSpec(scale=_floats(p["scale"]), clip=float(clip[1]) if clip[0].strip().lower() == "true" else None,
                         loss_weight=p.get("loss_weight"), weight_fn=p.get("weight_function"),
                         aux=(cfg.get("aux_loss") or {}).get("aux_loss"), n_iter=sp["n_iter"],
                         update_start=sp["update_start"], update_end=sp["update_end"],
                         start_guidance=sp["start_guidance"], stop_guidance=sp["stop_guidance"], pattern=sp["pattern"])
    return tab, op, g, phis, names


def maxdiff(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max())


def rel_err(a, b):
    """max |a-b| / max |b| - the normalised max error used for all floating-point parity bars."""
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
