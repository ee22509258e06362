"""Measurement operators (underwater / haze image-formation models with learnable phi) and noise models.

Mirrors the registries and classes of the reference `guided_diffusion/measurements.py`:
`get_operator` (:30-38, also stamps `__name__`), `HazePhysicalOperator` (:107-208),
`UnderWaterPhysicalRevisedOperator` (:211-329), `UnderWaterPhysicalOperator` (:332-433), `get_noise`
(:454-459), `Clean` (:471-474), `GaussianNoise` (:477-483), and the parameter-free `DenoiseOperator` (:61-77) /
`RGBGuidanceOperator` (:80-97) of the rgb_guidance demo.

Difference in mechanics, not in results: the water parameters of all images live in ONE device tensor
`phi[B, 9] = {a | b | inf}` that the fused guidance kernel (osm_guidance_phi_loop) reads and updates in place;
`phi_a`, `phi_b`, `phi_inf`, `phi_ab` are views of it with the reference's [B, c, 1, 1] shapes.  The SGD step
(`optimizer: sgd` / `GD` / `adam`, lr = phi_*_eta, 0 when the learn flag is off; :240-249, :266-303, utils.py:494-500) is part
of that kernel, so `optimize()` only reports the current values.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import lib as _lib
from ..osmosis_utils import utils as utilso

__OPERATOR__ = {}
__NOISE__ = {}

OP_KIND = {"underwater_physical_revised": 0, "underwater_physical": 1, "haze_physical": 2}
DEPTH_KIND = {None: 0, "original": 0, "gamma": 1, "move": 2}


def register_operator(name: str):
    def wrapper(cls):
        if __OPERATOR__.get(name, None):
            raise NameError(f"Name {name} is already registered!")
        __OPERATOR__[name] = cls
        return cls
    return wrapper


def get_operator(name: str, **kwargs):
    if __OPERATOR__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined.")
    operator = __OPERATOR__[name](**kwargs)
    operator.__name__ = name
    return operator


class LinearOperator:
    """measurements.py:41-58: operators without parameters."""

    def forward(self, data, **kwargs):
        raise NotImplementedError

    def transpose(self, data, **kwargs):
        raise NotImplementedError

    def ortho_project(self, data, **kwargs):
        return data - self.transpose(self.forward(data, **kwargs), **kwargs)

    def project(self, data, measurement, **kwargs):
        return self.ortho_project(measurement, **kwargs) - self.forward(data, **kwargs)


@register_operator(name="noise")
class DenoiseOperator(LinearOperator):
    """measurements.py:61-77: identity."""

    def __init__(self, device, **kwargs):
        self.device = device

    def forward(self, data, **kwargs):
        return data

    def transpose(self, data):
        return data

    def ortho_project(self, data):
        return data

    def project(self, data):
        return data


@register_operator(name="rgb_guidance")
class RGBGuidanceOperator(DenoiseOperator):
    """measurements.py:80-97: identity on the RGB channels the `ps` conditioning hands it.  The fused loop recognises it
    (`is_identity`) and runs osm_ps_guidance, which has the operator folded in."""
    is_identity = True

    def __init__(self, device, batch_size=1, **kwargs):
        self.device = device
        self.batch_size = batch_size


def depth_spec(depth_type, value):
    """(kind, [v0, v1, v2]) for the kernels from the YAML's depth_type / value (utils.py:529-566)."""
    if depth_type not in DEPTH_KIND:
        raise NotImplementedError
    v = utilso.get_depth_value(value) if value is not None else 0.0
    v = np.atleast_1d(np.asarray(v, dtype=np.float64)).tolist()
    v = (v + [0.0, 1.0, 1.0][len(v):])[:3] if len(v) < 3 else v[:3]
    return DEPTH_KIND[depth_type], [float(x) for x in v]


def _as_floats(v):
    if isinstance(v, str):
        return np.fromstring(v, dtype=float, sep=",").tolist()
    return np.atleast_1d(np.asarray(v, dtype=np.float64)).tolist()


class LearnableOperator:
    """Common state of the three physical operators."""
    kind_name = None
    groups = ()  # variable names in get_variable_list() order

    def _setup(self, device, batch_size, init, etas, kwargs):
        _lib.require_cuda()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.OsmError("operators run on CUDA devices only (no CPU fallback)")
        self.batch_size = batch_size
        self.depth_type = kwargs.get("depth_type", None)
        self.value = utilso.get_depth_value(kwargs.get("value", None)) if kwargs.get("value", None) is not None else None
        self.depth_kind, self.depth_val = depth_spec(self.depth_type, kwargs.get("value", None))
        optimizer = kwargs.get("optimizer", None)
        if optimizer is None:
            raise AttributeError("'NoneType' object has no attribute 'lower'")  # utils.py:495 behaviour: key must exist
        if optimizer.lower() not in ("", "gd", "sgd", "adam"):
            raise ValueError(f"Optimizer '{optimizer}' is not supported by the fused phi update (sgd / GD / adam).")
        self.optimizer = optimizer
        # torch.optim.Adam's per-parameter state (exp_avg[9], exp_avg_sq[9], step), kept across conditioning calls like the
        # reference's optimizer object (measurements.py:240-249)
        self.opt_state = torch.zeros(batch_size, 19, dtype=torch.float32, device=self.device) if optimizer.lower() == "adam" else None
        row = torch.zeros(9, dtype=torch.float32)
        for off, vals in init:
            row[off:off + len(vals)] = torch.tensor(vals, dtype=torch.float32)
        self.phi = row.repeat(batch_size, 1).to(self.device).contiguous()
        self.eta = [float(e) for e in etas]
        self._requires_grad = {g: False for g in self.groups}

    # views with the reference's shapes -------------------------------------------------------------
    def _view(self, lo, hi):
        return self.phi[:, lo:hi].unsqueeze(-1).unsqueeze(-1)

    def forward(self, data, **kwargs):
        B, Cc, H, W = data.shape
        assert Cc == 4 and B == self.phi.shape[0]
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=data.device)
        dv = (C.c_float * 3)(*self.depth_val)
        L = _lib.load()
        _lib.check(L.osm_operator_forward(OP_KIND[self.kind_name], self.depth_kind, dv, _lib.ptr(data.detach().contiguous()),
                                          _lib.ptr(self.phi), _lib.ptr(out), B, H * W, _lib.stream()))
        return out

    def optimize(self, **kwargs):
        return {g: getattr(self, g).detach() for g in self.groups}

    def get_variable_gradients(self, **kwargs):
        return dict(self._requires_grad)

    def set_variable_gradients(self, value=None, **kwargs):
        if value is None:
            raise ValueError("A value should be specified (True or False for general or dictionary)")
        for g in self.groups:
            self._requires_grad[g] = bool(value[g] if isinstance(value, dict) else value)

    def get_variable_list(self, **kwargs):
        return [getattr(self, g) for g in self.groups]


@register_operator(name="underwater_physical_revised")
class UnderWaterPhysicalRevisedOperator(LearnableOperator):
    kind_name = "underwater_physical_revised"
    groups = ("phi_a", "phi_b", "phi_inf")

    def __init__(self, device, phi_a, phi_b, phi_inf, phi_a_eta=1e-5, phi_b_eta=1e-5, phi_inf_eta=1e-5,
                 phi_a_learn_flag=True, phi_b_learn_flag=True, phi_inf_learn_flag=True, batch_size=1, **kwargs):
        self.phi_a_learn_flag, self.phi_b_learn_flag, self.phi_inf_learn_flag = phi_a_learn_flag, phi_b_learn_flag, phi_inf_learn_flag
        self.phi_a_eta = float(phi_a_eta) if phi_a_learn_flag else 0.0
        self.phi_b_eta = float(phi_b_eta) if phi_b_learn_flag else 0.0
        self.phi_inf_eta = float(phi_inf_eta) if phi_inf_learn_flag else 0.0
        self._setup(device, batch_size, [(0, _as_floats(phi_a)), (3, _as_floats(phi_b)), (6, _as_floats(phi_inf))],
                    [self.phi_a_eta, self.phi_b_eta, self.phi_inf_eta], kwargs)

    phi_a = property(lambda self: self._view(0, 3))
    phi_b = property(lambda self: self._view(3, 6))
    phi_inf = property(lambda self: self._view(6, 9))


class _TiedOperator(LearnableOperator):
    groups = ("phi_ab", "phi_inf")
    _n_ab = 3

    def __init__(self, device, phi_ab, phi_inf, phi_ab_eta=1e-5, phi_inf_eta=1e-5, phi_ab_learn_flag=True,
                 phi_inf_learn_flag=True, batch_size=1, **kwargs):
        self.phi_ab_learn_flag, self.phi_inf_learn_flag = phi_ab_learn_flag, phi_inf_learn_flag
        self.phi_ab_eta = float(phi_ab_eta) if phi_ab_learn_flag else 0.0
        self.phi_inf_eta = float(phi_inf_eta) if phi_inf_learn_flag else 0.0
        ab = [float(phi_ab)] if self._n_ab == 1 else _as_floats(phi_ab)
        self._setup(device, batch_size, [(0, ab), (6, _as_floats(phi_inf))], [self.phi_ab_eta, self.phi_inf_eta, 0.0], kwargs)

    phi_ab = property(lambda self: self._view(0, self._n_ab))
    phi_inf = property(lambda self: self._view(6, 9))


@register_operator(name="underwater_physical")
class UnderWaterPhysicalOperator(_TiedOperator):
    kind_name = "underwater_physical"


@register_operator(name="haze_physical")
class HazePhysicalOperator(_TiedOperator):
    kind_name = "haze_physical"
    _n_ab = 1


# ------------------------------------------------------------------------------------------------ noise


def register_noise(name: str):
    def wrapper(cls):
        if __NOISE__.get(name, None):
            raise NameError(f"Name {name} is already defined!")
        __NOISE__[name] = cls
        return cls
    return wrapper


def get_noise(name: str, **kwargs):
    if __NOISE__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined.")
    noiser = __NOISE__[name](**kwargs)
    noiser.__name__ = name
    return noiser


class Noise:
    def __call__(self, data):
        return self.forward(data)


@register_noise(name="clean")
class Clean(Noise):
    def forward(self, data):
        return data


@register_noise(name="gaussian")
class GaussianNoise(Noise):
    """measurements.py:477-483: y + sigma * randn (drawn once per image, before the sampling loop is seeded)."""

    def __init__(self, sigma):
        self.sigma = sigma

    def forward(self, data):
        return data + torch.randn_like(data) * self.sigma
