"""Auxiliary guidance losses (registry + weights).

Mirrors the reference `osmosis_utils/losses.py`: `get_loss` registry (:8-24), `avrg_loss` = sum_c |mean_hw rgb_c|
(:29-45), `val_loss` = mean relu(|rgb| - 0.7)^2 (:51-62) and the weighted sum `AuxiliaryLoss` (:67-83).
On the sampling path the two terms and their gradients are evaluated inside the fused guidance kernel
(osm_guidance_phi_loop); this module carries their names and weights to it.  Per-image semantics: every
image's terms are its own (a batch is B independent reference runs).
"""
from __future__ import annotations

__LOSS__ = {}


def register_loss(name: str):
    def wrapper(cls):
        if __LOSS__.get(name, None):
            raise NameError(f"Name {name} is already registered!")
        __LOSS__[name] = cls
        return cls
    return wrapper


def get_loss(name: str, **kwargs):
    if __LOSS__.get(name, None) is None:
        raise NameError(f"Name {name} is not defined.")
    return __LOSS__[name](**kwargs)


@register_loss(name="avrg_loss")
class Average_Loss:
    kernel_slot = "gamma_avrg"


@register_loss(name="val_loss")
class Value_Loss:
    kernel_slot = "gamma_val"
    value = 0.7


class AuxiliaryLoss:
    def __init__(self, losses_dictionary):
        self.losses_dictionary = dict(losses_dictionary)
        self.losses_list = [get_loss(k) for k in losses_dictionary]
        self.loss_gammas = [float(v) for v in losses_dictionary.values()]

    def kernel_weights(self):
        """{'gamma_avrg': w, 'gamma_val': w} for osm_guidance_params (0 for absent terms)."""
        out = {"gamma_avrg": 0.0, "gamma_val": 0.0}
        for loss, gamma in zip(self.losses_list, self.loss_gammas):
            out[loss.kernel_slot] = gamma
        return out
